"""A few AnymalTrajectory env steps at a given size (target for ncu captures of the trajectory-env kernels)."""
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import bench_configs as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
r = B.trajectory_env(num_envs=n, steps=3, warmup=2)
print(n, r["ms_per_step"])
