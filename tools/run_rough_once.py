"""A few rough-terrain env steps at a given size (target for ncu captures of post_physics_kernel<32,true> / lstm_torques)."""
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import bench_configs as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
r = B.rough_lstm(num_envs=n, steps=2, warmup=1)
print(n, r["ms_per_step"])
