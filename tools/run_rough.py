import sys; sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import bench_configs as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
r = B.rough_lstm(num_envs=n, steps=10, warmup=3)
print(n, "ms/step %.4f" % r["ms_per_step"], "lstm", r["lstm_torques"], "pp", r["post_physics_rough"])
