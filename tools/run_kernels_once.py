"""Runs each secondary kernel a few times at a representative size (for ncu captures; see profiles/)."""
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import torch
import bench_configs as B
r = B.rough_lstm(num_envs=262144, steps=3, warmup=2)
print("rough", r["ms_per_step"])
r = B.rom_rollout(num_envs=1 << 18, T=40)
print("rollout", r["ms_epoch"])
r = B.rom_per_call(num_envs=65536, loop_steps=20, cpu=False)
print("rom per call", r["ms_total"])
r = B.gae_update(num_envs=65536, T=24)
print("gae", r["gae_ms"], r["actor_forward"])
r = B.tube_dataset(num_envs=65536)
print("tube dataset", r["ms"])
r = B.trajectory_env(num_envs=262144, steps=3, warmup=2)
print("trajectory env", r["ms_per_step"])
