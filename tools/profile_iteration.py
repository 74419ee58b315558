"""Kernel list of one cfg-5 training iteration (4096 envs x 24 steps, graphed rollout + GAE + graphed update) from the torch profiler:
name, launches, total and mean device time — written as JSON to stdout (profiles/r2_iteration_kernels.json)."""
import json
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from legged_gym_dev_b200 import synthetic as S                                   # noqa: E402
from legged_gym_dev_b200.physics import ReplayPhysics                            # noqa: E402
from legged_gym_dev_b200.task_registry import task_registry                      # noqa: E402
from torch.profiler import profile, ProfilerActivity                            # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tape = S.make_state_tape(N, frames=4, seed=10, device="cuda")
args = SimpleNamespace(num_envs=N, sim_device="cuda", headless=True, physics_engine=None)
env, _ = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device="cuda"))
runner, _ = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
runner.graph_rollout = True
runner.learn(num_learning_iterations=3, init_at_random_ep_len=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    runner.learn(num_learning_iterations=1)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    runner.learn(num_learning_iterations=1)
    torch.cuda.synchronize()
rows = [dict(name=e.key[:110], count=e.count, total_us=round(e.device_time_total, 1), mean_us=round(e.device_time_total / max(e.count, 1), 2))
        for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r["total_us"])
print(json.dumps(dict(envs=N, ms_per_iteration=ms, kernel_time_sum_us=round(sum(r["total_us"] for r in rows), 1),
                      launches=sum(r["count"] for r in rows), kernels=rows), indent=1))
