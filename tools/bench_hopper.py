"""Timing of b200gym_hopper_torques (Hopper._compute_torques) with CUDA events; 188 algorithmic bytes per env (csrc/hopper.cu header).
Usage (GPU box): python tools/bench_hopper.py > gpurun_out/hopper_bench.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from legged_gym_dev_b200.hopper import HopperActuation   # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6535.7)
out = {"peak_gbs": PEAK, "bytes_per_env": 188, "sizes": {}}
for N in (4096, 65536, 1 << 20, 1 << 22):
    env = HopperActuation(N, device="cuda", torque_limits=[9000.0, 80.0, 80.0, 80.0])
    g = torch.Generator(device="cuda").manual_seed(1)
    env.dof_state.normal_(generator=g)
    env.root_states.normal_(generator=g)
    env.base_ang_vel.normal_(generator=g)
    env.contact_forces.normal_(generator=g)
    act = torch.randn(N, 4, device="cuda", generator=g)
    for _ in range(5):
        env._compute_torques(act)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    a.record()
    for _ in range(reps):
        env._compute_torques(act)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    out["sizes"][N] = dict(ms=ms, gbs=188 * N / ms / 1e6, frac=188 * N / ms / 1e6 / PEAK, env_calls_per_s=N / ms * 1e3)
    # compute_observations: 96 B read + 84 B written per env
    env.base_lin_vel.normal_(generator=g)
    env.actions.copy_(act)
    for _ in range(5):
        env.compute_observations()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        env.compute_observations()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    out["sizes"][N]["observations"] = dict(ms=ms, gbs=180 * N / ms / 1e6, frac=180 * N / ms / 1e6 / PEAK)
print(json.dumps(out))
