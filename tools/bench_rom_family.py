"""Timing of the generic rom-family kernels (csrc/rom_family.cu) at BASELINE cfg 4's size, CUDA events on the launching stream.
Usage (GPU box): python tools/bench_rom_family.py [num_envs] > gpurun_out/rom_family_bench.json
Algorithmic bytes per env (fp32, n states, m inputs, W = N*dN window rows, P = 9m + 6 generator parameters incl. weights):
  f       : read z, v; write z_next                                  4 (2n + m)
  reset   : read z, parameters; write parameters, both windows       4 (n + 2P + (W+1) n + W m)
  step    : every call: read parameters + last knot, write t, v      4 (P + n + 1 + m)
            due calls (1 in rom.dt / dt_loop): + shift both windows   + 8 ((W+1) n + W m)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_dev_b200 import rom as R   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6535.7)
SPEC = {"SingleInt2D": ([1e9] * 2, [1.0] * 2), "DoubleInt2D": ([1e9, 1e9, 0.6, 0.6], [1.0] * 2), "Unicycle": ([1e9] * 3, [1.0, 2.0]),
        "LateralUnicycle": ([1e9] * 3, [1.0, 0.5, 2.0]), "ExtendedUnicycle": ([1e9, 1e9, 1e9, 1.0, 2.0], [1.0, 4.0]),
        "ExtendedLateralUnicycle": ([1e9, 1e9, 1e9, 1.0, 0.5, 2.0], [1.0, 0.6, 4.0])}


def timed(fn, reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {"num_envs": N, "peak_gbs": PEAK, "classes": {}}
Wn, dt_loop, rom_dt = 10, 0.02, 0.1
for cls, (zmax, vmax) in SPEC.items():
    rom = R.ROM_CLASSES[cls](rom_dt, [-a for a in zmax], zmax, [-a for a in vmax], vmax, n_robots=N, backend="torch", device="cuda")
    gen = R.TrajectoryGenerator(rom, R.UniformSampleHoldDT(1.0, 2.0), R.UniformWeightSampler(), dt_loop=dt_loop, N=Wn, freq_low=0.01, freq_high=2.0,
                                seed=1, device="cuda", prob_stationary=0.0005, generic_kernels=True)
    n, m = rom.n, rom.m
    z0 = torch.randn(N, n, device="cuda") * 0.3
    v0 = torch.randn(N, m, device="cuda")
    P = 9 * m + 6
    rec = {}
    if cls not in ("SingleInt2D", "DoubleInt2D"):
        for _ in range(3):
            rom.f(z0, v0)
        ms = timed(lambda: rom.f(z0, v0), 20)
        by = 4 * (2 * n + m) * N
        rec["f"] = dict(ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / PEAK, bytes_per_env=4 * (2 * n + m))
    for _ in range(3):
        gen.reset(z0)
    ms = timed(lambda: gen.reset(z0), 10)
    by = 4 * (n + 2 * P + (Wn + 1) * n + Wn * m) * N
    rec["reset"] = dict(ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / PEAK, bytes_per_env=by // N)
    for _ in range(10):
        gen.step()
    steps = 100     # 20 ROM knots: one call in five shifts the windows
    ms = timed(gen.step, steps)
    per = 4 * (P + n + 1 + m) + 8 * ((Wn + 1) * n + Wn * m) * dt_loop / rom_dt
    rec["step"] = dict(ms=ms, gbs=per * N / ms / 1e6, frac=per * N / ms / 1e6 / PEAK, bytes_per_env=per, env_steps_per_s=N / ms * 1e3)
    if cls in ("SingleInt2D", "DoubleInt2D"):      # the register-resident kernels of csrc/rom.cu on the same state
        ref = R.TrajectoryGenerator(rom, R.UniformSampleHoldDT(1.0, 2.0), R.UniformWeightSampler(), dt_loop=dt_loop, N=Wn, freq_low=0.01, freq_high=2.0,
                                    seed=1, device="cuda", prob_stationary=0.0005, generic_kernels=False)
        ref.reset(z0)
        for _ in range(10):
            ref.step()
        rec["step_register_kernels_ms"] = timed(ref.step, steps)
    out["classes"][cls] = rec
    del gen, rom
print(json.dumps(out))
