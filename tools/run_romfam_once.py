"""ncu driver: one reset + 12 generator steps of the 6-state / 3-input rom class at 1M envs through the generic kernels
(steps 5 and 10 cross a ROM knot: the staged window shift).  Usage: ncu ... python tools/run_romfam_once.py [num_envs] [class]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_dev_b200 import rom as R   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cls = sys.argv[2] if len(sys.argv) > 2 else "ExtendedLateralUnicycle"
zmax, vmax = {"ExtendedLateralUnicycle": ([1e9, 1e9, 1e9, 1.0, 0.5, 2.0], [1.0, 0.6, 4.0]), "Unicycle": ([1e9] * 3, [1.0, 2.0])}[cls]
rom = R.ROM_CLASSES[cls](0.1, [-a for a in zmax], zmax, [-a for a in vmax], vmax, n_robots=N, device="cuda")
gen = R.TrajectoryGenerator(rom, R.UniformSampleHoldDT(1.0, 2.0), R.UniformWeightSampler(), dt_loop=0.02, N=10, freq_low=0.01, freq_high=2.0, seed=1,
                            device="cuda", prob_stationary=0.0005)
gen.reset(torch.randn(N, rom.n, device="cuda") * 0.3)
for _ in range(12):
    gen.step()
torch.cuda.synchronize()
print("ok", float(gen.k.float().mean()))
