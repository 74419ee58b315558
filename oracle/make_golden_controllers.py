"""Writes tests/golden/raibert_reference.npz by running the reference's own RaibertHeuristic (deep_tube_learning/controllers.py).
Build-container only.  Usage: python -m oracle.make_golden_controllers"""
import os

import numpy as np

from oracle import ref_harness as H
from oracle import port_controllers as PC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ref = H.import_reference()
    out = {}
    for tag, gains in (("a", PC.GAINS), ("b", PC.GAINS_B)):
        obs = PC.sample_obs(512, seed=3 if tag == "a" else 4)
        act = ref.controllers.RaibertHeuristic.raibert_policy(obs, gains["Kp"], gains["Kv"], gains["K_ff"], gains["clip_pos"],
                                                              gains["clip_vel"], gains["clip_ang"])
        out[f"obs_{tag}"], out[f"act_{tag}"] = obs.numpy(), act.numpy()
    path = os.path.join(ROOT, "tests", "golden", "raibert_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
