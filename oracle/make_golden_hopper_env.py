"""tests/golden/hopper_env_reference.npz: outputs of the UNMODIFIED HopperTrajectory env (legged_gym/envs/hopper/hopper_trajectory.py, run by
oracle/ref_harness.make_reference_hopper_trajectory) over 16 steps of the seeded case `tests/test_hopper_gpu._hopper_env_pair` builds, for two
reward tables.  Build-container only; the GPU test replays the file without the reference tree.  Usage: python -m oracle.make_golden_hopper_env"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ("yaml_table", "all_terms_spindown")
N, STEPS = 96, 16


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import ref_harness as H
    from oracle import port_hopper_env as E
    from test_hopper_cpu import HOPPER_ENV_CASES
    out = {}
    for name in CASES:
        hp = E.hopper_env_params(N, seed=3, **HOPPER_ENV_CASES[name])
        tape = E.make_hopper_tape(N, frames=8, seed=1, origins=E.grid_origins(N))
        dr = E.make_domain_rand(N, hp, seed=2)
        g = torch.Generator().manual_seed(5)
        tpush = 0.15 * torch.rand(N, generator=g)
        ep = torch.randint(0, 1002, (N,), generator=g)
        env = H.make_reference_hopper_trajectory(hp, dr, tape, tpush, episode_lengths=ep)
        env.reset_traj(torch.arange(N))                                   # reference code: generators start next to the robots
        rec = {k: [] for k in ("obs", "rew", "reset", "time_out", "root_states", "torques", "prev_error", "trajectory", "time_until_next_push")}
        for s in range(STEPS):
            o, _, r, d, _ = env.step(tape.actions[s % 8].clone())
            rec["obs"].append(o.clone()), rec["rew"].append(r.clone()), rec["reset"].append(d.bool().clone())
            rec["time_out"].append(env.time_out_buf.clone()), rec["root_states"].append(env.root_states.clone())
            rec["torques"].append(env.torques.clone()), rec["prev_error"].append(env.prev_error.clone())
            rec["trajectory"].append(env.trajectory.clone()), rec["time_until_next_push"].append(env.time_until_next_push.reshape(-1).clone())
        for k, v in rec.items():
            out[f"{name}/{k}"] = torch.stack(v).numpy()
    path = os.path.join(ROOT, "tests", "golden", "hopper_env_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
