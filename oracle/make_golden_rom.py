"""tests/golden/rom_*.npz from the UNMODIFIED reference CustomSim / TrajectoryGenerator / DoubleSingleTracking
(oracle/ref_harness.py).  Build-container only.  Run through `python -m oracle.make_golden`."""
import os

import numpy as np
import torch

from oracle import ref_harness as H

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {"default": {}, "fast_resample": dict(prob_stationary=0.05, t_low=0.2, t_high=0.5)}


def rom_golden(name, over, N=48, steps=120, seed=3):
    env, policy, cfg = H.make_reference_custom_sim(N, seed=seed, **over)
    out = dict(num_envs=np.int64(N), steps=np.int64(steps), seed=np.int64(seed), ramp_v_end0=env.traj_gen.ramp_v_end.numpy().copy())
    env.reset()
    obs = env.get_observations()
    keep = (0, 1, 2, 5, 19, 20, 59, 60, 61, 100, steps - 1)
    for s in range(steps):
        if s == 60:
            env.reset_idx(torch.arange(0, N, 3))
            obs = env.get_observations()
        a = policy(obs)
        obs, _, _, d, _ = env.step(a)
        if s in keep:
            tg = env.traj_gen
            snap = dict(action=a, obs=obs, root=env.root_states, traj=tg.trajectory, vtraj=tg.v_trajectory, v=tg.v, t=tg.t, k=tg.k,
                        t_final=tg.t_final, weights=tg.weights, stationary=tg.stationary_inds, env_trajectory=env.trajectory)
            for k, v in snap.items():
                out[f"s{s}_{k}"] = v.detach().numpy().copy()
    out["keep"] = np.array(keep)
    out["ctr"] = env._shim.ctr.copy()
    path = os.path.join(GOLD, f"rom_{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    for name, over in CASES.items():
        rom_golden(name, over)


if __name__ == "__main__":
    main()
