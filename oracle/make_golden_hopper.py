"""tests/golden/hopper_torques_reference.npz: outputs of the UNMODIFIED Hopper._compute_torques (legged_gym/envs/hopper/hopper.py:168-237,
run by oracle/ref_harness.reference_hopper_torques on pytorch3d.transforms = oracle/pytorch3d_restated) for the seeded cases of
oracle/port_hopper.hopper_case.  Build-container only.  Usage: python -m oracle.make_golden_hopper"""
import os

import numpy as np

from oracle import ref_harness as H
from oracle.port_hopper import hopper_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# name -> (num_envs, seed, control_type, max error angle, overrides); the wide limits keep most torques inside the clips
CASES = {"shipped": (192, 3, "orientation_spindown", 2.6, {}),
         "wide_limits": (192, 4, "orientation_spindown", 2.6, dict(torque_limits=[9000.0, 80.0, 80.0, 80.0])),
         "no_spindown": (160, 5, "orientation", 1.0, dict(torque_limits=[9000.0, 40.0, 40.0, 40.0], p_gains=[900.0, 15.0, 15.0, 12.0], d_gains=[60.0, 3.0, 3.0, 2.0])),
         "small_angles": (128, 6, "orientation_spindown", 0.02, dict(torque_limits=[9000.0, 80.0, 80.0, 80.0]))}


def main():
    out = {}
    for name, (N, seed, ct, ang, over) in CASES.items():
        case, act = hopper_case(N, seed=seed, control_type=ct, max_angle=ang, **over)
        clipped, torques = H.reference_hopper_torques(case, act)
        out[f"{name}_clipped"], out[f"{name}_torques"] = clipped.numpy(), torques.numpy()
    # observations (with / without noise) and the Hopper's own reward terms, from the unmodified methods (hopper.py:239-258,407-430,448-458)
    from oracle.port_hopper import OBS_CFG, obs_case
    case = obs_case(160, seed=9)
    for tag, cfg in (("noise", OBS_CFG), ("plain", dict(OBS_CFG, add_noise=False, clip_observations=1.5))):
        obs, nv, terms = H.reference_hopper_observations(case, cfg, seed=5, event=7)
        out[f"obs_{tag}"], out["noise_scale_vec"], out["reward_terms"] = obs.numpy(), nv.numpy(), terms.numpy()
    path = os.path.join(ROOT, "tests", "golden", "hopper_torques_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
