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


def hopper_trajectory_inputs(N=96, W=10):
    """Seeded inputs shared by the golden writer and the tests that replay it."""
    import torch
    from oracle.port_hopper import obs_case
    case = obs_case(N, seed=6)
    g = torch.Generator().manual_seed(2)
    traj = torch.randn(N, W, 2, generator=g) * 0.5 + case["root_states"][:, None, :2]
    scale = torch.tensor([2.0, 2.0])[None, :].repeat(W, 1)
    vdes = torch.randn(N, 2, generator=g) * 0.2
    state = dict(dof_state=torch.randn(N, 4, 2, generator=torch.Generator().manual_seed(1)),
                 root_states=torch.randn(N, 13, generator=torch.Generator().manual_seed(2)),
                 actions=torch.randn(N, 4, generator=torch.Generator().manual_seed(3)))
    origins = torch.randn(N, 3, generator=g)
    return case, traj, scale, vdes, state, origins, torch.arange(0, N, 3), torch.arange(1, N, 4)


def hopper_trajectory_golden():
    from oracle.port_controllers import GAINS
    from oracle.port_hopper import OBS_CFG, RESET_CFG
    case, traj, scale, vdes, state, origins, ids, push = hopper_trajectory_inputs()
    obs, nv, raib = H.reference_hopper_trajectory_observations(case, OBS_CFG, traj, scale, GAINS, vdes, seed=5, event=7)
    H.reference_hopper_trajectory_reset(state, ids, origins, RESET_CFG, seed=4, event=9, push_idx=push)
    out = dict(obs=obs.numpy(), noise_scale_vec=nv.numpy(), raibert=raib.numpy())
    out.update({f"reset_{k}": v.numpy() for k, v in state.items()})
    return out


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
    # HopperTrajectory groundwork (no kernel yet): trajectory-block observations, _reward_raibert, reset + push, from the unmodified methods
    for k, v in hopper_trajectory_golden().items():
        out[f"ht_{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "hopper_torques_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
