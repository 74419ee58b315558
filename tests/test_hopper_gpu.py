"""SURVEY 8f row 3 (part) on the GPU: b200gym_hopper_torques through the host mirror's `_compute_torques` against the CPU oracle port and the
reference-generated fixture.  Tolerance: 1e-5 relative with S = 1 N·m on torques whose orientation error stays away from pi; the so3 log
map amplifies one-ulp differences of acos / sin by 1 / sin(phi) next to pi (the reference shows the same sensitivity against itself), so the
all-angles case states its own, looser bound."""
import os

import numpy as np
import pytest
import torch

from legged_gym_dev_b200.hopper import HopperActuation
from oracle.compare import assert_close, max_err
from oracle.make_golden_hopper import CASES
from oracle.port_hopper import hopper_case, hopper_torques

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hopper_torques_reference.npz")
STATE = ("dof_state", "contact_forces", "root_states", "base_ang_vel", "p_gain_random", "d_gain_random", "torque_limit_random", "wheel_limit_random",
         "spring_stiffness", "spring_damping", "foot_pos_des", "torque_speed_bound_ratio_random")


def make_env(case, **over):
    env = HopperActuation(case["num_envs"], num_bodies=case["contact_forces"].shape[1], foot_body=case["foot_body"], device="cuda",
                          control_type=case["control_type"], action_scale=case["action_scale"], torque_speed_bound_ratio=case["torque_speed_bound_ratio"],
                          rot_actuator=case["rot_actuator"], **{k: case[k].tolist() for k in ("p_gains", "d_gains", "kd_spindown", "torque_limits",
                                                                                             "wheel_speed_limits")})
    env.load(**{k: case[k] for k in STATE})
    return env


@pytest.mark.parametrize("name", sorted(CASES))
def test_fused_matches_reference_golden(name):
    g = np.load(GOLD)
    N, seed, ct, ang, over = CASES[name]
    case, act = hopper_case(N, seed=seed, control_type=ct, max_angle=ang, **over)
    env = make_env(case)
    clipped = env._compute_torques(act.cuda())
    assert_close(clipped, g[f"{name}_clipped"], 1.0, f"{name}: returned torques")
    assert_close(env.torques, g[f"{name}_torques"], 1.0, f"{name}: self.torques")


@pytest.mark.parametrize("ct", ["orientation_spindown", "orientation"])
@pytest.mark.parametrize("N,ang,over", [(1, 1.0, {}), (257, 2.6, dict(torque_limits=[9000.0, 80.0, 80.0, 80.0])), (4099, 0.05, dict(torque_limits=[9000.0, 80.0, 80.0, 80.0])),
                                        (1000, 2.0, dict(action_scale=0.5, torque_speed_bound_ratio=2.0, wheel_speed_limits=[300.0, 600.0, 150.0],
                                                         torque_limits=[500.0, 30.0, 60.0, 90.0]))])
def test_parity_with_port(ct, N, ang, over):
    case, act = hopper_case(N, seed=N, control_type=ct, max_angle=ang, **over)
    want, want_t = hopper_torques(case, act)
    env = make_env(case)
    got = env._compute_torques(act.cuda())
    assert_close(got, want, 1.0, "returned torques")
    assert_close(env.torques, want_t, 1.0, "self.torques")
    contact = (case["contact_forces"][:, case["foot_body"], 2] > 0.1)
    if ct == "orientation_spindown" and N > 1:      # spin-down rows never touch the quaternion path: products of inputs, bit-exact
        assert torch.equal(env.torques.cpu()[contact], want_t[contact])


def test_all_angles_up_to_pi():
    case, act = hopper_case(20000, seed=8, control_type="orientation", max_angle=3.14159, torque_limits=[9000.0, 200.0, 200.0, 200.0])
    want, _ = hopper_torques(case, act)
    got = make_env(case)._compute_torques(act.cuda())
    assert bool(torch.isfinite(got).all())
    assert max_err(got, want, 1.0) < 2e-3      # within 0.01 rad of pi one ulp of sin(phi) is 1e-5 of phi / (2 sin phi)


def test_rejects_what_the_reference_cannot_run():
    for ct in ("orientation_w_foot", "V", "T"):
        with pytest.raises(NameError, match="Unknown controller type"):
            HopperActuation(8, control_type=ct)
    env = HopperActuation(8)
    with pytest.raises(RuntimeError, match="CUDA only"):
        env._compute_torques(torch.zeros(8, 4))
    with pytest.raises(ValueError, match="float32"):
        env._compute_torques(torch.zeros(7, 4, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        HopperActuation(8, device="cpu")


def test_full_size_properties_1m():
    """1 048 576 envs: contact envs get exactly the spring force on the foot and exactly the spin-down damping on the wheels; every torque
    respects its limit and the torque-speed envelope; the result does not depend on how the envs are split across launches."""
    N = 1 << 20
    case, act = hopper_case(N, seed=21, control_type="orientation_spindown", max_angle=2.0, torque_limits=[9000.0, 80.0, 80.0, 80.0])
    env = make_env(case)
    out = env._compute_torques(act.cuda()).clone()
    c = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in case.items()}
    contact = c["contact_forces"][:, case["foot_body"], 2] > 0.1
    fp, fv, wv = c["dof_state"][:, 0, 0], c["dof_state"][:, 0, 1], c["dof_state"][:, 1:, 1]
    spring = -c["spring_stiffness"][:, 0] * fp - c["spring_damping"][:, 0] * fv
    assert torch.equal(env.torques[contact, 0], spring[contact])
    tb = c["torque_limits"] * c["torque_limit_random"]
    wb = c["wheel_speed_limits"] * c["wheel_limit_random"]
    ts = case["torque_speed_bound_ratio"] * c["torque_speed_bound_ratio_random"]
    upper, lower = -ts * tb[:, 1:] / wb * (wv - wb), -ts * tb[:, 1:] / wb * (wv + wb)
    spin = torch.clip(-(c["kd_spindown"] * c["d_gain_random"][:, 1:]) * wv, lower, upper)
    assert torch.equal(env.torques[contact, 1:], spin[contact])
    assert bool((out.abs() <= tb).all()) and bool((env.torques[:, 1:] <= upper).all()) and bool((env.torques[:, 1:] >= lower).all())
    half = N // 2
    for sl in (slice(0, half), slice(half, N)):
        sub = dict(case, num_envs=half, **{k: case[k][sl] for k in STATE})
        assert torch.equal(make_env(sub)._compute_torques(act[sl].cuda()), out[sl])


# ---- observations + the Hopper's own reward terms ---------------------------------------------------------------------------------
OBS_STATE = ("root_states", "base_lin_vel", "base_ang_vel", "dof_state", "commands", "actions", "last_dof_vel", "torques")


def make_obs_env(case, cfg, seed=5, env_id_offset=0):
    env = HopperActuation(case["num_envs"], device="cuda", obs_cfg=cfg, seed=seed, env_id_offset=env_id_offset, dt=0.02)
    env.load(**{k: case[k] for k in OBS_STATE})
    env.common_step_counter = 7
    return env


def test_observations_match_reference_golden():
    from oracle.port_hopper import OBS_CFG, obs_case
    g = np.load(GOLD)
    case = obs_case(160, seed=9)
    env = make_obs_env(case, OBS_CFG)
    assert_close(env.noise_scale_vec, g["noise_scale_vec"], 1.0, "noise_scale_vec")
    assert_close(env.compute_observations(), g["obs_noise"], 1.0, "observations with noise")
    assert_close(make_obs_env(case, dict(OBS_CFG, add_noise=False, clip_observations=1.5)).compute_observations(), g["obs_plain"], 1.0, "plain observations")
    for j, name in enumerate(("_reward_torque_limits", "_reward_dof_acc", "_reward_unit_quat")):
        assert_close(getattr(env, name)(), g["reward_terms"][:, j], 1.0, name)


@pytest.mark.parametrize("N", [1, 127, 128, 1000, 4099])
def test_observations_parity_ragged(N):
    from oracle.port_hopper import OBS_CFG, hopper_observations, hopper_reward_terms, obs_case
    case = obs_case(N, seed=N)
    for cfg in (OBS_CFG, dict(OBS_CFG, add_noise=False), dict(OBS_CFG, noise_level=3.0, clip_observations=2.0)):
        env = make_obs_env(case, cfg)
        assert_close(env.compute_observations(), hopper_observations(case, cfg, seed=5, event=7), 1.0, f"observations {N}")
        assert_close(env._reward_terms(), hopper_reward_terms(case, 0.02), 1.0, f"reward terms {N}")


def test_observations_shard_invariance_1m():
    """1 048 576 envs: two half-size launches with global env ids equal the whole (noise keyed by global id); without noise the clipped
    observation is a pure function of the env's own row; the action block is a unit quaternion with qw >= 0."""
    from oracle.port_hopper import OBS_CFG, obs_case
    N = 1 << 20
    case = obs_case(N, seed=3)
    whole = make_obs_env(case, OBS_CFG).compute_observations().clone()
    half = N // 2
    for lo in (0, half):
        sub = dict(case, num_envs=half, **{k: case[k][lo:lo + half] for k in OBS_STATE})
        assert torch.equal(make_obs_env(sub, OBS_CFG, env_id_offset=lo).compute_observations(), whole[lo:lo + half])
    q = whole[:, 17:21]
    assert bool((q[:, 0] >= 0).all()) and float((q.norm(dim=1) - 1).abs().max()) < 1e-5
    plain = make_obs_env(case, dict(OBS_CFG, add_noise=False)).compute_observations()
    assert torch.equal(plain[:, 1:5], case["root_states"][:, 3:7].cuda()) and torch.equal(plain[:, 14:], whole[:, 14:])
    assert float((whole - plain)[:, :14].abs().max()) <= 0.0501 and float((whole - plain)[:, :14].abs().max()) > 0.04
