"""SURVEY 8f row 3 (part), CPU side: the restated pytorch3d functions against scipy, the oracle port of Hopper._compute_torques against the
UNMODIFIED reference method (build container) and against the reference-generated fixture (everywhere)."""
import os

import numpy as np
import pytest
import torch

from oracle import pytorch3d_restated as P3
from oracle.compare import assert_close, assert_exact
from oracle.make_golden_hopper import CASES
from oracle.port_hopper import hopper_case, hopper_torques

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hopper_torques_reference.npz")


def test_restated_pytorch3d_against_scipy():
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(500, 4, generator=g, dtype=torch.float64), torch.randn(500, 4, generator=g, dtype=torch.float64)
    a, b = a / a.norm(dim=1, keepdim=True), b / b.norm(dim=1, keepdim=True)
    to_scipy = lambda q: Rotation.from_quat(q[:, [1, 2, 3, 0]].numpy())          # scipy is scalar-last
    prod = P3.quaternion_multiply(P3.quaternion_invert(a), b)
    want = to_scipy(a).inv() * to_scipy(b)
    assert bool((prod[:, 0] >= 0).all())
    R = P3.quaternion_to_matrix(prod)
    assert np.allclose(R.numpy(), want.as_matrix(), atol=1e-12)
    keep = torch.from_numpy(want.magnitude() < 3.0)                               # away from pi the log map is well conditioned
    assert np.allclose(P3.so3_log_map(R)[keep].numpy(), want.as_rotvec()[keep.numpy()], atol=1e-7)
    M = Rotation.from_euler("xyz", [0.3, -0.2, 1.1]).as_matrix()
    pts = torch.randn(50, 3, generator=g, dtype=torch.float64)
    assert np.allclose(P3.Rotate(torch.tensor(M), dtype=torch.float64).transform_points(pts).numpy(), pts.numpy() @ M, atol=1e-12)
    # the linear continuation of acos: continuous and first-order exact at the bounds
    x = torch.tensor([1 - 1e-4, 1 - 0.5e-4, 1.0, -1 + 1e-4, -1.0], dtype=torch.float64)
    y = P3.acos_linear_extrapolation(x, (-1 + 1e-4, 1 - 1e-4))
    assert abs(float(y[0]) - np.arccos(1 - 1e-4)) < 1e-12 and abs(float(y[3]) - np.arccos(-1 + 1e-4)) < 1e-12
    assert float(y[1]) < float(y[0]) and float(y[2]) < float(y[1]) and float(y[4]) > float(y[3])


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference_golden(name):
    g = np.load(GOLD)
    N, seed, ct, ang, over = CASES[name]
    case, act = hopper_case(N, seed=seed, control_type=ct, max_angle=ang, **over)
    clipped, torques = hopper_torques(case, act)
    assert_close(clipped, g[f"{name}_clipped"], 1.0, f"{name}: clipped torques")
    assert_close(torques, g[f"{name}_torques"], 1.0, f"{name}: self.torques")


def test_cases_exercise_every_branch():
    """The fixtures are only worth something if the clips are not always active and both contact states occur."""
    for name, (N, seed, ct, ang, over) in CASES.items():
        case, act = hopper_case(N, seed=seed, control_type=ct, max_angle=ang, **over)
        clipped, torques = hopper_torques(case, act)
        contact = case["contact_forces"][:, case["foot_body"], 2] > 0.1
        assert 0.25 < float(contact.float().mean()) < 0.75
        tb = case["torque_limits"] * case["torque_limit_random"]
        inside = (clipped.abs() < tb)[:, 1:]
        if name != "shipped":
            assert float(inside.float().mean()) > 0.3, f"{name}: wheel torques saturate everywhere"
        assert bool(torch.isfinite(clipped).all())


@pytest.mark.reference
@pytest.mark.parametrize("ct", ["orientation_spindown", "orientation"])
@pytest.mark.parametrize("seed,N,ang", [(1, 257, 3.1), (2, 64, 0.01), (3, 500, 1.5)])
def test_port_equals_unmodified_reference_method(ct, seed, N, ang):
    from oracle import ref_harness as H
    case, act = hopper_case(N, seed=seed, control_type=ct, max_angle=ang, torque_limits=[9000.0, 80.0, 80.0, 80.0])
    want, want_t = H.reference_hopper_torques(case, act)
    got, got_t = hopper_torques(case, act)
    assert_exact(got, want, "returned torques")
    assert_exact(got_t, want_t, "self.torques")


@pytest.mark.reference
@pytest.mark.parametrize("ct", ["orientation_spindown", "orientation"])
def test_port_equals_hopper_trajectory_torque_law_too(ct):
    """HopperTrajectory carries its own copy of the torque law (hopper_trajectory.py:184-253, restructured index handling, same arithmetic):
    the port — and so the kernel — serves both classes."""
    from oracle import ref_harness as H
    for seed, N, ang in ((1, 257, 3.1), (4, 96, 0.3)):
        case, act = hopper_case(N, seed=seed, control_type=ct, max_angle=ang, torque_limits=[9000.0, 80.0, 80.0, 80.0])
        want, want_t = H.reference_hopper_torques(case, act, cls="HopperTrajectory")
        got, got_t = hopper_torques(case, act)
        assert_exact(got, want, "returned torques")
        assert_exact(got_t, want_t, "self.torques")


@pytest.mark.reference
@pytest.mark.parametrize("ct", ["orientation_w_foot", "V", "T_spindown"])
def test_reference_cannot_run_the_other_control_types(ct):
    """Why the port / kernel reject them: the reference method itself raises (shape errors at hopper.py:196 / :224 / :227)."""
    from oracle import ref_harness as H
    case, act = hopper_case(16, seed=1, control_type=ct)
    with pytest.raises((RuntimeError, IndexError)):
        H.reference_hopper_torques(case, act)
    with pytest.raises(ValueError):
        hopper_torques(case, act)


# ---- observations + the Hopper's own reward terms ---------------------------------------------------------------------------------
def test_observation_port_matches_reference_golden():
    from oracle.port_hopper import OBS_CFG, hopper_observations, hopper_reward_terms, noise_scale_vec, obs_case
    g = np.load(GOLD)
    case = obs_case(160, seed=9)
    assert_exact(noise_scale_vec(), g["noise_scale_vec"], "noise_scale_vec")
    assert_close(hopper_observations(case, OBS_CFG, seed=5, event=7), g["obs_noise"], 1.0, "observations with noise")
    plain = hopper_observations(case, dict(OBS_CFG, add_noise=False, clip_observations=1.5), seed=5, event=7)
    assert_close(plain, g["obs_plain"], 1.0, "observations, no noise, clipped at 1.5")
    assert float(plain.abs().max()) == 1.5 and bool((plain[:, 17] >= 0).all())       # the clip bites; qw >= 0 convention
    assert_close(hopper_reward_terms(case, 0.02), g["reward_terms"], 1.0, "reward terms")


@pytest.mark.reference
@pytest.mark.parametrize("N,seed", [(200, 1), (33, 2)])
def test_observation_port_equals_unmodified_reference_methods(N, seed):
    from oracle import ref_harness as H
    from oracle.port_hopper import OBS_CFG, hopper_observations, hopper_reward_terms, noise_scale_vec, obs_case
    case = obs_case(N, seed)
    for cfg in (OBS_CFG, dict(OBS_CFG, add_noise=False), dict(OBS_CFG, noise_level=3.0, clip_observations=2.0)):
        want, nv, terms = H.reference_hopper_observations(case, cfg, seed=5, event=7)
        assert_exact(noise_scale_vec(cfg), nv, "noise_scale_vec")
        assert_exact(hopper_observations(case, cfg, seed=5, event=7), want, "observations")
        assert_exact(hopper_reward_terms(case, 0.02), terms, "reward terms")


@pytest.mark.reference
def test_reference_hopper_reset_path_is_not_runnable_as_shipped():
    """Why the Hopper slice stops at torques / observations / reward terms: `Hopper._reset_dofs`, `_reset_root_states` and `_push_robots`
    (hopper.py:270-341) use torch_rand_vec_float, matrix_to_quaternion, euler_angles_to_matrix and `push_idx`, none of which hopper.py
    imports or defines — they raise NameError on the first reset.  The class that can run is HopperTrajectory (hopper_trajectory.py imports
    them, :38-40); the remaining Hopper rows have to be pinned on that file."""
    from types import SimpleNamespace
    from oracle import ref_harness as H
    hop = H.import_reference().hopper
    for name in ("torch_rand_vec_float", "matrix_to_quaternion", "euler_angles_to_matrix"):
        assert not hasattr(hop, name), f"hopper.py now defines {name}: the reset path may have become runnable"
    ids = torch.arange(4)
    stub = SimpleNamespace(dof_pos=torch.zeros(4, 4), dof_vel=torch.zeros(4, 4), default_dof_pos=torch.zeros(1, 4), num_dof=4, device="cpu",
                           default_dof_pos_noise_lower=torch.zeros(4), default_dof_pos_noise_upper=torch.ones(4),
                           default_dof_vel_noise_lower=torch.zeros(4), default_dof_vel_noise_upper=torch.ones(4))
    with pytest.raises(NameError, match="torch_rand_vec_float"):
        hop.Hopper._reset_dofs(stub, ids)
    with pytest.raises(NameError, match="not defined"):      # torch_rand_vec_float again, then the undefined push_idx / env_ids_int32 (:336-341)
        hop.Hopper._push_robots(SimpleNamespace(max_vel=torch.ones(6), device="cpu"), ids)
    import inspect
    src = inspect.getsource(hop.Hopper._push_robots)
    assert "push_idx" in src and "push_inds" in src and "env_ids_int32" in src


@pytest.mark.reference
@pytest.mark.parametrize("add_noise", [True, False])
def test_hopper_trajectory_observation_and_raibert_ports_equal_reference(add_noise):
    """Oracle groundwork for the remaining Hopper rows (no kernel yet): HopperTrajectory.compute_observations with its trajectory block,
    its noise vector and _reward_raibert, against the unmodified methods."""
    from oracle import ref_harness as H
    from oracle.port_controllers import GAINS
    from oracle.port_hopper import OBS_CFG, hopper_reward_raibert, hopper_traj_noise_scale_vec, hopper_traj_observations, obs_case
    N, W = 120, 10
    case = obs_case(N, seed=6)
    g = torch.Generator().manual_seed(2)
    traj = torch.randn(N, W, 2, generator=g) * 0.5 + case["root_states"][:, None, :2]
    scale = torch.tensor([2.0, 2.0])[None, :].repeat(W, 1)
    vdes = torch.randn(N, 2, generator=g) * 0.2
    cfg = dict(OBS_CFG, add_noise=add_noise)
    want, nv, raib = H.reference_hopper_trajectory_observations(case, cfg, traj, scale, GAINS, vdes, seed=5, event=7)
    assert_exact(hopper_traj_noise_scale_vec(W * 2, cfg), nv, "noise_scale_vec")
    assert_exact(hopper_traj_observations(case, traj, scale, cfg, seed=5, event=7), want, "observations")
    assert_exact(hopper_reward_raibert(case, traj[:, 0], vdes, GAINS), raib, "_reward_raibert")


def test_restated_euler_and_matrix_to_quaternion_against_scipy():
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(3)
    eul = (torch.rand(400, 3, generator=g, dtype=torch.float64) - 0.5) * 6.0
    R = P3.euler_angles_to_matrix(eul, "XYZ")
    assert np.allclose(R.numpy(), Rotation.from_euler("XYZ", eul.numpy()).as_matrix(), atol=1e-12)     # intrinsic X-Y-Z
    q = P3.matrix_to_quaternion(R)
    want = Rotation.from_matrix(R.numpy()).as_quat()[:, [3, 0, 1, 2]]
    want = np.where(want[:, :1] < 0, -want, want)
    assert bool((q[:, 0] >= 0).all()) and np.allclose(q.numpy(), want, atol=1e-9)
    yaw = torch.linspace(-3.1, 3.1, 50, dtype=torch.float64)          # the only use in the reference: pure yaw (hopper_trajectory.py:344)
    qy = P3.matrix_to_quaternion(P3.euler_angles_to_matrix(torch.stack((torch.zeros_like(yaw), torch.zeros_like(yaw), yaw), -1), "XYZ"))
    y = yaw.numpy()
    assert np.allclose(qy.numpy(), np.stack((np.cos(y / 2), 0 * y, 0 * y, np.sin(y / 2)), -1), atol=1e-12)


@pytest.mark.reference
@pytest.mark.parametrize("randomize_yaw", [True, False])
def test_hopper_trajectory_reset_and_push_ports_equal_reference(randomize_yaw):
    """Oracle groundwork (no kernel yet): HopperTrajectory._reset_dofs / _reset_root_states / _push_robots against the unmodified methods."""
    from oracle import ref_harness as H
    from oracle.port_hopper import RESET_CFG, hopper_traj_push, hopper_traj_reset
    N = 90
    g = torch.Generator().manual_seed(8)
    mk = lambda: dict(dof_state=torch.randn(N, 4, 2, generator=torch.Generator().manual_seed(1)),
                      root_states=torch.randn(N, 13, generator=torch.Generator().manual_seed(2)),
                      actions=torch.randn(N, 4, generator=torch.Generator().manual_seed(3)))
    origins = torch.randn(N, 3, generator=g)
    ids, push = torch.arange(0, N, 3), torch.arange(1, N, 4)
    cfg = dict(RESET_CFG, randomize_yaw=randomize_yaw)
    a, b = mk(), mk()
    H.reference_hopper_trajectory_reset(a, ids, origins, cfg, seed=4, event=9, push_idx=push)
    hopper_traj_reset(b, ids.numpy(), origins, cfg, seed=4, event=9)
    hopper_traj_push(b, push.numpy(), cfg, seed=4, event=9)
    for k in a:
        assert_exact(b[k], a[k], k)
    q = a["root_states"][ids, 3:7]
    assert float((q.norm(dim=1) - 1).abs().max()) < 1e-5


def test_hopper_trajectory_groundwork_ports_match_reference_golden():
    """The same ports against outputs the unmodified HopperTrajectory methods wrote into the fixture (runs without /root/reference)."""
    from oracle.make_golden_hopper import hopper_trajectory_inputs
    from oracle.port_controllers import GAINS
    from oracle.port_hopper import (OBS_CFG, RESET_CFG, hopper_reward_raibert, hopper_traj_noise_scale_vec, hopper_traj_observations, hopper_traj_push,
                                    hopper_traj_reset)
    g = np.load(GOLD)
    case, traj, scale, vdes, state, origins, ids, push = hopper_trajectory_inputs()
    assert_exact(hopper_traj_noise_scale_vec(20), g["ht_noise_scale_vec"], "noise_scale_vec")
    assert_close(hopper_traj_observations(case, traj, scale, OBS_CFG, seed=5, event=7), g["ht_obs"], 1.0, "observations")
    assert_close(hopper_reward_raibert(case, traj[:, 0], vdes, GAINS), g["ht_raibert"], 1.0, "_reward_raibert")
    hopper_traj_reset(state, ids.numpy(), origins, RESET_CFG, seed=4, event=9)
    hopper_traj_push(state, push.numpy(), RESET_CFG, seed=4, event=9)
    for k, v in state.items():
        assert_close(v, g[f"ht_reset_{k}"], 1.0, f"reset {k}")


HOPPER_ALL_SCALES = dict(termination=-500.0, tracking_rom=6.0, ang_vel_xy=-0.01, orientation=-80.0, torques=-1e-6, dof_acc=-2.5e-8, unit_quat=-0.01,
                         collision=-1.0, action_rate=-0.01, differential_error=10.0, raibert=-0.1, base_height=-1.0, dof_pos_limits=-10.0,
                         dof_vel=-1e-6, dof_vel_limits=-0.5, feet_air_time=1.0, feet_contact_forces=-0.01, lin_vel_z=-2.0, stumble=-0.3,
                         torque_limits=-0.02)
HOPPER_ENV_CASES = {
    "yaml_table": {},                                                                   # hopper_single_int.yaml's reward table
    "all_terms_spindown": dict(scales=HOPPER_ALL_SCALES, control_type="orientation_spindown", penalised_bodies=[1, 2, 3], only_positive_rewards=True),
    "nonoise_nopush": dict(obs=dict(add_noise=False), push_robots=False, reset=dict(randomize_yaw=False)),
}


def build_hopper_env_case(name, N, seed=3):
    from oracle import port_hopper_env as E
    hp = E.hopper_env_params(N, seed=seed, **HOPPER_ENV_CASES[name])
    tape = E.make_hopper_tape(N, frames=8, seed=1, origins=E.grid_origins(N))
    dr = E.make_domain_rand(N, hp, seed=2)
    g = torch.Generator().manual_seed(5)
    tpush = 0.15 * torch.rand(N, generator=g)
    ep = torch.randint(0, 1002, (N,), generator=g)
    return hp, tape, dr, tpush, ep


@pytest.mark.reference
@pytest.mark.parametrize("name", list(HOPPER_ENV_CASES))
def test_hopper_env_port_tracks_unmodified_reference(name):
    """The whole HopperTrajectory env (hopper_trajectory.py: reset, step, post_physics_step with pushes, resets, rewards, observations)
    executed UNMODIFIED through oracle/ref_harness.make_reference_hopper_trajectory against oracle/port_hopper_env.HopperTrajPort, every
    step, same Philox streams."""
    from oracle import ref_harness as H
    from oracle import port_hopper_env as E
    N = 64
    hp, tape, dr, tpush, ep = build_hopper_env_case(name, N)
    env = H.make_reference_hopper_trajectory(hp, dr, tape, tpush, episode_lengths=ep)
    port = E.HopperTrajPort(hp, dr, tape, tpush, episode_length_buf=ep, env_origins=env.env_origins)
    phys = E.HopperTapePhysics(tape)
    assert [n for n in port.active if n != "termination"] == env.reward_names
    o1, _ = env.reset()
    o2, _ = port.reset(phys)
    assert_close(o2, o1, 1.0, "reset() obs")
    resets = pushes = 0
    for s in range(24):
        a = tape.actions[s % 8] * (300.0 if s == 3 else 1.0)
        before = env.time_until_next_push.clone()
        o1, _, r1, d1, x1 = env.step(a.clone())
        o2, _, r2, d2, x2 = port.step(a.clone(), phys)
        pushes += int((env.time_until_next_push.reshape(-1) > before.reshape(-1)).sum())
        resets += int(d1.sum())
        tag = f"{name} step {s}: "
        assert_exact(d2, d1.bool(), tag + "reset")
        assert_exact(port.time_out_buf, env.time_out_buf, tag + "time_out")
        assert_exact(port.episode_length_buf, env.episode_length_buf, tag + "ep_len")
        assert_exact(port.last_contacts, env.last_contacts, tag + "last_contacts")
        assert_close(o2, o1, 1.0, tag + "obs")
        assert_close(r2, r1, 1.0, tag + "rew")
        assert_close(port.torques, env.torques, 300.0, tag + "torques")
        for k in ("root_states", "trajectory", "prev_error", "last_actions", "last_dof_vel", "last_root_vel", "base_ang_vel", "base_lin_vel",
                  "feet_air_time", "actions"):
            assert_close(getattr(port, k), getattr(env, k), 1.0, tag + k)
        assert_close(port.time_until_next_push, env.time_until_next_push.reshape(-1), 1.0, tag + "time_until_next_push")
        assert_close(port.dof_state, env.dof_state.view(N, 4, 2), 1.0, tag + "dof_state")
        for k in env.episode_sums:
            assert_close(port.episode_sums[k], env.episode_sums[k], 1.0, tag + "sum_" + k)
        if "episode" in x1:
            assert list(x1["episode"]) == list(x2["episode"])
            for k in x1["episode"]:
                assert_close(x2["episode"][k], x1["episode"][k], 1.0, tag + "extras " + k)
    assert resets > 0 and (pushes > 0 or not hp.push_robots)


@pytest.mark.parametrize("name", ["yaml_table", "all_terms_spindown"])
def test_hopper_env_port_replays_reference_golden(name):
    """The travelling Hopper env port against tests/golden/hopper_env_reference.npz (outputs of the unmodified reference class, written by
    oracle/make_golden_hopper_env.py): runs without the reference tree, as on the GPU box."""
    from oracle import port_hopper_env as E
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hopper_env_reference.npz"))
    N = gold[f"{name}/obs"].shape[1]
    hp, tape, dr, tpush, ep = build_hopper_env_case(name, N)
    port = E.HopperTrajPort(hp, dr, tape, tpush, episode_length_buf=ep, env_origins=E.grid_origins(N))
    phys = E.HopperTapePhysics(tape)
    port.gen.reset_traj(torch.arange(N), port.proj_z())
    for s in range(gold[f"{name}/obs"].shape[0]):
        port.step(tape.actions[s % 8].clone(), phys)
        g = lambda k: torch.from_numpy(gold[f"{name}/{k}"][s])
        tag = f"golden {name} step {s}: "
        assert_exact(port.reset_buf, g("reset"), tag + "reset")
        assert_close(port.obs_buf, g("obs"), 1.0, tag + "obs")
        assert_close(port.rew_buf, g("rew"), 1.0, tag + "rew")
        assert_close(port.root_states, g("root_states"), 1.0, tag + "root_states")
        assert_close(port.torques, g("torques"), 300.0, tag + "torques")
