"""SURVEY 8f row 3 (part) on the GPU: b200gym_hopper_torques through the host mirror's `_compute_torques` against the CPU oracle port and the
reference-generated fixture.  Tolerance: 1e-5 relative with S = 1 N·m on torques whose orientation error stays away from pi; the so3 log
map amplifies one-ulp differences of acos / sin by 1 / sin(phi) next to pi (the reference shows the same sensitivity against itself), so the
all-angles case states its own, looser bound."""
import os

import numpy as np
import pytest
import torch

from legged_gym_dev_b200.hopper import HopperActuation
from oracle.compare import assert_close, assert_exact, max_err
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


# ---- the whole HopperTrajectory env step against oracle/port_hopper_env.HopperTrajPort, which is pinned step by step to the
# unmodified reference class (tests/test_hopper_cpu.py::test_hopper_env_port_tracks_unmodified_reference) --------------------------------------
HOPPER_ALL_SCALES = dict(termination=-500.0, tracking_rom=6.0, ang_vel_xy=-0.01, orientation=-80.0, torques=-1e-6, dof_acc=-2.5e-8, unit_quat=-0.01,
                         collision=-1.0, action_rate=-0.01, differential_error=10.0, raibert=-0.1, base_height=-1.0, dof_pos_limits=-10.0,
                         dof_vel=-1e-6, dof_vel_limits=-0.5, feet_air_time=1.0, feet_contact_forces=-0.01, lin_vel_z=-2.0, stumble=-0.3,
                         torque_limits=-0.02)
HOPPER_ENV_CASES = {
    "yaml_table": {},
    "all_terms_spindown": dict(scales=HOPPER_ALL_SCALES, control_type="orientation_spindown", penalised_bodies=[1, 2, 3], only_positive_rewards=True),
    "nonoise_nopush": dict(obs=dict(add_noise=False), push_robots=False, reset=dict(randomize_yaw=False)),
}


def _hopper_env_pair(name, N, seed=3, env_id_offset=0, lo=0, hi=None):
    from oracle import port_hopper_env as E
    from legged_gym_dev_b200 import configs
    from legged_gym_dev_b200.hopper_trajectory import HopperTrajectory, HopperReplayPhysics
    from types import SimpleNamespace
    hi = N if hi is None else hi
    hp_all = E.hopper_env_params(N, seed=seed, **HOPPER_ENV_CASES[name])
    tape = E.make_hopper_tape(N, frames=8, seed=1, origins=E.grid_origins(N))
    dr = E.make_domain_rand(N, hp_all, seed=2)
    g = torch.Generator().manual_seed(5)
    tpush = 0.15 * torch.rand(N, generator=g)
    ep = torch.randint(0, 1002, (N,), generator=g)
    n = hi - lo
    hp = E.hopper_env_params(n, seed=seed, **HOPPER_ENV_CASES[name])
    cut = SimpleNamespace(dof=tape.dof[:, :, lo:hi].contiguous(), root=tape.root[:, :, lo:hi].contiguous(), contact=tape.contact[:, :, lo:hi].contiguous(),
                          actions=tape.actions[:, lo:hi].contiguous(), num_envs=n, frames=tape.frames, decimation=tape.decimation)
    dr_cut = {k: v[lo:hi].contiguous() for k, v in dr.items()}
    cfg = E.apply_params_to_cfg(configs.hopper_flat_trajectory_cfg(), hp)
    asset = dict(num_bodies=hp.num_bodies, foot_body=hp.foot_body, termination_bodies=hp.termination_bodies, penalised_bodies=hp.penalised_bodies,
                 torque_limits=hp.torque_limits, dof_pos_limits=hp.dof_pos_limits, dof_vel_limits=hp.dof_vel_limits)
    env = HopperTrajectory(cfg, SimpleNamespace(dt=hp.sim_dt), None, "cuda", True, physics=HopperReplayPhysics(cut, device="cuda"), asset=asset,
                           domain_rand_values=dr_cut, seed=seed, env_id_offset=env_id_offset + lo)
    env.time_until_next_push.copy_(tpush[lo:hi].reshape(-1, 1))
    env.episode_length_buf.copy_(ep[lo:hi])
    port = E.HopperTrajPort(hp, dr_cut, cut, tpush[lo:hi], episode_length_buf=ep[lo:hi], env_origins=env.env_origins.cpu(), env_id_offset=env_id_offset + lo)
    return hp, cut, env, port, E.HopperTapePhysics(cut)


def _compare_hopper_env(env, port, tag):
    N = port.N
    assert_exact(env.reset_buf.cpu(), port.reset_buf, tag + "reset")
    assert_exact(env.time_out_buf.cpu(), port.time_out_buf, tag + "time_out")
    assert_exact(env.episode_length_buf.cpu(), port.episode_length_buf, tag + "ep_len")
    assert_exact(env.last_contacts.cpu(), port.last_contacts, tag + "last_contacts")
    assert_close(env.obs_buf.cpu(), port.obs_buf, 1.0, tag + "obs")
    assert_close(env.rew_buf.cpu(), port.rew_buf, 1.0, tag + "rew")
    assert_close(env.torques.cpu(), port.torques, 300.0, tag + "torques")
    for k in ("root_states", "trajectory", "prev_error", "last_actions", "last_dof_vel", "last_root_vel", "base_ang_vel", "base_lin_vel",
              "projected_gravity", "feet_air_time", "actions"):
        assert_close(getattr(env, k).cpu(), getattr(port, k), 1.0, tag + k)
    assert_close(env.time_until_next_push.cpu().reshape(-1), port.time_until_next_push, 1.0, tag + "time_until_next_push")
    assert_close(env.dof_state.cpu().view(N, 4, 2), port.dof_state, 1.0, tag + "dof_state")
    assert list(env.episode_sums) == list(port.episode_sums)
    for k in port.episode_sums:
        assert_close(env.episode_sums[k].cpu(), port.episode_sums[k], 1.0, tag + "sum_" + k)
    for k, v in port.extras.get("episode", {}).items():
        assert_close(env.extras["episode"][k].cpu(), v, 1.0, tag + "extras " + k)


@pytest.mark.parametrize("name,N", [("yaml_table", 1000), ("all_terms_spindown", 516), ("nonoise_nopush", 67), ("yaml_table", 3)])
def test_hopper_env_step_parity(name, N):
    """HopperTrajectory.step on the fused path (4 x torque law + generator step + prologue / post-physics / finaliser + generator reset) against
    the port, every step: flags and counters exact, everything else to 1e-5 (S = 1; 300 for torques = the foot's torque limit)."""
    hp, tape, env, port, phys = _hopper_env_pair(name, N)
    ids = torch.arange(N)
    port.gen.reset_traj(ids, port.proj_z())          # the generators start next to the robots on both sides (a never-reset generator evaluates 0/0)
    env.reset_traj_all()
    resets = 0
    for s in range(24):
        a = tape.actions[s % 8] * (300.0 if s == 3 else 1.0)
        port.step(a.clone(), phys)
        env.step(a.cuda())
        resets += int(port.reset_buf.sum())
        _compare_hopper_env(env, port, f"{name} N={N} step {s}: ")
    assert resets > 0


def test_hopper_env_shard_invariance():
    """Per-env results do not depend on the split over ranks, except for the reference's env-0 push quirk (local env 0 of every process rides
    along with every push of its process): compared on a case without pushes."""
    N, half = 128, 64
    _, tape, env, _, _ = _hopper_env_pair("nonoise_nopush", N)
    _, tape_b, env_b, _, _ = _hopper_env_pair("nonoise_nopush", N, lo=half, hi=N)
    env_b.env_origins.copy_(env.env_origins[half:])   # the shard keeps the origins of its global env ids (a per-process input)
    env.reset_traj_all()
    env_b.reset_traj_all()
    for s in range(10):
        env.step(tape.actions[s % 8].cuda())
        env_b.step(tape_b.actions[s % 8].cuda())
        assert torch.equal(env.obs_buf[half:], env_b.obs_buf) and torch.equal(env.rew_buf[half:], env_b.rew_buf)
        assert torch.equal(env.reset_buf[half:], env_b.reset_buf) and torch.equal(env.root_states[half:], env_b.root_states)


@pytest.mark.parametrize("name", ["yaml_table", "all_terms_spindown"])
def test_hopper_env_replays_reference_golden(name):
    """The fused Hopper env against outputs of the UNMODIFIED reference class itself (tests/golden/hopper_env_reference.npz, written by
    oracle/make_golden_hopper_env.py in the build container): observations, rewards, flags, root state, torques, prev_error, push timers."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hopper_env_reference.npz")
    gold = np.load(path)
    N = gold[f"{name}/obs"].shape[1]
    hp, tape, env, port, phys = _hopper_env_pair(name, N)
    env.reset_traj_all()
    for s in range(gold[f"{name}/obs"].shape[0]):
        env.step(tape.actions[s % 8].cuda())
        tag = f"golden {name} step {s}: "
        g = lambda k: torch.from_numpy(gold[f"{name}/{k}"][s])
        assert_exact(env.reset_buf.cpu(), g("reset"), tag + "reset")
        assert_exact(env.time_out_buf.cpu(), g("time_out"), tag + "time_out")
        assert_close(env.obs_buf.cpu(), g("obs"), 1.0, tag + "obs")
        assert_close(env.rew_buf.cpu(), g("rew"), 1.0, tag + "rew")
        assert_close(env.torques.cpu(), g("torques"), 300.0, tag + "torques")
        assert_close(env.root_states.cpu(), g("root_states"), 1.0, tag + "root_states")
        assert_close(env.prev_error.cpu(), g("prev_error"), 1.0, tag + "prev_error")
        assert_close(env.trajectory.cpu(), g("trajectory"), 1.0, tag + "trajectory")
        assert_close(env.time_until_next_push.cpu().reshape(-1), g("time_until_next_push"), 1.0, tag + "time_until_next_push")


@pytest.mark.parametrize("name,N", [("yaml_table", 300), ("all_terms_spindown", 67)])
def test_hopper_env_external_reset_matches_oracle(name, N):
    """HopperTrajectory.reset() / reset_idx(env_ids) from outside step() (hopper_trajectory.py:286-296, legged_robot_trajectory.py:204-246)
    against the port, whose reset() is pinned to the reference's (tests/test_hopper_cpu.py): masked reset launch + generator reset, immediate."""
    hp, tape, env, port, phys = _hopper_env_pair(name, N)
    port.reset(phys)
    obs, _ = env.reset()
    assert torch.isfinite(obs).all()
    _compare_hopper_env(env, port, f"{name} reset(): ")
    assert not bool(env.time_out_buf.any())
    for s in range(3):
        a = tape.actions[s % 8]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        _compare_hopper_env(env, port, f"{name} step {s} after reset(): ")
    ids = torch.arange(1, N, 3)
    port.reset_idx(ids)
    env.reset_idx(ids.cuda())
    _compare_hopper_env(env, port, f"{name} partial reset: ")
    assert bool(env.reset_buf[ids.cuda()].all())
    for s in range(3, 8):
        a = tape.actions[s % 8]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        _compare_hopper_env(env, port, f"{name} step {s} after the partial reset: ")
