"""BASELINE.json's full sizes (1 048 576 envs) through size-independent properties — the oracle cannot run these sizes in
seconds, so the checks are identities that must hold for any N: shard invariance (results keyed by global env id), flag
identities recomputed with plain torch ops from the same inputs, clip bounds, determinism of the persistent ROM rollout, and
the sliding-window kernel against a torch gather of the same definition."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

N_FULL = 1 << 20


def _env(num_envs, tape, ep, offset=0):
    from legged_gym_dev_b200 import configs
    from legged_gym_dev_b200.legged_robot import Anymal
    from legged_gym_dev_b200.physics import ReplayPhysics
    import legged_case as LC
    cfg = configs.with_upstream_rewards(configs.anymal_c_flat_cfg(), pd_control=True)
    cfg.env.num_envs = num_envs
    env = Anymal(cfg, SimpleNamespace(dt=cfg.sim.dt), None, "cuda", True, physics=ReplayPhysics(tape, device="cuda", copy=True),
                 asset=LC.dof_limits(), seed=0, env_id_offset=offset)
    env.episode_length_buf.copy_(ep)
    return env


def _half(tape, lo, hi):
    N = tape.num_envs
    t = SimpleNamespace(num_envs=hi - lo, frames=tape.frames, decimation=tape.decimation)
    t.root, t.contact, t.actions = tape.root[:, lo:hi].contiguous(), tape.contact[:, lo:hi].contiguous(), tape.actions[:, lo:hi].contiguous()
    d = tape.dof.view(tape.frames, tape.decimation, N, 12, 2)[:, :, lo:hi]
    t.dof = d.reshape(tape.frames, tape.decimation, (hi - lo) * 12, 2).contiguous()
    return t


def test_flat_step_at_1m_envs_shard_invariance_and_flag_identities():
    from legged_gym_dev_b200 import synthetic as S
    N, F = N_FULL, 2
    tape = S.make_state_tape(N, frames=F, seed=11, device="cuda")
    ep = S.make_episode_lengths(N, seed=3, device="cuda")
    full = _env(N, tape, ep)
    lo = N // 2
    part = _env(N - lo, _half(tape, lo, N), ep[lo:], offset=lo)
    for s in range(3):
        ep_before = full.episode_length_buf.clone()
        a = tape.actions[s % F]
        obs, _, rew, rst, extras = full.step(a)
        obs2, _, rew2, rst2, _ = part.step(a[lo:])
        # shard invariance: the upper half computed on its own (global ids lo..N-1) is bit-identical
        assert torch.equal(obs[lo:], obs2) and torch.equal(rew[lo:], rew2) and torch.equal(rst[lo:], rst2)
        assert torch.equal(full.commands[lo:], part.commands) and torch.equal(full.episode_length_buf[lo:], part.episode_length_buf)
        # flag identities (legged_robot.py:139-145) recomputed with torch from the frame the step consumed
        f = s % F
        base_force = tape.contact[f][:, 0, :]
        term = torch.norm(base_force, dim=-1) > 1.0
        tout = (ep_before + 1).float() > full.max_episode_length
        assert torch.equal(full.time_out_buf, tout)
        assert torch.equal(rst, term | tout)
        assert torch.equal(extras["time_outs"], tout)
        # counters: reset envs restart at 0, the others advance by one
        assert torch.equal(full.episode_length_buf, torch.where(rst, torch.zeros_like(ep_before), ep_before + 1))
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        assert float(obs.abs().max()) <= full.cfg.normalization.clip_observations
    assert int(full.reset_buf.sum()) > 0


def test_rom_rollout_at_1m_envs_is_deterministic_and_consistent():
    from legged_gym_dev_b200 import configs
    from legged_gym_dev_b200.rom import CustomSim
    N, T = N_FULL, 12
    logs = []
    for _ in range(2):
        env = CustomSim(configs.double_single_int_cfg(N, seed=4), device="cuda")
        obs = torch.zeros(N, 8, device="cuda")
        logs.append((env.collect_epoch(obs, T), obs.clone(), env))
    (l1, o1, e1), (l2, o2, _) = logs
    for k in ("z", "v", "pz_x", "done"):
        assert torch.equal(l1[k], l2[k]), k
    assert torch.equal(o1, o2)
    assert not bool(l1["done"].any())                       # CustomSim never terminates (custom_sim.py:75)
    vmax = torch.tensor([0.2, 0.2], device="cuda")
    assert bool((l1["v"].abs() <= vmax + 1e-6).all())       # generator inputs respect the ROM's input bounds
    # single-integrator ROM (rom_dynamics.py:188-193): the logged reference point (oldest, interpolated horizon entry,
    # data_collection_trajectory.py:145) moves at most two ROM steps' worth (dt_rom * v_max each) between two log rows
    dz = l1["z"][:, 1:] - l1["z"][:, :-1]
    assert float(dz.abs().max()) <= 2 * 0.1 * 0.2 + 1e-6
    assert torch.equal(l1["z"][:, -1], e1.trajectory[:, 0])


def test_sliding_window_at_scale_matches_torch_gather():
    from legged_gym_dev_b200 import datasets as DS
    B, T, D, N, dN, m = 1 << 17, 64, 3, 10, 1, 2
    data = torch.randn(B, T, D, device="cuda")
    got = DS.sliding_window(data, N, dN, m)
    start = data[:, :1, :].clone()
    start[:, :, -m:] = 0
    for i in (0, 1, 5, 9):
        want = torch.cat((start.expand(B, i, D), data[:, :T - i, :]), dim=1)       # dN == 1: slice i is the log shifted by i
        assert torch.equal(got[:, :, i * D:(i + 1) * D], want), i


def test_trajectory_env_at_1m_envs_shard_invariance_and_identities():
    """SURVEY 8f row 1 at 1 048 576 envs: the upper half stepped on its own (global ids) is bit-identical, generator state included;
    flags, counters, push timers and prev_error obey the identities of legged_robot_trajectory.py:150-246."""
    from legged_gym_dev_b200 import configs, synthetic as S
    from legged_gym_dev_b200.legged_robot_trajectory import AnymalTrajectory
    from legged_gym_dev_b200.physics import ReplayPhysics
    import legged_case as LC
    N, F = N_FULL, 2
    tape = S.make_state_tape(N, frames=F, seed=12, device="cuda")
    ep = S.make_episode_lengths(N, seed=4, device="cuda")
    tpush = 0.1 * torch.rand(N, 1, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))

    def make(n, tp, e, tp_push, offset, origins=None):
        cfg = configs.anymal_c_flat_trajectory_cfg()
        for k, v in configs.TRAJECTORY_ALL_REWARD_SCALES.items():
            setattr(cfg.rewards.scales, k, v)
        cfg.control.use_actuator_network = False
        cfg.env.num_envs = n
        env = AnymalTrajectory(cfg, SimpleNamespace(dt=cfg.sim.dt), None, "cuda", True, physics=ReplayPhysics(tp, device="cuda", copy=True),
                               asset=LC.dof_limits(), seed=0, env_id_offset=offset)
        env.episode_length_buf.copy_(e)
        env.time_until_next_push.copy_(tp_push)
        if origins is not None:
            env.env_origins.copy_(origins)
        env.reset_traj(torch.arange(n, device="cuda"))
        return env

    full = make(N, tape, ep, tpush, 0)
    lo = N // 2
    part = make(N - lo, _half(tape, lo, N), ep[lo:], tpush[lo:], lo, origins=full.env_origins[lo:])
    for s in range(3):
        ep_before, tp_before = full.episode_length_buf.clone(), full.time_until_next_push.clone()
        traj_prev_gen = full.traj_gen.k.clone()
        a = tape.actions[s % F]
        obs, _, rew, rst, extras = full.step(a)
        obs2, _, rew2, rst2, _ = part.step(a[lo:])
        assert torch.equal(obs[lo:], obs2) and torch.equal(rew[lo:], rew2) and torch.equal(rst[lo:], rst2)
        for name in ("trajectory", "prev_error", "time_until_next_push"):
            assert torch.equal(getattr(full, name)[lo:], getattr(part, name)), name
        for name in ("trajectory", "k", "t", "t_final", "weights", "rng_ctr", "stationary_inds"):
            assert torch.equal(getattr(full.traj_gen, name)[lo:], getattr(part.traj_gen, name)), "traj_gen." + name
        f = s % F
        term = torch.norm(tape.contact[f][:, 0, :], dim=-1) > 1.0
        tout = (ep_before + 1).float() > full.max_episode_length
        assert torch.equal(rst, term | tout) and torch.equal(extras["time_outs"], tout)
        assert torch.equal(full.episode_length_buf, torch.where(rst, torch.zeros_like(ep_before), ep_before + 1))
        # push timers: decremented by dt; the ones that fired were redrawn inside [0.5, 10) s
        dec = tp_before - 0.02
        fired = (dec <= 0).reshape(-1)
        assert torch.equal(full.time_until_next_push[~fired], dec[~fired])
        tf = full.time_until_next_push[fired]
        assert bool((tf >= 0.5).all()) and bool((tf < 10.0).all()) and int(fired.sum()) > 0
        # prev_error only changes for envs that reset, where it is (stale trajectory knot - new root position)^2
        want = torch.square(full.trajectory[rst, 0, :] - full.root_states[rst, :2])
        assert torch.allclose(full.prev_error[rst], want, rtol=1e-5, atol=1e-6)
        # generators of reset envs restart at k = 0 after their N warm-up knots; the others advance by at most one knot
        assert bool((full.traj_gen.k[rst] == 0).all())
        dk = full.traj_gen.k[~rst] - traj_prev_gen[~rst]
        assert bool(((dk == 0) | (dk == 1)).all())
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        assert float(obs.abs().max()) <= full.cfg.normalization.clip_observations
    assert int(full.reset_buf.sum()) > 0


def test_rough_lstm_step_at_cfg3_size_matches_oracle():
    """BASELINE.json configs[2] at its stated size: anymal_c_rough with the actuator-net torques, the 187-point height scan (terrain windows
    staged by TMA) and every reward term, 16 384 envs x 10 steps, every tensor against the oracle port (flags / height cells exact, 1e-5 else)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import legged_case as LC
    N = 16384
    case = LC.build_case("rough_lstm_allterms", N)
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    resets = 0
    for s in range(10):
        a = case.tape.actions[s % case.tape.frames]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        resets += int(port.reset_buf.sum())
        LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"cfg3 size step {s}: ")
    assert resets > 0
