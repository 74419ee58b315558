"""Group M parity on the GPU: fused ROM kernels (through the C ABI) vs the CPU oracle port.
Masks / counters (k, t, stationary flags, draw-event counters) bit-exact; fp32 state within 1e-5 (S = 1)."""
import pytest
import torch

from legged_gym_dev_b200 import configs
from oracle.compare import assert_close, assert_exact
from oracle.port_rom import RomPort, rom_params

pytestmark = pytest.mark.gpu


def make_pair(num_envs, seed=3, **over):
    N = num_envs
    from legged_gym_dev_b200.rom import CustomSim, DoubleSingleTracking
    port = RomPort(rom_params(num_envs, seed=seed, **{("weight_sampler" if k == "weight_samp_cls" else k): v for k, v in over.items()}))
    env = CustomSim(configs.double_single_int_cfg(num_envs, seed=seed, **over), device="cuda")
    policy = DoubleSingleTracking(10, 10, env.model.clip_v_z)
    return port, env, policy


def compare_state(port, env, tag):
    g = env.traj_gen
    c = lambda t: t.detach().cpu()
    assert_exact(c(g.t), port.t, tag + "t")
    assert_exact(c(g.k), port.k, tag + "k")
    assert_exact(c(g.stationary_inds), port.stationary, tag + "stationary")
    assert_exact(c(g.rng_ctr).long(), torch.from_numpy(port.ctr), tag + "draw-event counters")
    assert_close(c(g.t_final), port.t_final, 1.0, tag + "t_final")
    assert_close(c(env.root_states), port.root_states, 1.0, tag + "root_states")
    assert_close(c(g.trajectory), port.traj, 1.0, tag + "trajectory")
    assert_close(c(g.v_trajectory), port.v_traj, 1.0, tag + "v_trajectory")
    assert_close(c(g.v), port.v, 1.0, tag + "v")
    assert_close(c(g.weights), port.weights, 1.0, tag + "weights")
    assert_close(c(env.trajectory), port.trajectory, 1.0, tag + "CustomSim.trajectory")
    for a, b in (("sample_hold_input", "sample_hold_input"), ("extreme_input", "extreme_input"), ("ramp_v_start", "ramp_v_start"),
                 ("ramp_v_end", "ramp_v_end"), ("ramp_t_start", "ramp_t_start"), ("sin_mag", "sin_mag"), ("sin_freq", "sin_freq"),
                 ("sin_off", "sin_off"), ("sin_mean", "sin_mean")):
        assert_close(c(getattr(g, a)), getattr(port, b), 1.0, tag + a)
    assert_close(c(g.get_trajectory()), port.get_trajectory(), 1.0, tag + "get_trajectory()")


@pytest.mark.parametrize("N,over", [(256, {}), (1000, dict(prob_stationary=0.05, t_low=0.2, t_high=0.5)),
                                     (77, dict(weight_samp_cls="UniformWeightSamplerNoExtreme", randomize_rom_distance=False)),
                                     (130, dict(N=6, dN=2))])
def test_stepwise_parity(N, over):
    port, env, policy = make_pair(N, **over)
    assert_close(env.traj_gen.ramp_v_end.cpu(), port.ramp_v_end, 1.0, "ramp_v_end at construction")
    o2, _ = port.reset()
    env.reset()
    o1 = env.get_observations()
    compare_state(port, env, "after reset: ")
    for s in range(260):
        if s == 120:   # partial reset: side effects on every env
            ids = torch.arange(0, N, 3)
            o2, _ = port.reset_idx(ids)
            env.reset_idx(ids.cuda())
            o1 = env.get_observations()
            compare_state(port, env, f"partial reset: ")
        a2 = port.policy(o2)
        a1 = policy(o1)
        assert_close(a1.cpu(), a2, 1.0, f"step {s}: action")
        o2, _ = port.step(a2)
        o1, _, _, d1, _ = env.step(a1)
        assert_close(o1.cpu(), o2, 1.0, f"step {s}: obs")
        assert not bool(d1.any())
        if s % 20 == 0 or s > 250:
            compare_state(port, env, f"step {s}: ")
    compare_state(port, env, "end: ")


def test_cfg1_4096x1000():
    """BASELINE config 1: 4096 envs x 1000 loop steps; k must reach 500."""
    port, env, policy = make_pair(4096, seed=0)
    o2, _ = port.reset()
    env.reset()
    o1 = env.get_observations()
    for s in range(1000):
        o2, _ = port.step(port.policy(o2))
        o1, _, _, _, _ = env.step(policy(o1))
    compare_state(port, env, "cfg1 end: ")
    assert float(env.traj_gen.k[0]) == 500.0


@pytest.mark.parametrize("N,T,dbg", [(300, 40, True), (129, 7, False), (64, 0, True)])
def test_collect_epoch_matches_loop(N, T, dbg):
    """The persistent rollout kernel == the reference's while-loop epoch (data_collection_trajectory.py:104-149)."""
    port, env, policy = make_pair(N, seed=11)
    obs_p = torch.zeros(N, 8)
    obs_g = torch.zeros(N, 8, device="cuda")
    for epoch in range(2):   # epoch 1 starts from epoch 0's stale observation, like the reference
        want, obs_p = port.collect_epoch(obs_p, T)
        got = env.collect_epoch(obs_g, T, save_debugging_data=dbg)
        for k in ("z", "pz_x", "v") + (("x",) if dbg else ()):
            assert_close(got[k].cpu(), want[k], 1.0, f"epoch {epoch}: {k}")
        assert_exact(got["done"].cpu(), want["done"], f"epoch {epoch}: done")
        assert_close(obs_g.cpu(), obs_p, 1.0, f"epoch {epoch}: obs")
        compare_state(port, env, f"epoch {epoch}: ")


def test_rollout_equals_per_call_path():
    """Size-independent property at a larger N: one rollout launch == T x (policy + step) launches, bit for bit."""
    from legged_gym_dev_b200.rom import CustomSim, DoubleSingleTracking
    N, T = 20000, 12
    a = CustomSim(configs.double_single_int_cfg(N, seed=5), device="cuda")
    b = CustomSim(configs.double_single_int_cfg(N, seed=5), device="cuda")
    pol = DoubleSingleTracking(10, 10, b.model.clip_v_z)
    obs = torch.zeros(N, 8, device="cuda")
    log = a.collect_epoch(obs, T, save_debugging_data=True)
    b.reset()
    o = torch.zeros(N, 8, device="cuda")
    for t in range(T):
        k = b.traj_gen.k.clone()
        while bool(torch.any(b.traj_gen.k == k)):
            o, _, _, _, _ = b.step(pol(o))
        assert torch.equal(log["x"][:, t + 1], b.root_states)
        assert torch.equal(log["v"][:, t], b.traj_gen.v)
        # torch's CUDA division by a Python scalar multiplies by the reciprocal: 1-ulp close, not equal
        assert_close(log["z"][:, t + 1].cpu(), b.traj_gen.get_trajectory()[:, 0].cpu(), 1.0, "z log vs get_trajectory()")
        assert torch.equal(log["z"][:, t + 1], b.trajectory[:, 0])
    assert torch.equal(a.traj_gen.trajectory, b.traj_gen.trajectory)
    assert torch.equal(obs, o)


def test_shard_invariance():
    from legged_gym_dev_b200.rom import CustomSim
    N = 512
    full = CustomSim(configs.double_single_int_cfg(N, seed=9), device="cuda")
    half = CustomSim(configs.double_single_int_cfg(N // 2, seed=9), device="cuda", env_id_offset=N // 2)
    o1, o2 = torch.zeros(N, 8, device="cuda"), torch.zeros(N // 2, 8, device="cuda")
    l1, l2 = full.collect_epoch(o1, 20), half.collect_epoch(o2, 20)
    for k in ("z", "v", "pz_x"):
        assert torch.equal(l1[k][N // 2:], l2[k])
