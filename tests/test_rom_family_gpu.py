"""SURVEY 8f row 4 on the GPU: the generic rom-family kernels (csrc/rom_family.cu, through the C ABI) against the CPU oracle port and the
reference-generated fixture.  Masks / clocks / counters (t, k, stationary flags, draw-event counters) bit-exact; fp32 state within 1e-5
(S = 1).  For the integrator classes the generic kernels must equal the register-resident ones of csrc/rom.cu bit for bit."""

import numpy as np
import pytest
import torch

from legged_gym_dev_b200 import rom as R
from oracle.compare import assert_close, assert_exact
from oracle.make_golden_rom_family import CLASSES, algebra_inputs, case_params
from oracle.port_rom import FAMILY, GenPort, Rom, gen_params
from test_rom_family_cpu import GOLD, check_algebra, check_snapshot

pytestmark = pytest.mark.gpu
SAMPLERS = {"UniformWeightSampler": R.UniformWeightSampler, "UniformWeightSamplerNoExtreme": R.UniformWeightSamplerNoExtreme,
            "UniformWeightSamplerNoRamp": R.UniformWeightSamplerNoRamp}
PARAM_NAMES = ("sample_hold_input", "extreme_input", "ramp_v_start", "ramp_v_end", "ramp_t_start", "sin_mag", "sin_freq", "sin_off", "sin_mean")


def make_rom(p, n_robots=None):
    return R.ROM_CLASSES[p.rom_cls](p.rom_dt, p.z_min, p.z_max, p.v_min, p.v_max, n_robots=n_robots or p.num_envs, backend="torch", device="cuda")


def make_gen(p, env_id_offset=0, n_robots=None, **kw):
    return R.TrajectoryGenerator(make_rom(p, n_robots), R.UniformSampleHoldDT(p.t_low, p.t_high), SAMPLERS[p.weight_sampler](), dt_loop=p.dt_loop, N=p.N,
                                 freq_low=p.freq_low, freq_high=p.freq_high, seed=p.seed, backend="torch", device="cuda",
                                 prob_stationary=p.prob_stationary, dN=p.dN, env_id_offset=env_id_offset, **kw)


def gen_view(g):
    c = lambda t: t.detach().cpu()
    return dict(traj=c(g.trajectory), vtraj=c(g.v_trajectory), v=c(g.v), t=c(g.t), k=c(g.k), t_final=c(g.t_final), weights=c(g.weights),
                stationary=c(g.stationary_inds), get_trajectory=c(g.get_trajectory()))


def compare(port, g, tag):
    c = lambda t: t.detach().cpu()
    assert_exact(c(g.t), port.t, tag + "t")
    assert_exact(c(g.k), port.k, tag + "k")
    assert_exact(c(g.stationary_inds), port.stationary, tag + "stationary")
    assert_exact(c(g.rng_ctr).long(), torch.from_numpy(port.ctr), tag + "draw-event counters")
    for a, b in (("t_final", "t_final"), ("trajectory", "traj"), ("v_trajectory", "v_traj"), ("v", "v"), ("weights", "weights")):
        assert_close(c(getattr(g, a)), getattr(port, b), 1.0, tag + a)
    for name in PARAM_NAMES:
        assert_close(c(getattr(g, name)), getattr(port, name), 1.0, tag + name)
    assert_close(c(g.get_trajectory()), port.get_trajectory(), 1.0, tag + "get_trajectory()")


@pytest.mark.parametrize("cls", CLASSES)
def test_fused_matches_reference_golden(cls):
    g = np.load(GOLD)
    p = case_params(cls)
    check_algebra(make_rom(p), g, cls, to_dev=lambda t: t.cuda(), tag="fused ")
    gen = make_gen(p)
    assert_close(gen.ramp_v_end.cpu(), g[f"{cls}_ramp_v_end0"], 1.0, "ramp_v_end at construction")
    gen.reset(torch.from_numpy(g[f"{cls}_z0"]).cuda())
    check_snapshot(gen_view(gen), g, f"{cls}_reset", f"{cls} after reset: ")
    keep = [int(k) for k in g["keep"]]
    for s in range(max(keep) + 1):
        gen.step()
        if s in keep:
            check_snapshot(gen_view(gen), g, f"{cls}_s{s}", f"{cls} step {s}: ")
    assert np.array_equal(gen.rng_ctr.cpu().numpy(), g[f"{cls}_ctr"])


@pytest.mark.parametrize("cls", CLASSES)
@pytest.mark.parametrize("rows", [1, 257, 5000])
def test_algebra_parity_ragged(cls, rows):
    p = case_params(cls)
    n, m = (FAMILY.get(cls, (2 if cls == "SingleInt2D" else 4, 2)))[:2]
    z, v, x = algebra_inputs(cls, n, m, rows=rows, seed=rows)
    z[:, n - 1] *= 3.0                      # push velocity states / headings well past their bounds
    port, rom = Rom(cls, p.rom_dt, p.z_min, p.z_max, p.v_min, p.v_max), make_rom(p)
    zc, vc, xc = z.cuda(), v.cuda(), x.cuda()
    assert_close(rom.f(zc, vc), port.f(z, v), 1.0, f"{cls}.f")
    for a, b, name in zip(rom.des_pose_vel(zc, vc), port.des_pose_vel(z, v), ("pose", "vel")):
        assert_close(a, b, 1.0, f"{cls}.des_pose_vel {name}")
    for a, b, name in zip(rom.compute_state_dependent_input_bounds(zc), port.bounds(z), ("lo", "hi")):
        assert_close(a, b, 1.0, f"{cls} bounds {name}")
    assert_close(rom.clip_v_z(zc, vc), port.clip_v_z(z, v), 1.0, f"{cls}.clip_v_z")
    assert_close(rom.clip_v(vc), torch.max(torch.min(v, port.v_max), port.v_min), 1.0, f"{cls}.clip_v")
    assert_close(rom.proj_z(xc), np.asarray(port.proj_z(x)), 1.0, f"{cls}.proj_z")
    assert_exact(rom.vel_inds.cpu(), port.vel_inds, "vel_inds")


def test_algebra_rejects_bad_input():
    p = case_params("Unicycle")
    rom = make_rom(p)
    with pytest.raises(RuntimeError, match="CUDA only"):
        rom.f(torch.zeros(4, 3), torch.zeros(4, 2))
    with pytest.raises(ValueError, match="float32"):
        rom.f(torch.zeros(4, 4, device="cuda"), torch.zeros(4, 2, device="cuda"))
    assert rom.f(torch.zeros(0, 3, device="cuda"), torch.zeros(0, 2, device="cuda")).shape == (0, 3)
    with pytest.raises(ValueError, match="entries"):
        R.Unicycle(0.1, [0, 0], [1, 1], [0, 0], [1, 1], device="cuda")
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        R.Unicycle(0.1, [0] * 3, [1] * 3, [0] * 2, [1] * 2, device="cpu")


@pytest.mark.parametrize("cls,N,over", [
    ("Unicycle", 300, {}), ("LateralUnicycle", 77, dict(weight_sampler="UniformWeightSamplerNoRamp")),
    ("ExtendedUnicycle", 1000, dict(t_low=0.05, t_high=0.3, prob_stationary=0.2)), ("ExtendedLateralUnicycle", 513, dict(N=4, dN=3)),
    ("ExtendedLateralUnicycle", 130, dict(dt_loop=0.1, weight_sampler="UniformWeightSamplerNoExtreme")),
    ("ExtendedLateralUnicycle", 260, dict(N=10)), ("ExtendedLateralUnicycle", 140, dict(N=16, dN=2)),     # 48 KB / 150 KB of staged windows per CTA
    ("SingleInt2D", 200, {}), ("DoubleInt2D", 200, dict(N=10, dt_loop=0.05))])
def test_generator_stepwise_parity(cls, N, over):
    """reset(z), steps with dt_loop < rom.dt (envs only advance when their ROM clock is due), a partial reset_idx whose warm-up re-evaluates
    every env, masked step_idx calls that put the envs' clocks out of phase, get_input_t — each compared tensor by tensor."""
    p = gen_params(N, cls, seed=9, **over)
    port, gen = GenPort(p), make_gen(p)
    assert_close(gen.ramp_v_end.cpu(), port.ramp_v_end, 1.0, "ramp_v_end at construction")
    rng = torch.Generator().manual_seed(N)
    z0 = torch.randn(N, port.rom.n, generator=rng) * 0.4
    port.reset(z0.clone())
    gen.reset(z0.cuda())
    compare(port, gen, f"{cls} after reset: ")
    seen_stationary = 0
    for s in range(240):
        if s == 100:
            ids = torch.arange(0, N, 3)
            z1 = torch.randn(N, port.rom.n, generator=rng) * 0.4
            port.gen_reset_idx(ids, z1.clone())
            gen.reset_idx(ids.cuda(), z1.cuda())
            compare(port, gen, f"{cls} partial reset: ")
        if s in (40, 41, 150):     # step a subset only: clocks go out of phase
            ids = torch.arange(s % 2, N, 2)
            port.gen_step_idx(ids)
            gen.step_idx(ids.cuda())
        else:
            port.step()
            gen.step()
        compare(port, gen, f"{cls} step {s}: ")
        seen_stationary += int(port.stationary.sum())
    assert port.ctr.max() > 4 and seen_stationary > 0 and not bool(port.stationary.all())
    # open-loop use (trajopt/trajectory_gen.py:35-41): the caller owns clock and state
    for tq in (float(port.t.max()) + 0.05, float(port.t.max()) + 2.5):
        zq = torch.randn(N, port.rom.n, generator=rng) * 0.4
        want = port.get_input_t(torch.full((N,), tq), zq)
        got = gen.get_input_t(tq, zq.cuda())
        assert_close(got, want, 1.0, f"{cls} get_input_t({tq})")
        assert_exact(gen.rng_ctr.cpu().long(), torch.from_numpy(port.ctr), "draw-event counters after get_input_t")
        assert_close(gen.t_final.cpu(), port.t_final, 1.0, "t_final after get_input_t")


@pytest.mark.parametrize("cls,over", [("SingleInt2D", {}), ("DoubleInt2D", {}), ("SingleInt2D", dict(N=5, dN=2)), ("DoubleInt2D", dict(N=10, t_low=0.05, t_high=0.3))])
def test_generic_kernels_equal_register_kernels_bitwise(cls, over):
    """rom types 0 / 1 are served by both kernel sets: same reset, then TrajectoryGenerator.step through b200gym_rom_step (horizon in
    registers) and through b200gym_romfam_gen_step (horizon in HBM) — every tensor identical, every step."""
    N = 777
    p = gen_params(N, cls, seed=4, **over)
    a, b = make_gen(p, generic_kernels=False), make_gen(p, generic_kernels=True)
    assert not a._family and b._family and make_gen(p)._family      # stand-alone generators default to the window-staging kernels
    z0 = (torch.randn(N, a.rom.n, generator=torch.Generator().manual_seed(2)) * 0.4).cuda()
    a.reset(z0)
    b.reset(z0)
    mask = torch.arange(N, device="cuda") % 3 != 0
    for s in range(150):
        if s in (20, 21):
            a.step_idx(mask)
            b.step_idx(mask)
        else:
            a.step()
            b.step()
        for name in ("trajectory", "v_trajectory", "v", "t", "k", "t_final", "weights", "stationary_inds", "rng_ctr") + PARAM_NAMES:
            assert torch.equal(getattr(a, name), getattr(b, name)), f"{cls} step {s}: {name} differs between the two kernel sets"
    assert int(a.rng_ctr.max()) > 3


def test_generator_shard_invariance_and_bounds_1m():
    """BASELINE cfg 4 size (1 048 576 envs), the 6-state / 3-input class: two half-size generators with global env ids equal the whole
    (no cross-env term, RNG keyed by global id); velocity states never leave their bounds (clip_v_z); stationary envs hold zero input;
    every env's window holds W+1 finite knots."""
    N, cls = 1 << 20, "ExtendedLateralUnicycle"
    p = gen_params(N, cls, seed=12, prob_stationary=0.1, t_low=0.1, t_high=0.4)
    whole = make_gen(p)
    halves = [make_gen(p, env_id_offset=o, n_robots=N // 2) for o in (0, N // 2)]
    z0 = (torch.randn(N, 6, generator=torch.Generator().manual_seed(5)) * 0.3).cuda()
    z0[:, 3:] = z0[:, 3:].clamp(-0.4, 0.4)
    whole.reset(z0)
    for h, sl in zip(halves, (slice(0, N // 2), slice(N // 2, N))):
        h.reset(z0[sl].contiguous())
    for _ in range(40):
        whole.step()
        for h in halves:
            h.step()
    for name in ("trajectory", "v_trajectory", "v", "t", "k", "t_final", "weights", "stationary_inds", "rng_ctr"):
        assert torch.equal(getattr(whole, name), torch.cat([getattr(h, name) for h in halves])), f"{name}: shards differ from the whole"
    tr = whole.trajectory
    assert bool(torch.isfinite(tr).all()) and bool(torch.isfinite(whole.v_trajectory).all())
    zmax = torch.tensor(p.z_max[3:], device="cuda")
    assert bool((tr[:, -1, 3:].abs() <= zmax * (1 + 1e-5) + 1e-6).all()), "velocity states left their bounds"
    st = whole.stationary_inds
    assert 0.02 < float(st.float().mean()) < 0.3
    assert bool((whole.v[st] == 0).all())
    assert 7 <= int(whole.k.min()) and int(whole.k.max()) <= 9      # 40 loop steps of 0.02 s on a 0.1 s ROM clock
