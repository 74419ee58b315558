"""SURVEY 8f row 4 on the CPU: the oracle's restatement of the whole RomDynamics family (oracle/port_rom.py) against the UNMODIFIED
reference classes (build container) and against the reference-generated fixture tests/golden/romfam_reference.npz (everywhere)."""
import os

import numpy as np
import pytest
import torch

from oracle.compare import assert_close, assert_exact
from oracle.make_golden_rom_family import CLASSES, algebra_inputs, case_params
from oracle.port_rom import FAMILY, GenPort, Rom, gen_params

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "romfam_reference.npz")
SNAP_EXACT = ("t", "k", "stationary")
SNAP_CLOSE = ("traj", "vtraj", "v", "t_final", "weights", "get_trajectory")


def port_view(port):
    return dict(traj=port.traj, vtraj=port.v_traj, v=port.v, t=port.t, k=port.k, t_final=port.t_final, weights=port.weights,
                stationary=port.stationary, get_trajectory=port.get_trajectory())


def check_snapshot(view, g, prefix, tag):
    for key in SNAP_EXACT:
        assert_exact(view[key].to(torch.from_numpy(g[f"{prefix}_{key}"]).dtype), g[f"{prefix}_{key}"], f"{tag}{key}")
    for key in SNAP_CLOSE:
        assert_close(view[key], g[f"{prefix}_{key}"], 1.0, f"{tag}{key}")


def check_algebra(rom_like, g, cls, to_dev=lambda t: t, tag=""):
    """rom_like: an object with the RomDynamics methods; compares every recorded output of the reference class."""
    z, v, x = (to_dev(torch.from_numpy(g[f"{cls}_alg_{k}"])) for k in ("z", "v", "x"))
    pose, vel = rom_like.des_pose_vel(z, v)
    lo, hi = rom_like.compute_state_dependent_input_bounds(z) if hasattr(rom_like, "compute_state_dependent_input_bounds") else rom_like.bounds(z)
    got = dict(f=rom_like.f(z, v), pose=pose, vel=vel, lo=lo, hi=hi, clip=rom_like.clip_v_z(z, v))
    if f"{cls}_alg_proj" in g.files:
        got["proj"] = rom_like.proj_z(x)
    for k, t in got.items():
        assert_close(t, g[f"{cls}_alg_{k}"], 1.0, f"{tag}{cls}.{k}")


@pytest.mark.parametrize("cls", CLASSES)
def test_port_matches_reference_golden(cls):
    g = np.load(GOLD)
    p = case_params(cls)
    check_algebra(Rom(cls, p.rom_dt, p.z_min, p.z_max, p.v_min, p.v_max), g, cls, tag="port ")
    port = GenPort(p)
    assert_exact(port.ramp_v_end, g[f"{cls}_ramp_v_end0"], "ramp_v_end at construction")
    port.reset(torch.from_numpy(g[f"{cls}_z0"]).clone())
    check_snapshot(port_view(port), g, f"{cls}_reset", f"{cls} after reset: ")
    keep = [int(k) for k in g["keep"]]
    for s in range(max(keep) + 1):
        port.step()
        if s in keep:
            check_snapshot(port_view(port), g, f"{cls}_s{s}", f"{cls} step {s}: ")
    assert np.array_equal(port.ctr, g[f"{cls}_ctr"])


@pytest.mark.reference
@pytest.mark.parametrize("cls", CLASSES)
def test_port_tracks_unmodified_reference_generator(cls):
    """Every generator tensor after every one of 200 steps, bit for bit (same torch CPU arithmetic under the same draws)."""
    from oracle import ref_harness as H
    N = 64
    p = gen_params(N, cls, seed=5, **({} if cls not in FAMILY else dict(dt_loop=0.1)))   # the reference's unicycle f() needs every env due
    tg, rom = H.make_reference_generator(p, seed=5)
    port = GenPort(p)
    assert_exact(port.ramp_v_end, tg.ramp_v_end, "ramp_v_end at construction")
    z0 = torch.randn(N, rom.n, generator=torch.Generator().manual_seed(1)) * 0.3
    tg.reset(z0.clone())
    port.reset(z0.clone())
    for s in range(200):
        tg.step()
        port.step()
        for a, b, name in ((port.t, tg.t, "t"), (port.k, tg.k, "k"), (port.stationary, tg.stationary_inds, "stationary"),
                           (port.traj, tg.trajectory, "trajectory"), (port.v_traj, tg.v_trajectory, "v_trajectory"), (port.v, tg.v, "v"),
                           (port.weights, tg.weights, "weights"), (port.t_final, tg.t_final, "t_final"),
                           (port.get_trajectory(), tg.get_trajectory(), "get_trajectory()")):
            assert_exact(a, b, f"{cls} step {s}: {name}")
    assert np.array_equal(port.ctr, tg._shim.ctr)
    assert port.ctr.max() > 3 and port.stationary.any()


@pytest.mark.reference
@pytest.mark.parametrize("cls", CLASSES)
def test_port_algebra_equals_reference_classes(cls):
    from oracle import ref_harness as H
    rd = H.import_reference().rom_dynamics
    p = case_params(cls)
    t = lambda v: torch.tensor(v, dtype=torch.float32)
    z, v, x = algebra_inputs(cls, *(FAMILY.get(cls, (2 if cls == "SingleInt2D" else 4, 2))[:2]), rows=96, seed=23)
    ref = getattr(rd, cls)(p.rom_dt, t(p.z_min), t(p.z_max), t(p.v_min), t(p.v_max), n_robots=z.shape[0], backend="torch", device="cpu")
    port = Rom(cls, p.rom_dt, p.z_min, p.z_max, p.v_min, p.v_max)
    assert_exact(port.f(z, v), ref.f(z, v), "f")
    for a, b, name in zip(port.des_pose_vel(z, v), ref.des_pose_vel(z, v), ("pose", "vel")):
        assert_exact(a, b, name)
    for a, b, name in zip(port.bounds(z), ref.compute_state_dependent_input_bounds(z), ("lo", "hi")):
        assert_exact(a, b, name)
    assert_exact(port.clip_v_z(z, v), ref.clip_v_z(z, v), "clip_v_z")
    assert_exact(port.vel_inds, ref.vel_inds, "vel_inds")
    if cls != "ExtendedLateralUnicycle":     # its proj_z raises in the reference (torch.squeeze of a numpy array, :426)
        ref_np = getattr(rd, cls)(p.rom_dt, np.array(p.z_min), np.array(p.z_max), np.array(p.v_min), np.array(p.v_max), n_robots=z.shape[0],
                                  backend="numpy")
        assert np.array_equal(np.asarray(port.proj_z(x)), np.asarray(ref_np.proj_z(x.numpy())))


def test_extended_lateral_proj_z_extends_extended_unicycle():
    """The one proj_z without a reference output: columns 0-3 and the last equal ExtendedUnicycle's, column 4 is the lateral component."""
    _, _, x = algebra_inputs("ExtendedLateralUnicycle", 6, 3)
    a = Rom("ExtendedUnicycle", 0.1, [0] * 5, [0] * 5, [0] * 2, [0] * 2).proj_z(x)
    b = Rom("ExtendedLateralUnicycle", 0.1, [0] * 6, [0] * 6, [0] * 3, [0] * 3).proj_z(x)
    assert np.array_equal(a[:, :4], b[:, :4]) and np.array_equal(a[:, 4], b[:, 5])
    yaw = b[:, 2]
    v = x.numpy()[:, 7:9].astype(np.float64)
    assert np.allclose(b[:, 3] ** 2 + b[:, 4] ** 2, (v ** 2).sum(1), rtol=1e-12)            # a rotation of (vx, vy)
    assert np.allclose(np.cos(yaw) * b[:, 3] - np.sin(yaw) * b[:, 4], v[:, 0], atol=1e-12)   # back in the world frame


def test_product_classes_fail_loudly_without_cuda():
    """No CPU fallback: the host mirrors of the family refuse CPU devices / CPU tensors instead of computing something else."""
    from legged_gym_dev_b200 import rom as R
    from legged_gym_dev_b200.hopper import HopperActuation
    for cls in ("Unicycle", "LateralUnicycle", "ExtendedUnicycle", "ExtendedLateralUnicycle"):
        n, m, _ = FAMILY[cls]
        assert (R.ROM_CLASSES[cls].n, R.ROM_CLASSES[cls].m) == (n, m)
        with pytest.raises(RuntimeError, match="CUDA devices only"):
            R.ROM_CLASSES[cls](0.1, [0.0] * n, [1.0] * n, [0.0] * m, [1.0] * m, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        R.TrajectoryGenerator(R.SingleInt2D(0.1, [0, 0], [1, 1], [0, 0], [1, 1], device="cpu"), R.UniformSampleHoldDT(1, 2), R.UniformWeightSampler(),
                              device="cpu")
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        HopperActuation(4, device="cpu")
    with pytest.raises(NameError, match="Unknown controller type"):
        HopperActuation(4, control_type="V")


def test_c_abi_argument_checks_need_no_gpu():
    """The argument validation of the new entry points runs before any CUDA call: wrong rom types / null buffers / bad windows come back as
    B200GYM_EINVAL with a message, on a machine without a GPU too."""
    import ctypes as C
    from legged_gym_dev_b200 import _lib
    L = _lib.lib()
    err = lambda: L.b200gym_last_error().decode()
    one = C.c_void_p(16)       # never dereferenced: the checks reject the call first
    assert L.b200gym_romfam_f(9, 0.1, one, one, one, 4, None) == -1 and "unknown rom_type 9" in err()
    assert L.b200gym_romfam_f(2, 0.1, None, one, one, 4, None) == -1 and "null argument" in err()
    assert L.b200gym_romfam_proj_z(3, one, one, 0, None) == -1 and "n_rows" in err()
    p = _lib.RomFamilyParamsPOD()
    s = _lib.RomStatePOD()
    p.num_envs, p.rom_type, p.window, p.dN, p.rom_dt = 8, 5, 1, 1, 0.1
    assert L.b200gym_romfam_gen_step(p, s, None, 0, None) == -1 and "window" in err()
    p.window = 6
    assert L.b200gym_romfam_gen_step(p, s, None, 0, None) == -1 and "null state tensor" in err()
    p.rom_dt = 0.0
    assert L.b200gym_romfam_input_bounds(p, one, None, one, one, None, 4, None) == -1 and "rom_dt" in err()
    hp, hb = _lib.HopperTorqueParamsPOD(), _lib.HopperTorqueBuffersPOD()
    hp.num_envs, hp.num_bodies, hp.foot_body = 8, 5, 7
    assert L.b200gym_hopper_torques(hp, hb, None) == -1 and "foot_body" in err()
    hp.foot_body = 4
    assert L.b200gym_hopper_torques(hp, hb, None) == -1 and "null buffer" in err()
    assert L.b200gym_hopper_reward_terms(8, 0.0, one, one, one, one, one, None) == -1 and "positive" in err()
