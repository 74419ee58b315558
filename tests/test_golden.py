"""Golden vectors written by the UNMODIFIED reference (oracle/make_golden.py) replayed through
(a) the CPU oracle port  [-m "not gpu"]  and (b) the fused CUDA path through the C ABI  [-m gpu]."""
import glob
import os

import numpy as np
import pytest
import torch

import legged_case as LC
from legged_gym_dev_b200 import synthetic as S
from oracle.compare import assert_close, assert_exact

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LEGGED = sorted(os.path.basename(p)[len("legged_"):-4] for p in glob.glob(os.path.join(GOLD, "legged_*.npz")))


def _load_case(name):
    g = np.load(os.path.join(GOLD, f"legged_{name}.npz"))
    N, frames = int(g["num_envs"]), int(g["frames"])
    case = LC.build_case(name, N, frames=frames, seed=int(g["seed"]))
    case.tape.root = torch.from_numpy(g["tape_root"])
    case.tape.dof = torch.from_numpy(g["tape_dof"])
    case.tape.contact = torch.from_numpy(g["tape_contact"])
    case.tape.actions = torch.from_numpy(g["tape_actions"])
    case.ep = torch.from_numpy(g["ep"])
    if case.traj:
        case.tpush = torch.from_numpy(g["tpush"])
    if case.rough:
        case.terrain = {k[len("terrain_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("terrain_")}
    return case, g


def _check(snap, g, s, tag):
    keys = [k[len(f"s{s}_"):] for k in g.files if k.startswith(f"s{s}_")]
    assert keys
    for k in keys:
        want = torch.from_numpy(g[f"s{s}_{k}"])
        got = snap[k]
        if k in LC.EXACT:
            assert_exact(got.to(want.dtype), want, f"{tag}{k}")
        elif k == "obs" and "trajectory" in snap:     # the trajectory block is a difference of positions: see LC.compare_snapshots
            LC.compare_snapshots(dict(obs=got, trajectory=snap["trajectory"]),
                                 dict(obs=want, trajectory=torch.from_numpy(g[f"s{s}_trajectory"])), tag)
        else:
            assert_close(got.reshape(want.shape), want, LC.SCALES.get(k, 1.0), f"{tag}{k}")


@pytest.mark.parametrize("name", LEGGED)
def test_port_matches_reference_golden(name):
    case, g = _load_case(name)
    port, phys = LC.make_port(case)
    if case.traj:
        port.gen.reset_traj(torch.arange(case.num_envs), port.proj_z())
    for s in range(int(g["steps"])):
        a = case.tape.actions[s % case.tape.frames] * (150.0 if s == 3 else 1.0)
        port.step(a.clone(), phys)
        _check(LC.snapshot_port(port), g, s, f"golden {name} step {s}: ")


@pytest.mark.gpu
@pytest.mark.parametrize("name", LEGGED)
def test_fused_matches_reference_golden(name):
    case, g = _load_case(name)
    env = LC.make_fused(case)
    if case.traj:
        env.reset_traj(torch.arange(case.num_envs, device="cuda"))
    for s in range(int(g["steps"])):
        a = case.tape.actions[s % case.tape.frames] * (150.0 if s == 3 else 1.0)
        env.step(a.cuda())
        _check(LC.snapshot_fused(env), g, s, f"golden {name} step {s}: ")


# ---------------------------------------------------------------------------------------------------
# Group M goldens (reference CustomSim + TrajectoryGenerator + DoubleSingleTracking)
# ---------------------------------------------------------------------------------------------------
ROM = sorted(os.path.basename(p)[len("rom_"):-4] for p in glob.glob(os.path.join(GOLD, "rom_*.npz")))
ROM_OVER = {"default": {}, "fast_resample": dict(prob_stationary=0.05, t_low=0.2, t_high=0.5)}
ROM_EXACT = {"t", "k", "stationary"}


def _rom_check(get, g, s, tag):
    for key in ("action", "obs", "root", "traj", "vtraj", "v", "t", "k", "t_final", "weights", "stationary", "env_trajectory"):
        want = torch.from_numpy(g[f"s{s}_{key}"])
        got = get(key)
        if key in ROM_EXACT:
            assert_exact(got.to(want.dtype), want, f"{tag}{key}")
        else:
            assert_close(got, want, 1.0, f"{tag}{key}")


@pytest.mark.parametrize("name", ROM)
def test_rom_port_matches_reference_golden(name):
    from oracle.port_rom import RomPort, rom_params
    g = np.load(os.path.join(GOLD, f"rom_{name}.npz"))
    N, steps = int(g["num_envs"]), int(g["steps"])
    port = RomPort(rom_params(N, seed=int(g["seed"]), **ROM_OVER[name]))
    assert_close(port.ramp_v_end, torch.from_numpy(g["ramp_v_end0"]), 1.0, "ramp_v_end0")
    obs, _ = port.reset()
    keep = set(int(k) for k in g["keep"])
    for s in range(steps):
        if s == 60:
            obs, _ = port.reset_idx(torch.arange(0, N, 3))
        a = port.policy(obs)
        obs, _ = port.step(a)
        if s in keep:
            view = dict(action=a, obs=obs, root=port.root_states, traj=port.traj, vtraj=port.v_traj, v=port.v, t=port.t, k=port.k,
                        t_final=port.t_final, weights=port.weights, stationary=port.stationary, env_trajectory=port.trajectory)
            _rom_check(lambda k: view[k], g, s, f"rom golden {name} step {s}: ")
    assert np.array_equal(port.ctr, g["ctr"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ROM)
def test_rom_fused_matches_reference_golden(name):
    from legged_gym_dev_b200 import configs
    from legged_gym_dev_b200.rom import CustomSim, DoubleSingleTracking
    g = np.load(os.path.join(GOLD, f"rom_{name}.npz"))
    N, steps = int(g["num_envs"]), int(g["steps"])
    env = CustomSim(configs.double_single_int_cfg(N, seed=int(g["seed"]), **ROM_OVER[name]), device="cuda")
    policy = DoubleSingleTracking(10, 10, env.model.clip_v_z)
    assert_close(env.traj_gen.ramp_v_end.cpu(), torch.from_numpy(g["ramp_v_end0"]), 1.0, "ramp_v_end0")
    env.reset()
    obs = env.get_observations()
    keep = set(int(k) for k in g["keep"])
    for s in range(steps):
        if s == 60:
            env.reset_idx(torch.arange(0, N, 3, device="cuda"))
            obs = env.get_observations()
        a = policy(obs)
        a_keep = a.clone()
        obs, _, _, _, _ = env.step(a)
        if s in keep:
            tg = env.traj_gen
            view = dict(action=a_keep, obs=obs, root=env.root_states, traj=tg.trajectory, vtraj=tg.v_trajectory, v=tg.v, t=tg.t, k=tg.k,
                        t_final=tg.t_final, weights=tg.weights, stationary=tg.stationary_inds, env_trajectory=env.trajectory)
            _rom_check(lambda k: view[k].detach().cpu(), g, s, f"rom golden {name} step {s}: ")
    assert np.array_equal(env.traj_gen.rng_ctr.cpu().numpy().astype(np.int64), g["ctr"])
