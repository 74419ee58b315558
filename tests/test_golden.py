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
        else:
            assert_close(got.reshape(want.shape), want, LC.SCALES.get(k, 1.0), f"{tag}{k}")


@pytest.mark.parametrize("name", LEGGED)
def test_port_matches_reference_golden(name):
    case, g = _load_case(name)
    port, phys = LC.make_port(case)
    for s in range(int(g["steps"])):
        a = case.tape.actions[s % case.tape.frames] * (150.0 if s == 3 else 1.0)
        port.step(a.clone(), phys)
        _check(LC.snapshot_port(port), g, s, f"golden {name} step {s}: ")


@pytest.mark.gpu
@pytest.mark.parametrize("name", LEGGED)
def test_fused_matches_reference_golden(name):
    case, g = _load_case(name)
    env = LC.make_fused(case)
    for s in range(int(g["steps"])):
        a = case.tape.actions[s % case.tape.frames] * (150.0 if s == 3 else 1.0)
        env.step(a.cuda())
        _check(LC.snapshot_fused(env), g, s, f"golden {name} step {s}: ")
