"""RaibertHeuristic (deep_tube_learning/controllers.py:4-81, SURVEY 8f row 3's tracking controller): oracle port vs the reference's own
class and vs reference-generated goldens on CPU; the CUDA kernel through the C ABI vs the port and the goldens on the GPU."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import port_controllers as PC
from oracle.compare import assert_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raibert_reference.npz")


@pytest.mark.reference
def test_port_equals_reference_class():
    from oracle import ref_harness as H
    ref = H.import_reference()
    for gains in (PC.GAINS, PC.GAINS_B):
        obs = PC.sample_obs(1000, seed=11)
        want = ref.controllers.RaibertHeuristic.raibert_policy(obs, *gains.values())
        assert torch.equal(PC.raibert_policy(obs, **gains), want)
        cfg = SimpleNamespace(controller=SimpleNamespace(K_p=gains["Kp"], K_v=gains["Kv"], K_ff=gains["K_ff"], clip_value_pos=gains["clip_pos"],
                                                         clip_value_vel=gains["clip_vel"], clip_value_total=gains["clip_ang"]))
        assert torch.equal(ref.controllers.RaibertHeuristic(cfg).get_inference_policy("cpu")(obs), want)


def test_port_matches_reference_golden():
    g = np.load(GOLD)
    for tag, gains in (("a", PC.GAINS), ("b", PC.GAINS_B)):
        got = PC.raibert_policy(torch.from_numpy(g[f"obs_{tag}"]), **gains)
        assert_close(got, torch.from_numpy(g[f"act_{tag}"]), 1.0, f"raibert golden {tag}")
        assert torch.allclose(got.norm(dim=-1), torch.ones(got.shape[0]), atol=1e-5)      # unit quaternions


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 33, 4096, 1 << 20])
def test_kernel_matches_port_and_golden(n):
    from legged_gym_dev_b200.rom import RaibertHeuristic
    g = np.load(GOLD)
    for tag, gains in (("a", PC.GAINS), ("b", PC.GAINS_B)):
        obs = torch.from_numpy(g[f"obs_{tag}"]).cuda()
        got = RaibertHeuristic.raibert_policy(obs, gains["Kp"], gains["Kv"], gains["K_ff"], gains["clip_pos"], gains["clip_vel"], gains["clip_ang"])
        assert_close(got.cpu(), torch.from_numpy(g[f"act_{tag}"]), 1.0, f"raibert kernel vs reference golden {tag}")
    gains = PC.GAINS_B
    obs = PC.sample_obs(n, seed=n)
    cfg = SimpleNamespace(controller=SimpleNamespace(K_p=gains["Kp"], K_v=gains["Kv"], K_ff=gains["K_ff"], clip_value_pos=gains["clip_pos"],
                                                     clip_value_vel=gains["clip_vel"], clip_value_total=gains["clip_ang"]))
    policy = RaibertHeuristic(cfg).get_inference_policy("cuda")
    got = policy(obs.cuda())
    if n <= 4096:
        assert_close(got.cpu(), PC.raibert_policy(obs, **gains), 1.0, f"raibert kernel vs port, n={n}")
    assert torch.allclose(got.norm(dim=-1), torch.ones(n, device="cuda"), atol=1e-5)
    with pytest.raises(RuntimeError):
        policy(obs)                                   # CPU tensor: no fallback
    with pytest.raises(ValueError):
        policy(obs[:, :8].cuda())
