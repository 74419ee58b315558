"""Tube-dataset construction.  CPU: oracle/port_datasets.py against the reference's own functions (when
/root/reference is present) and against tests/golden/datasets_reference.npz (written from the reference).  GPU: the device path
(legged_gym_dev_b200/datasets.py -> b200gym_sliding_window / b200gym_tube_error) against the port and the goldens — pure data
movement, so every comparison is BIT-EXACT."""
import os

import numpy as np
import pytest
import torch

from oracle import port_datasets as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "datasets_reference.npz")


def _gold():
    return np.load(GOLD)


def test_port_matches_reference_goldens():
    g = _gold()
    for k, (B, T, D, N, dN, m) in enumerate(g["window_cases"]):
        data = g[f"w{k}_data"]
        assert np.array_equal(P.sliding_window(data, N, dN, m), g[f"w{k}_out"])
        assert np.array_equal(P.get_slice(data, N - 1, dN, m), g[f"w{k}_slice_last"])
    order = [int(g["cd_first_epoch"]), 1 - int(g["cd_first_epoch"])]
    epochs = [{k: g[f"cd_e{e}_{k}"] for k in ("z", "v", "pz_x", "done")} for e in order]
    ds = P.construct_dataset(epochs)
    for k, v in ds.items():
        assert np.array_equal(v, g[f"cd_out_{k}"]), k
    ep = {k: g[f"cd_e0_{k}"] for k in ("z", "v", "pz_x", "done")}
    for rec in (0, 1):
        w, win = P.tube_windows(ep, 4, 1, recursive=bool(rec))
        assert np.array_equal(w, g[f"tw{rec}_w"]) and np.array_equal(win, g[f"tw{rec}_win"])


def test_port_matches_reference_functions_live():
    from oracle import ref_harness as H
    if not H.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    from oracle.make_golden_datasets import reference_datasets_module
    D = reference_datasets_module()
    rng = np.random.default_rng(1)
    for (B, T, Dd, N, dN, m) in [(3, 13, 2, 4, 1, 1), (2, 30, 3, 10, 2, 2), (1, 5, 2, 7, 1, 0), (2, 8, 6, 2, 4, 3)]:
        data = rng.standard_normal((B, T, Dd)).astype(np.float32)
        assert np.array_equal(P.sliding_window(data, N, dN, m), D.sliding_window(data, N, dN, m))
        for i in range(N):
            assert np.array_equal(P.get_slice(data, i, dN, m), D.get_slice(data, i, dN, m))


@pytest.mark.gpu
def test_device_windows_match_goldens_and_port():
    from legged_gym_dev_b200 import datasets as DS
    g = _gold()
    for k, (B, T, D, N, dN, m) in enumerate(g["window_cases"]):
        data = torch.from_numpy(g[f"w{k}_data"]).cuda()
        assert np.array_equal(DS.sliding_window(data, int(N), int(dN), int(m)).cpu().numpy(), g[f"w{k}_out"])
        assert np.array_equal(DS.get_slice(data, int(N) - 1, int(dN), int(m)).cpu().numpy(), g[f"w{k}_slice_last"])
    rng = np.random.default_rng(3)
    for (B, T, Dd, N, dN, m) in [(64, 200, 3, 10, 1, 2), (7, 33, 5, 4, 3, 0), (1000, 50, 2, 6, 2, 2), (3, 4, 1, 9, 1, 1)]:
        data = rng.standard_normal((B, T, Dd)).astype(np.float32)
        got = DS.sliding_window(torch.from_numpy(data).cuda(), N, dN, m).cpu().numpy()
        assert np.array_equal(got, P.sliding_window(data, N, dN, m)), (B, T, Dd, N, dN, m)


@pytest.mark.gpu
def test_device_dataset_from_epoch_logs():
    from legged_gym_dev_b200 import datasets as DS
    g = _gold()
    order = [int(g["cd_first_epoch"]), 1 - int(g["cd_first_epoch"])]
    epochs = [{k: torch.from_numpy(g[f"cd_e{e}_{k}"]).cuda() for k in ("z", "v", "pz_x", "done")} for e in order]
    ds = DS.construct_dataset(epochs)
    for k, v in ds.items():
        assert np.array_equal(v.cpu().numpy(), g[f"cd_out_{k}"]), k
    ep = {k: torch.from_numpy(g[f"cd_e0_{k}"]).cuda() for k in ("z", "v", "pz_x", "done")}
    for rec in (0, 1):
        w, win = DS.tube_windows(ep, 4, 1, recursive=bool(rec))
        assert np.array_equal(w.cpu().numpy(), g[f"tw{rec}_w"]), "tube error"
        assert np.array_equal(win.cpu().numpy(), g[f"tw{rec}_win"]), "window data"


@pytest.mark.gpu
def test_rollout_logs_to_windows_end_to_end():
    """collect_epoch (persistent rollout kernel) -> tube windows, all on the device; checked against the port on the same logs."""
    from legged_gym_dev_b200 import configs, datasets as DS
    from legged_gym_dev_b200.rom import CustomSim
    env = CustomSim(configs.double_single_int_cfg(256, seed=0), device="cuda")
    obs = torch.zeros(256, 8, device="cuda")
    log = env.collect_epoch(obs, 40)
    w, win = DS.tube_windows(log, 10, 1, recursive=True)
    host = {k: v.cpu().numpy() for k, v in log.items()}
    w_ref, win_ref = P.tube_windows(host, 10, 1, recursive=True)
    assert np.array_equal(w.cpu().numpy(), w_ref) and np.array_equal(win.cpu().numpy(), win_ref)
    assert win.shape == (256, 40, 10 * 3)
