"""tests/golden/datasets_*.npz from the UNMODIFIED reference functions deep_tube_learning/datasets.py `get_slice`,
`sliding_window`, `construct_dataset` (run on pickles in a temp folder) and the window block of
evaluation/evaluate_tube_simple.py:28-46 (restated inline there, executed here with the reference's sliding_window).
Build-container only: python -m oracle.make_golden_datasets"""
import os
import pickle
import tempfile

import numpy as np

from oracle import ref_harness as H

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (B, T, D, N, dN, m)
WINDOW_CASES = [(5, 17, 3, 4, 1, 2), (3, 20, 4, 5, 2, 2), (2, 9, 2, 6, 3, 0), (4, 12, 5, 3, 1, 5), (1, 6, 1, 8, 1, 1)]


def reference_datasets_module():
    H.import_reference()           # installs the mocks (wandb, matplotlib, ...) and puts /root/reference on sys.path
    import sys
    from unittest.mock import MagicMock
    sys.modules.setdefault("wandb", MagicMock())
    import deep_tube_learning.datasets as D
    return D


def main():
    D = reference_datasets_module()
    rng = np.random.default_rng(7)
    out = {"window_cases": np.array(WINDOW_CASES)}
    for k, (B, T, Dd, N, dN, m) in enumerate(WINDOW_CASES):
        data = rng.standard_normal((B, T, Dd)).astype(np.float32)
        out[f"w{k}_data"] = data
        out[f"w{k}_out"] = D.sliding_window(data, N, dN, m)
        out[f"w{k}_slice_last"] = D.get_slice(data, N - 1, dN, m)
    # construct_dataset on two epochs of rollout-shaped logs
    epochs = []
    with tempfile.TemporaryDirectory() as tmp:
        for e in range(2):
            B, T, n = 6 + e, 11, 2
            ep = dict(z=rng.standard_normal((B, T + 1, n)).astype(np.float32), v=rng.standard_normal((B, T, n)).astype(np.float32),
                      pz_x=rng.standard_normal((B, T + 1, n)).astype(np.float32), done=rng.random((B, T)) < 0.1)
            epochs.append({k: v.copy() for k, v in ep.items()})
            with open(os.path.join(tmp, f"epoch_{e}.pickle"), "wb") as f:
                pickle.dump(ep, f)
        ds = D.construct_dataset(tmp)
    # the reference globs the folder: order of concatenation = glob order; record which epoch came first
    first_B = ds["z"].shape[0] and (6 if np.array_equal(ds["z"][:6], epochs[0]["z"]) else 7)
    out["cd_first_epoch"] = np.int64(0 if first_B == 6 else 1)
    for e, ep in enumerate(epochs):
        for k, v in ep.items():
            out[f"cd_e{e}_{k}"] = v
    for k, v in ds.items():
        out[f"cd_out_{k}"] = v
    # evaluate_tube_simple.py:28-46 with the reference's sliding_window
    ep = epochs[0]
    for rec in (0, 1):
        z, pz_x, v = ep["z"][:, :-1, :], ep["pz_x"][:, :-1, :], ep["v"]
        w = np.linalg.norm(pz_x - z, axis=-1)
        z_no_pos = z[:, :, 2:]
        if rec:
            win = D.sliding_window(np.concatenate((w[:, :, None], z_no_pos, v), axis=-1), 4, 1, v.shape[-1])
        else:
            win = np.concatenate((w[:, :, None], D.sliding_window(np.concatenate((z_no_pos, v), axis=-1), 4, 1, v.shape[-1])), axis=-1)
        out[f"tw{rec}_w"], out[f"tw{rec}_win"] = w, win
    path = os.path.join(GOLD, "datasets_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
