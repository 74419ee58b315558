"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of deep_tube_learning/datasets.py:60-71 (`get_slice`,
`sliding_window`), of `construct_dataset`'s array handling (:11-57, without the file I/O) and of the window construction in
deep_tube_learning/evaluation/evaluate_tube_simple.py:28-46.  Pinned: tests/test_oracle_cpu.py runs it against the reference's
own functions imported from /root/reference, tests/golden/datasets_*.npz hold reference outputs for the GPU box."""
import numpy as np


def get_slice(data, i, dN, m):
    """datasets.py:60-66."""
    dc = data.copy()
    slc = np.flip(np.arange(dc.shape[-2] - (i * dN) - 1, -1, step=-dN))
    start = dc[:, 0, :].reshape((dc.shape[0], 1, dc.shape[2])).copy()
    start[:, :, -m:] = 0
    return np.concatenate((np.repeat(start, dc.shape[-2] - len(slc), axis=-2), dc[:, slc, :]), axis=-2)


def sliding_window(data, N, dN, m):
    """datasets.py:69-71."""
    return np.concatenate([get_slice(data, i, dN, m) for i in range(N)], axis=-1)


def construct_dataset(epochs):
    """datasets.py:11-57 on in-memory epoch dicts (the reference reads them from epoch_*.pickle)."""
    z = v = pz_x = done = None
    for e in epochs:
        d = e["done"].copy()
        d[-1, :] = True                      # datasets.py:24 — last ROBOT's row
        if z is None:
            z, v, done, pz_x = e["z"], e["v"], d, e["pz_x"]
        else:
            z, v = np.concatenate((z, e["z"]), 0), np.concatenate((v, e["v"]), 0)
            done, pz_x = np.concatenate((done, d), 0), np.concatenate((pz_x, e["pz_x"]), 0)
    return {"z": z, "pz_x": pz_x, "v": v, "z_p1": z[:, 1:, :].copy(), "pz_x_p1": pz_x[:, 1:, :].copy(), "done": done}


def tube_windows(epoch_data, N, dN, recursive=False):
    """evaluate_tube_simple.py:28-46."""
    z, pz_x, v = epoch_data["z"][:, :-1, :], epoch_data["pz_x"][:, :-1, :], epoch_data["v"]
    w = np.linalg.norm(pz_x - z, axis=-1)
    z_no_pos = z[:, :, 2:]
    if recursive:
        data = np.concatenate((w[:, :, None], z_no_pos, v), axis=-1)
        return w, sliding_window(data, N, dN, v.shape[-1])
    zv_slide = sliding_window(np.concatenate((z_no_pos, v), axis=-1), N, dN, v.shape[-1])
    return w, np.concatenate((w[:, :, None], zv_slide), axis=-1)
