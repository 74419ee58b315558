"""Device-side mirror of deep_tube_learning/datasets.py (`construct_dataset`, `get_slice`, `sliding_window`) and of the
window construction in deep_tube_learning/evaluation/evaluate_tube_simple.py:28-46: the epoch logs written by
`CustomSim.collect_epoch` (b200gym_rom_rollout) stay in HBM and become the tube dataset without the pickle / numpy
round trip.  Same names, argument meaning and shapes as the reference; tensors are CUDA tensors instead of ndarrays."""
import glob
import pickle

import torch

from . import _lib


def _f32(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32")
    return t.contiguous()


def construct_dataset(epochs, data_folder=None):
    """datasets.py:11-57.  `epochs`: list of epoch dicts {'z','v','pz_x','done'} of device tensors (or a folder of
    epoch_*.pickle files, as the reference takes).  The reference sets `done_e[-1, :] = True` — the LAST ROBOT's row, not
    the last time step (datasets.py:24) — on every epoch before concatenating along the robot axis; mirrored as is."""
    if isinstance(epochs, str):
        data_folder, files, epochs = epochs, sorted(glob.glob(f"{epochs}/epoch_*.pickle")), []
        for f in files:
            with open(f, "rb") as fh:
                e = pickle.load(fh)
            epochs.append({k: torch.as_tensor(v).cuda() for k, v in e.items()})
    zs, vs, ds, ps = [], [], [], []
    for e in epochs:
        d = e["done"].clone()
        d[-1, :] = True
        zs.append(e["z"]), vs.append(e["v"]), ds.append(d), ps.append(e["pz_x"])
    z, v, done, pz_x = torch.cat(zs, 0), torch.cat(vs, 0), torch.cat(ds, 0), torch.cat(ps, 0)
    dataset = {"z": z, "pz_x": pz_x, "v": v, "z_p1": z[:, 1:, :].contiguous(), "pz_x_p1": pz_x[:, 1:, :].contiguous(), "done": done}
    if data_folder is not None:
        with open(f"{data_folder}/dataset.pickle", "wb") as f:
            pickle.dump({k: t.cpu().numpy() for k, t in dataset.items()}, f)
    return dataset


def sliding_window(data, N, dN, m):
    """datasets.py:69-71: [B,T,D] -> [B,T,N*D], one launch."""
    data = _f32(data, "data")
    B, T, D = data.shape
    out = torch.empty(B, T, N * D, device=data.device)
    _lib.check(_lib.lib().b200gym_sliding_window(_lib.ptr(data), _lib.ptr(out), B, T, D, N, dN, m, _lib.stream_ptr(data.device)),
               "sliding_window")
    return out


def get_slice(data, i, dN, m):
    """datasets.py:60-66: slice i of the window = columns [i*D, (i+1)*D) of sliding_window(data, i+1, dN, m)."""
    D = data.shape[-1]
    return sliding_window(data, i + 1, dN, m)[..., i * D:].contiguous()


def tube_error(z, pz_x, T=None):
    """w = ||pz_x - z|| over the first T samples (evaluate_tube_simple.py:28-31: the logs carry T+1 samples)."""
    z, pz_x = _f32(z, "z"), _f32(pz_x, "pz_x")
    B, T1, n = z.shape
    T = T1 - 1 if T is None else T
    w = torch.empty(B, T, device=z.device)
    _lib.check(_lib.lib().b200gym_tube_error(_lib.ptr(z), _lib.ptr(pz_x), _lib.ptr(w), B, T, T1, n, _lib.stream_ptr(z.device)), "tube_error")
    return w


def tube_windows(epoch_data, N, dN, recursive=False):
    """evaluate_tube_simple.py:28-46: (w, window_data) from one epoch's logs."""
    z, pz_x, v = epoch_data["z"], epoch_data["pz_x"], epoch_data["v"]
    w = tube_error(z, pz_x)
    z_no_pos = z[:, :-1, 2:]
    m = v.shape[-1]
    if recursive:
        data = torch.cat((w[:, :, None], z_no_pos, v), dim=-1)
        return w, sliding_window(data, N, dN, m)
    zv_slide = sliding_window(torch.cat((z_no_pos, v), dim=-1), N, dN, m)
    return w, torch.cat((w[:, :, None], zv_slide), dim=-1)
