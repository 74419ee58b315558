"""Shared comparison helpers for parity tests.  TEST INFRASTRUCTURE.

The contract (BASELINE.json north_star): bit-exact for masks / flags / indices;
`|a-b| <= rtol*|b| + rtol*S` for fp32 quantities, S a per-quantity scale stated at the call site."""
import numpy as np
import torch

RTOL = 1e-5
# per-quantity scales S (atol = 1e-5 * S)
SCALE = dict(torques=80.0, obs=1.0, rew=1.0, rom=1.0, state=1.0, heights=1.0, sums=1.0, lstm=1.0, gae=1.0)


def _np(x):
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def assert_exact(a, b, what=""):
    a, b = _np(a), _np(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = a != b
    if a.dtype.kind == "f":
        bad &= ~(np.isnan(a) & np.isnan(b))
    assert not bad.any(), f"{what}: {int(bad.sum())} of {a.size} entries differ (first at {np.argwhere(bad)[0]})"


def max_err(a, b, scale=1.0):
    """max over entries of |a-b| / (|b| + S), the normalised error compared against RTOL."""
    a, b = _np(a).astype(np.float64), _np(b).astype(np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / (np.abs(b) + scale)))


def assert_close(a, b, scale=1.0, what="", rtol=RTOL):
    a_, b_ = _np(a), _np(b)
    assert a_.shape == b_.shape, f"{what}: shape {a_.shape} vs {b_.shape}"
    e = max_err(a_, b_, scale)
    assert e <= rtol, f"{what}: normalised error {e:.3e} > {rtol:.1e} (S={scale})"
    return e
