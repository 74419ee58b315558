"""RNG shim for running the UNMODIFIED reference under the counter-based stream of oracle/philox.py.

TEST INFRASTRUCTURE (container only).  The reference draws from torch's global generator
(SURVEY.md A.2/A.3); here torch.rand / rand_like / randint / randint_like are replaced, *only while a
draw context is active*, by functions that look up which envs the rows of the draw belong to and
return Philox numbers keyed (seed, env, event, site, column).  Contexts are published by thin
wrappers installed around the reference methods that draw (see oracle/ref_harness.py); the reference
code itself runs as written.
"""
import contextlib
import numpy as np
import torch

from . import philox

_orig = {name: getattr(torch, name) for name in ("rand", "rand_like", "randint", "randint_like")}


class _Ctx:
    def __init__(self, seed, env_ids, event, plan):
        self.seed = seed
        self.env_ids = env_ids      # np.int64 [n] or callable(call_index, history) -> np.int64 [n]
        self.event = event          # scalar or np.int64 [n] (or callable like env_ids)
        self.plan = plan            # list of (site, col0) per successive draw call
        self.calls = 0
        self.history = []           # returned tensors, for lazily-resolved contexts

    def next(self):
        if self.calls >= len(self.plan):
            raise RuntimeError(f"rng_shim: unexpected extra draw (plan {self.plan})")
        site, col0 = self.plan[self.calls]
        ids = self.env_ids(self.calls, self.history) if callable(self.env_ids) else self.env_ids
        ev = self.event(self.calls, self.history) if callable(self.event) else self.event
        self.calls += 1
        return site, col0, np.asarray(ids, dtype=np.int64).reshape(-1), ev


_stack = []


def _shape_of(size, kw):
    if "size" in kw:
        return tuple(kw["size"])
    if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
        return tuple(size[0])
    return tuple(int(s) for s in size)


def _draw_uniform(shape):
    ctx = _stack[-1]
    site, col0, ids, ev = ctx.next()
    n = shape[0] if len(shape) else 1
    ncols = int(np.prod(shape[1:])) if len(shape) > 1 else 1
    if n != len(ids):
        raise RuntimeError(f"rng_shim: draw of {shape} rows does not match {len(ids)} env ids (site {site})")
    u = philox.uniform01(ctx.seed, ids, ev, site, col0 + ncols)[:, col0:]
    t = torch.from_numpy(np.ascontiguousarray(u)).reshape(shape)
    ctx.history.append(t)
    return t


def _draw_int(shape, bound):
    ctx = _stack[-1]
    site, col0, ids, ev = ctx.next()
    n = shape[0] if len(shape) else 1
    ncols = int(np.prod(shape[1:])) if len(shape) > 1 else 1
    if n != len(ids):
        raise RuntimeError(f"rng_shim: int draw of {shape} rows does not match {len(ids)} env ids (site {site})")
    r = philox.randint(ctx.seed, ids, ev, site, col0 + ncols, bound)[:, col0:]
    t = torch.from_numpy(np.ascontiguousarray(r)).reshape(shape)
    ctx.history.append(t)
    return t


def _rand(*size, **kw):
    if not _stack:
        return _orig["rand"](*size, **kw)
    return _draw_uniform(_shape_of(size, kw))


def _rand_like(t, **kw):
    if not _stack:
        return _orig["rand_like"](t, **kw)
    return _draw_uniform(tuple(t.shape)).to(t.dtype)


def _randint(*args, **kw):
    if not _stack:
        return _orig["randint"](*args, **kw)
    if len(args) == 3:
        low, high, shape = args
    else:
        low, (high, shape) = 0, args
    assert low == 0
    return _draw_int(tuple(shape), int(high))


def _randint_like(t, *args, **kw):
    if not _stack:
        return _orig["randint_like"](t, *args, **kw)
    high = args[-1]
    return _draw_int(tuple(t.shape), int(high)).to(t.dtype)


def install():
    torch.rand, torch.rand_like, torch.randint, torch.randint_like = _rand, _rand_like, _randint, _randint_like


def uninstall():
    for k, v in _orig.items():
        setattr(torch, k, v)


@contextlib.contextmanager
def draws(seed, env_ids, event, plan):
    """Publish which envs / event / sites the draws inside the block belong to."""
    if torch.is_tensor(env_ids):
        env_ids = env_ids.detach().cpu().numpy()
    ctx = _Ctx(seed, env_ids, event, plan)
    _stack.append(ctx)
    try:
        yield ctx
    finally:
        _stack.pop()
