"""b200gym_gemm_f16 (csrc/gemm.cu): the three modes of the grouped tcgen05 GEMM against torch fp32 matmuls of the SAME
fp16-rounded operands (so the only differences are accumulation order and the fp16 rounding of an fp16 output)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(mode, a, b, out, m, n, k, aux=None, bias=None, flags=0, m_real=0, n_real=0, splits=1, scale=1.0, ldo=None):
    from legged_gym_dev_b200 import _lib
    p = _lib.GemmProblemPOD()
    p.a, p.b, p.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    p.aux = aux.data_ptr() if aux is not None else None
    p.bias = bias.data_ptr() if bias is not None else None
    p.mode, p.flags, p.m, p.n, p.k = mode, flags, m, n, k
    p.lda, p.ldb, p.ldo = a.stride(0), b.stride(0), out.stride(0) if ldo is None else ldo
    p.ldaux = aux.stride(0) if aux is not None else 0
    p.m_real, p.n_real, p.splits, p.scale = m_real, n_real, splits, scale
    return p


def _launch(problems):
    from legged_gym_dev_b200 import _lib
    arr = (_lib.GemmProblemPOD * len(problems))(*problems)
    _lib.check(_lib.lib().b200gym_gemm_f16(arr, len(problems), _lib.stream_ptr("cuda")), "gemm_f16")
    torch.cuda.synchronize()


def _report(name, got, want, tol):
    err = (got.float() - want.float()).abs()
    scale = want.float().abs().max().item() + 1e-6
    worst = err.max().item()
    if worst > tol * scale:
        idx = (err == err.max()).nonzero()[0].tolist()
        bad_rows = (err.max(dim=1).values > tol * scale).nonzero().flatten()[:16].tolist()
        bad_cols = (err.max(dim=0).values > tol * scale).nonzero().flatten()[:32].tolist()
        raise AssertionError(f"{name}: max |err| {worst:.4g} (scale {scale:.4g}) at {idx}: got {got[tuple(idx)].item():.6g} want "
                             f"{want[tuple(idx)].item():.6g}; first bad rows {bad_rows}, bad cols {bad_cols}")


def _h(*shape, gen, scale=1.0):
    return (scale * torch.randn(*shape, device="cuda", generator=gen)).half()


@pytest.mark.parametrize("rows,k,n,n_real,elu,f32", [(256, 48, 128, 128, True, False), (1000, 240, 512, 512, True, False),
                                                      (130, 128, 64, 64, True, False), (4096, 32, 16, 12, False, True),
                                                      (77, 512, 256, 256, True, False), (24576, 64, 32, 32, True, False)])
def test_forward_mode(rows, k, n, n_real, elu, f32):
    from legged_gym_dev_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(rows + k)
    x = _h(rows, k, gen=g)
    w = _h(n, k, gen=g, scale=k ** -0.5)
    w[n_real:] = 0
    bias = torch.randn(n_real, device="cuda", generator=g)
    out = torch.full((rows, n), 7.0, device="cuda", dtype=torch.float32 if f32 else torch.float16)
    _launch([_problem(_lib.GEMM_FWD, x, w, out, rows, n, k, bias=bias, flags=(1 if elu else 0) | (2 if f32 else 0), n_real=n_real)])
    z = x.float() @ w.float().t()
    z[:, :n_real] += bias
    want = torch.nn.functional.elu(z) if elu else z
    _report(f"FWD {rows}x{k}->{n}", out, want, 2e-3 if not f32 else 2e-5)


@pytest.mark.parametrize("rows,nl,kl", [(256, 16, 32), (1000, 128, 256), (130, 64, 128), (24576, 32, 64), (300, 256, 512)])
def test_dgrad_mode(rows, nl, kl):
    """dZ_{l-1} = (dZ_l . W_l) * elu'(H_{l-1}); W_l [nl, kl] is read MN-major from its plain row-major copy."""
    from legged_gym_dev_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(rows + nl)
    dz = _h(rows, nl, gen=g)
    w = _h(nl, kl, gen=g, scale=nl ** -0.5)
    h = torch.nn.functional.elu(torch.randn(rows, kl, device="cuda", generator=g)).half()
    out = torch.full((rows, kl), 7.0, device="cuda", dtype=torch.float16)
    _launch([_problem(_lib.GEMM_DGRAD, dz, w, out, rows, kl, nl, aux=h)])
    hf = h.float()
    want = (dz.float() @ w.float()) * torch.where(hf > 0, torch.ones_like(hf), hf + 1.0)
    _report(f"DGRAD {rows}x{nl}->{kl}", out, want, 2e-3)
    out2 = torch.full((rows, kl), 7.0, device="cuda", dtype=torch.float16)
    _launch([_problem(_lib.GEMM_DGRAD, dz, w, out2, rows, kl, nl)])
    _report(f"DGRAD (no aux) {rows}x{nl}->{kl}", out2, dz.float() @ w.float(), 2e-3)


@pytest.mark.parametrize("rows,nl,nl_real,kl,kl_real,splits", [(256, 128, 128, 48, 48, 1), (1000, 64, 64, 128, 128, 3), (24576, 16, 12, 32, 32, 16),
                                                                (5000, 512, 512, 240, 235, 7), (130, 16, 1, 128, 128, 2),
                                                                (3000, 256, 256, 512, 512, 5)])
def test_wgrad_mode(rows, nl, nl_real, kl, kl_real, splits):
    """dW_l += scale * dZ_l^T H_{l-1}, db_l += scale * colsum(dZ_l): both operands MN-major, K = batch rows, split-K with red.add."""
    from legged_gym_dev_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(rows + nl)
    dz = _h(rows, nl, gen=g)
    h = _h(rows, kl, gen=g)
    dw = torch.randn(nl_real, kl_real, device="cuda", generator=g)
    db = torch.randn(nl_real, device="cuda", generator=g)
    dw0, db0 = dw.clone(), db.clone()
    scale = 0.25
    _launch([_problem(_lib.GEMM_WGRAD, dz, h, dw, nl, kl, rows, bias=db, m_real=nl_real, n_real=kl_real, splits=splits, scale=scale)])
    want_w = dw0 + scale * (dz.float().t() @ h.float())[:nl_real, :kl_real]
    want_b = db0 + scale * dz.float().sum(0)[:nl_real]
    _report(f"WGRAD dW {rows}: {nl}x{kl}", dw, want_w, 1e-4)
    _report(f"WGRAD db {rows}: {nl}", db[None], want_b[None], 1e-4)


def test_grouped_launch_and_errors():
    """Several problems of different modes in ONE launch (actor + critic layers share launches in PPO.update)."""
    from legged_gym_dev_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(5)
    rows = 700
    x1, w1, x2, w2 = _h(rows, 48, gen=g), _h(128, 48, gen=g, scale=0.1), _h(rows, 64, gen=g), _h(32, 64, gen=g, scale=0.1)
    o1 = torch.zeros(rows, 128, device="cuda", dtype=torch.float16)
    o2 = torch.zeros(rows, 32, device="cuda", dtype=torch.float16)
    dz, w3 = _h(rows, 32, gen=g), _h(32, 64, gen=g, scale=0.1)
    o3 = torch.zeros(rows, 64, device="cuda", dtype=torch.float16)
    dw = torch.zeros(32, 64, device="cuda")
    _launch([_problem(_lib.GEMM_FWD, x1, w1, o1, rows, 128, 48, flags=1, n_real=128),
             _problem(_lib.GEMM_FWD, x2, w2, o2, rows, 32, 64, n_real=32),
             _problem(_lib.GEMM_DGRAD, dz, w3, o3, rows, 64, 32),
             _problem(_lib.GEMM_WGRAD, dz, x2, dw, 32, 64, rows, m_real=32, n_real=64, splits=4)])
    _report("group/FWD elu", o1, torch.nn.functional.elu(x1.float() @ w1.float().t()), 2e-3)
    _report("group/FWD lin", o2, x2.float() @ w2.float().t(), 2e-3)
    _report("group/DGRAD", o3, dz.float() @ w3.float(), 2e-3)
    _report("group/WGRAD", dw, dz.float().t() @ x2.float(), 1e-4)
    with pytest.raises(RuntimeError):   # n not a multiple of 16
        _launch([_problem(_lib.GEMM_FWD, x1, w1, o1, rows, 100, 48)])
    with pytest.raises(RuntimeError):   # unaligned leading dimension
        _launch([_problem(_lib.GEMM_FWD, torch.zeros(rows, 44, device="cuda", dtype=torch.float16), w1, o1, rows, 128, 32)])
