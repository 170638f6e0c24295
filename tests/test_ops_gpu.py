"""GPU parity tests of the individual sm_100a kernels, called through the C ABI (include/ndit.h).
Reference for each op = plain fp32 PyTorch (or the oracle's helper functions) on the same inputs.
bf16 outputs: tolerance is stated per test in units of the output's bf16 resolution."""
import ctypes as C
import math

import pytest
import torch

from oracle import nextdit_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from lumina_t2x_b200 import _lib
    return _lib.load()


def ptr(t):
    return C.c_void_p(t.data_ptr())


def _rel_err(out, ref):
    return ((out.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-20)).item()


def _diag(out, ref, tile=(128, 64)):
    """Compact description of where a 2-D result is wrong (for debugging without a local GPU)."""
    d = (out.float() - ref.float()).abs()
    bad = d > (0.02 * ref.float().abs().max())
    M, N = d.shape
    rows = bad.any(dim=1).nonzero().flatten()
    cols = bad.any(dim=0).nonzero().flatten()
    msg = f"bad {int(bad.sum())}/{bad.numel()} max_abs {d.max().item():.4g} ref_absmax {ref.float().abs().max().item():.4g}"
    if rows.numel():
        msg += f" rows[{rows.min().item()}..{rows.max().item()}] n={rows.numel()} cols[{cols.min().item()}..{cols.max().item()}] n={cols.numel()}"
        msg += f" first_bad_rows {rows[:8].tolist()} first_bad_cols {cols[:8].tolist()}"
        r0, c0 = rows[0].item(), cols[0].item()
        msg += f" out[{r0},{c0}:{c0 + 4}]={out[r0, c0:c0 + 4].float().tolist()} ref={ref[r0, c0:c0 + 4].float().tolist()}"
    return msg


GEMM_SHAPES = [
    (128, 128, 64), (128, 256, 64), (128, 256, 256), (256, 512, 576), (512, 864, 576), (300, 1096, 584),
    (2048, 3456, 2304), (1024, 2304, 6144), (64, 1152, 256),
    # large enough for the CTA-pair (cta_group::2) kernel: M % 256 == 0, N % 256 == 0, >= 74 tiles of 256x256
    (4096, 2304, 2304), (8192, 2304, 6144), (2560, 2048, 192),
    # CTA-pair kernel with a ragged last M tile (Flag-DiT: 2 x 4160 tokens; the second CTA of the last pair is all padding)
    (8320, 3072, 3072), (4900, 2304, 256), (2700, 2048, 192),
    # CTA-pair kernel with a narrow last-N tile (fused q|k|v: 3456 = 13 x 256 + 128), down to a 32-wide one, also with a ragged M tile
    (8192, 3456, 2304), (2560, 2336, 192), (4100, 3488, 320),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_store(lib, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device="cuda", generator=g)).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    Cc = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.ndit_op_gemm(ptr(A), ptr(W), ptr(Cc), M, N, K, 0, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.ndit_last_error(None)
    ref = A.float() @ W.float().t()
    assert torch.isfinite(Cc.float()).all(), "non-finite / unwritten outputs: " + _diag(torch.nan_to_num(Cc.float(), nan=1e9), ref)
    # bf16 output of an fp32-accumulated product: <= 1 bf16 ulp (2^-8 relative) + accumulation-order noise
    err = (Cc.float() - ref).abs()
    tol = 2 ** -7 * ref.abs() + 1e-3 * ref.abs().max()
    assert (err <= tol).all(), _diag(Cc, ref)


def test_gemm_identity_layout(lib):
    """W = I picks out columns of A exactly: any swizzle/descriptor error shows up as a permutation."""
    M, N, K = 128, 256, 256
    A = torch.arange(M * K, device="cuda", dtype=torch.float32).reshape(M, K).remainder(251).to(torch.bfloat16)
    W = torch.eye(N, K, device="cuda", dtype=torch.bfloat16)
    Cc = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    rc = lib.ndit_op_gemm(ptr(A), ptr(W), ptr(Cc), M, N, K, 0, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.ndit_last_error(None)
    assert torch.equal(Cc, A[:, :N]), _diag(Cc, A[:, :N])


@pytest.mark.parametrize("M,F,K", [(128, 128, 64), (256, 512, 576), (1024, 6144, 2304), (200, 1536, 576),
                                   (8192, 6144, 2304), (8320, 8192, 3072), (2700, 2048, 384)])
def test_gemm_swiglu(lib, M, F, K):
    g = torch.Generator(device="cuda").manual_seed(M + F + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W1 = (torch.randn(F, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    W3 = (torch.randn(F, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    # block interleave: per 128 output features [w1 rows | w3 rows]
    W13 = torch.stack([W1.view(F // 128, 128, K), W3.view(F // 128, 128, K)], dim=1).reshape(2 * F, K).contiguous()
    Hh = torch.full((M, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.ndit_op_gemm(ptr(A), ptr(W13), ptr(Hh), M, 2 * F, K, 1, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.ndit_last_error(None)
    x1 = (A.float() @ W1.float().t()).to(torch.bfloat16)
    x3 = (A.float() @ W3.float().t()).to(torch.bfloat16)
    ref = (torch.nn.functional.silu(x1.float()).to(torch.bfloat16).float() * x3.float())
    assert torch.isfinite(Hh.float()).all()
    err = (Hh.float() - ref).abs()
    tol = 2 ** -6 * ref.abs() + 4e-3 * ref.abs().max()      # two stacked bf16 roundings of inputs that differ by 1 ulp
    assert (err <= tol).all(), _diag(Hh, ref)


def _ln_rope_ref(qkv, qw, qb, kw, kb, B, Hp, Wp, H, Hkv, hd, theta, lin):
    N = Hp * Wp
    q = qkv[:, : H * hd].float()
    k = qkv[:, H * hd: (H + Hkv) * hd].float()
    q = torch.nn.functional.layer_norm(q, (H * hd,), qw.float(), qb.float(), 1e-5)
    k = torch.nn.functional.layer_norm(k, (Hkv * hd,), kw.float(), kb.float(), 1e-5)
    # scale_watershed > timestep branch with scale_factor = lin reproduces (theta, lin)
    ang = O.rope_angles(hd, Hp, Wp, lin, 2.0, 0.0, theta=theta).to(qkv.device)
    q = O.apply_rope(q.view(B, N, H, hd), ang).reshape(B * N, H * hd)
    k = O.apply_rope(k.view(B, N, Hkv, hd), ang).reshape(B * N, Hkv * hd)
    out = qkv.clone()
    out[:, : H * hd] = q.to(torch.bfloat16)
    out[:, H * hd: (H + Hkv) * hd] = k.to(torch.bfloat16)
    return out


@pytest.mark.parametrize("B,Hp,Wp,H,Hkv,theta,lin", [(2, 8, 8, 8, 2, 10000.0, 1.0), (2, 16, 12, 32, 8, 20000.0, 1.0),
                                                       (1, 64, 64, 32, 8, 10000.0, 2.0)])
def test_ln_rope(lib, B, Hp, Wp, H, Hkv, theta, lin):
    hd = 72
    g = torch.Generator(device="cuda").manual_seed(11)
    W = (H + 2 * Hkv) * hd
    qkv = torch.randn(B * Hp * Wp, W, device="cuda", generator=g).to(torch.bfloat16)
    qw = (1 + 0.1 * torch.randn(H * hd, device="cuda", generator=g)).to(torch.bfloat16)
    qb = (0.1 * torch.randn(H * hd, device="cuda", generator=g)).to(torch.bfloat16)
    kw = (1 + 0.1 * torch.randn(Hkv * hd, device="cuda", generator=g)).to(torch.bfloat16)
    kb = (0.1 * torch.randn(Hkv * hd, device="cuda", generator=g)).to(torch.bfloat16)
    ref = _ln_rope_ref(qkv, qw, qb, kw, kb, B, Hp, Wp, H, Hkv, hd, theta, lin)
    got = qkv.clone()
    rc = lib.ndit_op_ln_rope(ptr(got), ptr(qw), ptr(qb), ptr(kw), ptr(kb), B, Hp, Wp, H, Hkv, hd, theta, lin, 0, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.ndit_last_error(None)
    assert torch.equal(got[:, (H + Hkv) * hd:], qkv[:, (H + Hkv) * hd:]), "v must be untouched"
    err = (got.float() - ref.float()).abs()
    # fp32 math rounded once to bf16: allow 1 bf16 ulp (sincos / reduction-order differences flip roundings)
    tol = 2 ** -7 * ref.float().abs() + 1e-3
    assert (err <= tol).all(), _diag(got, ref)


def _attn_ref(qkv, kvy, ymask, gate_tanh, B, N, T, H, Hkv, ss, sc):
    hd = 72
    rep = H // Hkv
    q = qkv[:, : H * hd].float().view(B, N, H, hd).permute(0, 2, 1, 3)
    k = qkv[:, H * hd: (H + Hkv) * hd].float().view(B, N, Hkv, hd).repeat_interleave(rep, 2).permute(0, 2, 1, 3)
    v = qkv[:, (H + Hkv) * hd:].float().view(B, N, Hkv, hd).repeat_interleave(rep, 2).permute(0, 2, 1, 3)
    ky = kvy[:, : Hkv * hd].float().view(B, T, Hkv, hd).repeat_interleave(rep, 2).permute(0, 2, 1, 3)
    vy = kvy[:, Hkv * hd:].float().view(B, T, Hkv, hd).repeat_interleave(rep, 2).permute(0, 2, 1, 3)
    a = torch.softmax(q @ k.transpose(-1, -2) * ss, -1) @ v
    s2 = (q @ ky.transpose(-1, -2) * sc).masked_fill(~ymask.bool()[:, None, None, :], float("-inf"))
    ay = torch.softmax(s2, -1) @ vy
    a = a.to(torch.bfloat16).float()
    ay = ay.to(torch.bfloat16).float()
    out = a + (ay * gate_tanh.view(1, H, 1, 1)).to(torch.bfloat16).float()
    return out.permute(0, 2, 1, 3).reshape(B * N, H * hd).to(torch.bfloat16)


ATTN_CASES = [
    # B, N, T, H, Hkv, valid caption tokens per batch row
    (1, 128, 8, 4, 1, [8]), (2, 256, 16, 8, 2, [16, 8]), (2, 384, 136, 8, 2, [136, 8]), (2, 64, 24, 8, 2, [24, 3]),
    (2, 1024, 128, 32, 8, [128, 8]), (1, 200, 40, 4, 4, [33]),
]


@pytest.mark.parametrize("B,N,T,H,Hkv,valid", ATTN_CASES)
@pytest.mark.parametrize("use_ref", [1, 0, 2, 3], ids=["refkernel", "tcgen05", "tcgen05_gen3_halfrow_tmemP", "tcgen05_gen1"])
def test_attention(lib, B, N, T, H, Hkv, valid, use_ref):
    hd = 72
    g = torch.Generator(device="cuda").manual_seed(N + T)
    qkv = torch.randn(B * N, (H + 2 * Hkv) * hd, device="cuda", generator=g).to(torch.bfloat16)
    kvy = torch.randn(B * T, 2 * Hkv * hd, device="cuda", generator=g).to(torch.bfloat16)
    ymask = torch.zeros(B, T, dtype=torch.uint8, device="cuda")
    for b, n in enumerate(valid):
        ymask[b, :n] = 1
    gate_tanh = torch.tanh(0.5 * torch.randn(H, device="cuda", generator=g)).to(torch.bfloat16).float()
    ss = math.sqrt(math.log(N, 64) / hd)
    sc = 1 / math.sqrt(hd)
    out = torch.full((B * N, H * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.ndit_op_attention(ptr(qkv), ptr(kvy), ptr(ymask), ptr(gate_tanh), ptr(out), B, N, T, H, Hkv, ss, sc, use_ref, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.ndit_last_error(None)
    ref = _attn_ref(qkv, kvy, ymask, gate_tanh, B, N, T, H, Hkv, ss, sc)
    assert torch.isfinite(out.float()).all(), "non-finite / unwritten: " + _diag(torch.nan_to_num(out.float(), nan=1e9), ref)
    err = (out.float() - ref.float()).abs()
    # The op rounds to bf16 three times (self part, gated caption part, sum) and feeds bf16 probabilities to P.V.  Measured
    # (tools/attn_error.py, B200): max|err| / max|ref| <= 9.6e-3 for both tcgen05 generations (P truncated to bf16), 7.8e-3 for the
    # CUDA-core kernel (P rounded to nearest) - i.e. two to three bf16 ulps of the largest outputs, dominated by the output
    # roundings, not by the truncation.  Tolerance: 1.2e-2 of the output scale.
    tol = 1.2e-2 * ref.float().abs().max()
    assert (err <= tol).all(), _diag(out, ref)


@pytest.mark.parametrize("M,rows,D,with_o", [(64, 32, 576, True), (64, 32, 576, False), (512, 256, 2304, True),
                                             (100, 50, 2304, True)])
def test_resid_rms_mod(lib, M, rows, D, with_o):
    g = torch.Generator(device="cuda").manual_seed(5)
    Bn = M // rows
    X = torch.randn(M, D, device="cuda", generator=g).to(torch.bfloat16)
    o = torch.randn(M, D, device="cuda", generator=g).to(torch.bfloat16)
    w_post = (1 + 0.1 * torch.randn(D, device="cuda", generator=g)).to(torch.bfloat16)
    w_pre = (1 + 0.1 * torch.randn(D, device="cuda", generator=g)).to(torch.bfloat16)
    tanh_g_b = torch.tanh(torch.randn(Bn, D, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
    onepls_b = (1 + 0.3 * torch.randn(Bn, D, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
    tanh_g, onepls = tanh_g_b.float(), onepls_b.float()
    p = O._Prec("bf16")
    Xr = X.float()
    if with_o:
        n = O.rms_norm(p, o.float(), w_post, 1e-5)
        Xr = p.r(Xr + p.r(tanh_g.repeat_interleave(rows, 0) * n))
    ur = p.r(O.rms_norm(p, Xr, w_pre, 1e-5) * onepls.repeat_interleave(rows, 0))
    Xg = X.clone()
    u = torch.full((M, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.ndit_op_resid_rms_mod(ptr(Xg), ptr(o) if with_o else None, ptr(w_post), ptr(tanh_g_b), ptr(w_pre), ptr(onepls_b),
                                   None, ptr(u), M, rows, D, 1e-5, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.ndit_last_error(None)
    # identical rounding points; reduction order may flip a bf16 rounding: <= 1 ulp on a few elements
    for got, ref, name in ((Xg, Xr, "X"), (u, ur, "u")):
        err = (got.float() - ref).abs()
        tol = 2 ** -7 * ref.abs() + 1e-6
        frac_exact = (err == 0).float().mean().item()
        assert (err <= tol).all(), name + " " + _diag(got, ref)
        assert frac_exact > 0.98, (name, frac_exact)


@pytest.mark.parametrize("M,D,E", [(512, 1152, 4), (300, 576, 8), (2048, 1536, 2)])
def test_moe_token_gate_routing(lib, M, D, E):
    """Token gate of the MoE FFN (models1.py:461-470) against an fp64 evaluation: every token whose top-2 set is decided by more than
    the bf16 resolution of the logits must be routed identically and its two weights must be the bf16 softmax of the bf16 logits;
    near-ties (a bf16 logit tie between the 2nd and 3rd expert, or between the two selected ones) may go either way and are counted."""
    g = torch.Generator(device="cuda").manual_seed(M + E)
    u = torch.randn(M, D, device="cuda", generator=g).to(torch.bfloat16)
    Wg = (torch.randn(E, D, device="cuda", generator=g) / D ** 0.5).to(torch.bfloat16)
    wtok = torch.full((M, E), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.ndit_op_moe_gate(ptr(u), ptr(Wg), ptr(wtok), M, D, E, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.ndit_last_error(None)
    w = wtok.float()
    assert torch.isfinite(w).all() and ((w != 0).sum(1) == 2).all() and (w >= 0).all()
    logits = (u.double() @ Wg.double().T)
    lb = logits.float().to(torch.bfloat16).float()                       # what the gate Linear returns under autocast
    top = torch.topk(lb, min(3, E), dim=1)
    clear = torch.ones(M, dtype=torch.bool, device="cuda")
    if E > 2:
        clear &= (top.values[:, 1] - top.values[:, 2]) > 2.0 ** -6 * top.values[:, 1].abs().clamp_min(1e-3)
    sel_ref = torch.zeros(M, E, dtype=torch.bool, device="cuda").scatter_(1, top.indices[:, :2], True)
    agree = ((w != 0) == sel_ref).all(1)
    assert agree[clear].all(), f"{(~agree[clear]).sum().item()} clearly decided tokens routed differently"
    assert clear.float().mean() > 0.9
    # weights: softmax over the two selected bf16 logits, rounded to bf16
    l0, l1 = top.values[:, 0], top.values[:, 1]
    e1 = torch.exp(l1 - l0)
    w0, w1 = (1 / (1 + e1)).to(torch.bfloat16).float(), (e1 / (1 + e1)).to(torch.bfloat16).float()
    got0 = w.gather(1, top.indices[:, :1]).squeeze(1)
    got1 = w.gather(1, top.indices[:, 1:2]).squeeze(1)
    ok = clear & agree
    assert ((got0 - w0).abs()[ok] <= 2.0 ** -8).all() and ((got1 - w1).abs()[ok] <= 2.0 ** -8).all()
