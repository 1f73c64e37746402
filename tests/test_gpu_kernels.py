"""Per-kernel parity on a real B200: every C-ABI kernel against a plain fp32 restatement of the same op.

bf16 tolerances: operands are rounded to bf16 before BOTH sides see them, the reference accumulates in
fp32/fp64, so the only differences are accumulation order and the final bf16 rounding of the output
(2^-9 relative) -> rtol 1e-2 on bf16 outputs, 2e-3 on fp32 outputs.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from weathermodel_b200 import ops  # noqa: E402


def _cmp(name, got, ref, rtol, atol):
    got = got.double().cpu()
    ref = ref.double().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite values in output"
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    if bad.any():
        idx = torch.nonzero(bad)
        first = tuple(idx[0].tolist())
        rows = sorted(set(idx[:, 0].tolist()))[:8]
        cols = sorted(set(idx[:, -1].tolist()))[:16]
        rel = (got - ref).norm() / (ref.norm() + 1e-30)
        raise AssertionError(
            f"{name}: {int(bad.sum())}/{bad.numel()} mismatches, max abs err {err.max():.4g}, rel fro {rel:.3g}, "
            f"first {first} got {got[first]:.5g} ref {ref[first]:.5g}; bad rows~{rows} cols~{cols}")


GEMM_DEFAULTS = {"gemm_two_cta": -1, "gemm_epi_warps": 0, "gemm_staged": -1}  # -1 / 0 = automatic choice  # library defaults (wm_gemm.cu)


def _bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


# ------------------------------------------------------------------------------------------ UMMA probe
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("N,K", [(64, 48), (128, 32), (48, 64), (128, 128), (48, 128), (32, 16), (16, 128)])
def test_umma_probe(a_mn, b_mn, N, K):
    a = _bf(128, K, seed=1)
    b = _bf(N, K, seed=2)
    d = ops.umma_probe(a, b, bool(a_mn), bool(b_mn))
    ref = a.float() @ b.float().t()
    _cmp(f"probe a_mn={a_mn} b_mn={b_mn} N={N} K={K}", d, ref, 1e-4, 1e-3)


# ------------------------------------------------------------------------------------------ gemm_tn
@pytest.mark.parametrize("M,N,K,tile_n", [
    (128, 128, 64, 128), (300, 200, 48, 0), (1000, 576, 576, 0), (256, 1728, 576, 192), (512, 2304, 576, 256),
    (640, 576, 2304, 0), (200, 600, 200, 0), (130, 48, 48, 0), (257, 64, 32, 0), (4096, 800, 200, 224),
    (1024, 1344, 336, 0), (365, 336, 1344, 0), (128, 32, 64, 32), (23360, 576, 576, 0),
])
def test_gemm_tn_plain(M, N, K, tile_n):
    a = _bf(M, K, seed=3)
    b = _bf(N, K, scale=K ** -0.5, seed=4)
    out = ops.gemm_tn(a, b, tile_n=tile_n)
    ref = a.float() @ b.float().t()
    _cmp(f"gemm_tn {M}x{N}x{K}", out, ref, 1e-2, 1e-2)


@pytest.mark.parametrize("M,N,K", [(1024, 576, 576), (4099, 1728, 576), (2048, 2304, 192), (1500, 192, 2304)])
def test_gemm_cta_pair_kernel_is_bit_identical(M, N, K):
    """The cta_group::2 kernel (256-row tiles over two SMs) must reproduce the single-CTA kernel exactly."""
    from weathermodel_b200._lib import lib

    a = _bf(M, K, seed=30)
    b = _bf(N, K, scale=K ** -0.5, seed=31)
    bias = torch.randn(N, device="cuda")
    res = _bf(M, N, seed=32)
    try:
        lib().wm_set_option(b"gemm_two_cta", 0)
        ref = ops.gemm_tn(a, b, bias=bias, relu=True, residual=res)
        lib().wm_set_option(b"gemm_two_cta", 1)
        got = ops.gemm_tn(a, b, bias=bias, relu=True, residual=res)
    finally:
        lib().wm_set_option(b"gemm_two_cta", GEMM_DEFAULTS["gemm_two_cta"])
    assert torch.equal(ref, got)
    _cmp("cta-pair gemm", got, torch.relu(a.float() @ b.float().t() + bias) + res.float(), 1e-2, 2e-2)


@pytest.mark.parametrize("M,N,K,tile_n", [
    (1024, 576, 576, 0), (4099, 1728, 576, 0), (2085, 2304, 576, 0), (1500, 576, 2304, 0), (1031, 200, 200, 0),
    (130, 48, 48, 0), (1111, 600, 200, 64), (2048, 1344, 336, 128), (365, 40, 64, 0),
])
def test_gemm_kernel_variants_are_bit_identical(M, N, K, tile_n):
    """8 vs 16 epilogue warps (gemm_epi_warps), single-CTA vs CTA-pair tiles (gemm_two_cta) and thread-per-row vs
    staged coalesced stores (gemm_staged) must agree bit for bit, incl. clipped rows / columns, dropout, the sign
    side channel, a residual and the fp32 output."""
    from weathermodel_b200._lib import lib

    a = _bf(M, K, seed=40)
    b = _bf(N, K, scale=K ** -0.5, seed=41)
    bias = torch.randn(N, device="cuda")
    res = _bf(M, N, seed=42)
    nbytes = lib().wm_gemm_sign_bits_bytes(M, N)

    def run():
        bits = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        o1 = ops.gemm_tn(a, b, bias=bias, relu=True, dropout_p=0.1, seed=11, stream_id=5, sign_bits_out=bits, tile_n=tile_n)
        o2 = ops.gemm_tn(a, b, bias=bias, residual=res, dropout_p=0.1, seed=12, stream_id=6, tile_n=tile_n)
        o3 = ops.gemm_tn(a, b, tile_n=tile_n)
        o4 = ops.gemm_tn(a, b, bias=bias, out_fp32=True, tile_n=tile_n)
        o5 = ops.gemm_tn(a, b, gate_bits=bits, gate_scale=1.0 / 0.9, tile_n=tile_n)
        return o1, bits, o2, o3, o4, o5

    try:
        lib().wm_set_option(b"gemm_two_cta", 0)
        lib().wm_set_option(b"gemm_epi_warps", 8)
        lib().wm_set_option(b"gemm_staged", 0)
        ref = run()
        for two, ew, stg in ops.GEMM_VARIANTS[1:]:
            lib().wm_set_option(b"gemm_two_cta", two)
            lib().wm_set_option(b"gemm_epi_warps", ew)
            lib().wm_set_option(b"gemm_staged", stg)
            got = run()
            for i, (r, g) in enumerate(zip(ref, got)):
                assert torch.equal(r, g), f"two_cta={two} epi_warps={ew} staged={stg}: output {i} differs"
    finally:
        for k, v in GEMM_DEFAULTS.items():
            lib().wm_set_option(k.encode(), v)
    _cmp("gemm variants", ref[3], a.float() @ b.float().t(), 1e-2, 1e-2)
    assert ops.device_error() == 0


def test_gemm_site_tuner_records_a_variant_and_keeps_results():
    """tune_gemm_sites times the four variants per call site and stores one; outputs before == after."""
    from weathermodel_b200._lib import lib

    M, D, FF = 8192, 192, 768
    a = _bf(M, D, seed=50)
    b = _bf(FF, D, scale=D ** -0.5, seed=51)
    bias = torch.randn(FF, device="cuda")
    before = ops.gemm_tn(a, b, bias=bias, relu=True)
    chosen = ops.tune_gemm_sites(M, D, FF, 0.1, a.device)
    assert set(chosen) == {"qkv", "out_proj", "linear1", "linear2", "linear2_dgrad", "linear1_dgrad", "out_proj_dgrad",
                           "qkv_dgrad"}
    assert all(v in ops.GEMM_VARIANTS for v in chosen.values())
    assert ops.tune_gemm_sites(M, D, FF, 0.1, a.device) == {}  # once per shape and process
    assert torch.equal(before, ops.gemm_tn(a, b, bias=bias, relu=True))
    assert lib().wm_gemm_set_variant(M, FF, D, None, 0, 2, 8, 0) != 0  # rejects unknown variants
    assert ops.device_error() == 0


def test_gemm_tn_epilogues():
    M, N, K = 777, 576, 576
    a = _bf(M, K, seed=5)
    b = _bf(N, K, scale=K ** -0.5, seed=6)
    bias = torch.randn(N, device="cuda")
    res = _bf(M, N, seed=7)
    acc = a.float() @ b.float().t()
    _cmp("bias", ops.gemm_tn(a, b, bias=bias), acc + bias, 1e-2, 1e-2)
    _cmp("bias+relu", ops.gemm_tn(a, b, bias=bias, relu=True), torch.relu(acc + bias), 1e-2, 1e-2)
    _cmp("bias+res", ops.gemm_tn(a, b, bias=bias, residual=res), acc + bias + res.float(), 1e-2, 2e-2)
    _cmp("fp32 out", ops.gemm_tn(a, b, bias=bias, out_fp32=True), acc + bias, 2e-3, 2e-3)
    gate = _bf(M, N, seed=8)
    _cmp("gate", ops.gemm_tn(a, b, gate=gate, gate_scale=1.25),
         torch.where(gate.float() > 0, acc * 1.25, torch.zeros_like(acc)), 1e-2, 1e-2)
    # sign side channel: the bits a ReLU GEMM writes gate a later GEMM exactly like the activation itself would
    bits = ops.gemm_sign_bits(M, N, "cuda")
    h = ops.gemm_tn(a, b, bias=bias, relu=True, sign_bits_out=bits)
    g1 = ops.gemm_tn(a, b, gate=h, gate_scale=1.25)
    g2 = ops.gemm_tn(a, b, gate_bits=bits, gate_scale=1.25)
    assert torch.equal(g1, g2), "gate through sign bits differs from gate through the bf16 activation"
    assert ((h > 0) == (g2 != 0)).float().mean().item() > 0.999
    # dropout: kept fraction ~ 1-p_eff, kept values scaled by 1/(1-p_eff) with p_eff = round(32768 p)/32768 (0.100006 for
    # the reference's p = 0.1), identical across calls
    p = 0.1
    d1 = ops.gemm_tn(a, b, bias=bias, dropout_p=p, seed=11, stream_id=5, out_fp32=True)
    d2 = ops.gemm_tn(a, b, bias=bias, dropout_p=p, seed=11, stream_id=5, out_fp32=True)
    d3 = ops.gemm_tn(a, b, bias=bias, dropout_p=p, seed=11, stream_id=6, out_fp32=True)
    assert torch.equal(d1, d2), "dropout mask is not a pure function of (seed, stream, index)"
    kept = d1 != 0
    frac = kept.float().mean().item()
    t15 = (int(p * 65536 + 0.5) + 1) >> 1
    assert abs(t15 / 32768 - p) < 1e-5, "effective dropout probability must be the reference's to 1e-5"
    assert abs(frac - (1 - t15 / 32768)) < 5e-3, f"kept fraction {frac}"
    assert (kept != (d3 != 0)).float().mean().item() > 0.1, "different stream ids give the same mask"
    scale = 32768.0 / (32768 - t15)
    _cmp("dropout kept values", d1[kept], ((acc + bias) * scale)[kept], 2e-3, 2e-3)


@pytest.mark.parametrize("M,N,K", [(777, 576, 576), (1030, 2304, 576), (515, 576, 2304)])
def test_gemm_straight_line_epilogues_match_the_generic_body(M, N, K):
    """bf16 outputs of the nine epilogue flag sets of the training / evaluation schedules run through straight-line
    packed-math bodies (csrc/wm_gemm.cu epi_fast16); fp32 outputs run through the generic run-time-flag body. Same
    masks (zero pattern bit for bit), same values up to the bf16 rounding of the output."""
    a = _bf(M, K, seed=70)
    b = _bf(N, K, scale=K ** -0.5, seed=71)
    bias = torch.randn(N, device="cuda")
    res = _bf(M, N, seed=72)
    drop = dict(dropout_p=0.1, seed=21, stream_id=9)
    bits = ops.gemm_sign_bits(M, N, "cuda")
    bits32 = ops.gemm_sign_bits(M, N, "cuda")
    cases = {
        "plain": {}, "bias": dict(bias=bias), "bias+relu": dict(bias=bias, relu=True),
        "bias+relu+bits": dict(bias=bias, relu=True, sign_bits_out=bits),
        "bias+relu+drop+bits": dict(bias=bias, relu=True, sign_bits_out=bits, **drop),
        "bias+drop+res": dict(bias=bias, residual=res, **drop), "bias+res": dict(bias=bias, residual=res),
        "res": dict(residual=res),
    }
    for name, kw in cases.items():
        kw32 = dict(kw)
        if "sign_bits_out" in kw32:
            kw32["sign_bits_out"] = bits32
        fast = ops.gemm_tn(a, b, **kw)
        slow = ops.gemm_tn(a, b, out_fp32=True, **kw32)
        assert fast.dtype == torch.bfloat16 and slow.dtype == torch.float32
        if "residual" not in kw:  # (with a residual a dropped element still carries the residual)
            differ = (fast == 0) != (slow.to(torch.bfloat16) == 0)  # (only where acc + bias cancels to fp32 round-off)
            assert bool((slow.abs()[differ] < 1e-5).all()) and int(differ.sum()) <= 4, f"{name}: zero pattern differs"
        err = (fast.float() - slow).abs()
        bound = slow.abs() * 2.0 ** -8 + 1e-6  # half an ulp of bf16 plus the fma-vs-(add, mul) difference
        assert bool((err <= bound).all()), f"{name}: max excess {(err - bound).max().item():.3e}"
        if "sign_bits_out" in kw:  # the sign side channel: bits of the bf16 output the next GEMM will see
            g_fast = ops.gemm_tn(res, torch.eye(N, device="cuda", dtype=torch.bfloat16), gate_bits=bits, gate_scale=1.0)
            assert torch.equal(g_fast != 0, (fast > 0) & (res != 0)), f"{name}: sign bits do not describe the output"
    # gate bits (dgrad through dropout(relu(.))): straight-line body against the generic one
    h = ops.gemm_tn(a, b, bias=bias, relu=True, sign_bits_out=bits, **drop)
    g_fast = ops.gemm_tn(a, b, gate_bits=bits, gate_scale=1.0 / 0.9)
    g_slow = ops.gemm_tn(a, b, gate_bits=bits, gate_scale=1.0 / 0.9, out_fp32=True)
    assert torch.equal(g_fast != 0, (h > 0) & (g_slow.to(torch.bfloat16) != 0))
    assert bool(((g_fast.float() - g_slow).abs() <= g_slow.abs() * 2.0 ** -8 + 1e-6).all())
    assert ops.device_error() == 0


# ------------------------------------------------------------------------------------------ gemm_wgrad
@pytest.mark.parametrize("Mtok,Nout,Kout", [
    (64, 128, 64), (1000, 576, 576), (777, 64, 576), (3000, 200, 600), (4097, 1728, 576), (2048, 576, 2304),
    (1500, 48, 192), (5000, 576, 64), (365 * 8, 336, 1344), (46720, 200, 200),
    # big enough for two A^T tiles per CTA (wgrad_pick_mh): even / odd tile counts, a clipped last tile, narrow outputs
    (16384 + 77, 1728, 576), (20000, 2304, 576), (17000, 640, 192), (16500, 200, 128), (18000, 576, 576),
])
def test_gemm_wgrad(Mtok, Nout, Kout):
    a = _bf(Mtok, Nout, seed=9)
    b = _bf(Mtok, Kout, seed=10)
    out = ops.gemm_wgrad(a, b)
    ref = a.double().t() @ b.double()
    _cmp(f"wgrad {Mtok}x{Nout}x{Kout}", out, ref, 2e-3, 2e-3 * math.sqrt(Mtok))
    out2, db = ops.gemm_wgrad(a, b, want_bias_grad=True)
    assert torch.equal(out, out2), "split-K wgrad is not deterministic"
    _cmp(f"wgrad bias {Mtok}x{Nout}", db, a.double().sum(0), 2e-3, 2e-3 * math.sqrt(Mtok))
    # one and two A^T tiles per CTA (forced): the same product under another token-split plan
    from weathermodel_b200._lib import lib
    outs = {}
    try:
        for mh in (1, 2):
            lib().wm_set_option(b"wgrad_mh", mh)
            outs[mh] = ops.gemm_wgrad(a, b, want_bias_grad=True)
    finally:
        lib().wm_set_option(b"wgrad_mh", 0)
    _cmp("wgrad: two tiles per CTA vs one", outs[2][0], outs[1][0], 1e-4, 1e-4 * math.sqrt(Mtok))
    _cmp("wgrad bias: two tiles per CTA vs one", outs[2][1], outs[1][1], 1e-4, 1e-4 * math.sqrt(Mtok))
    assert ops.device_error() == 0


# ------------------------------------------------------------------------------------------ masks
@pytest.mark.parametrize("n,p", [(64, 0.15), (4096, 0.30), (1, 0.15), (37, 0.5)])
def test_mask_bert_bit_exact(n, p):
    torch.manual_seed(1234)
    ref_rand = torch.rand(n, 365, 31, device="cuda")
    ref_next = torch.rand(5, device="cuda")
    torch.manual_seed(1234)
    mask, rnd = ops.mask_bert(365, 31, n, p, return_rand=True)
    got_next = torch.rand(5, device="cuda")
    assert torch.equal(rnd, ref_rand), f"uniforms differ in {(rnd != ref_rand).sum().item()} places"
    assert torch.equal(mask, ref_rand < p)
    assert torch.equal(ref_next, got_next), "generator offset bookkeeping differs from torch.rand"


@pytest.mark.parametrize("n,k", [(64, 10), (4096, 1), (5000, 25), (3, 31), (8, 0)])
def test_mask_former_bit_exact(n, k):
    torch.manual_seed(1234)
    rv = torch.rand(n, 31, device="cuda")
    ref = (torch.argsort(rv, dim=-1) < k).unsqueeze(1).expand(-1, 365, -1)
    ref_next = torch.rand(3, device="cuda")
    torch.manual_seed(1234)
    got = ops.mask_former(365, 31, n, k)
    got_next = torch.rand(3, device="cuda")
    assert got.shape == ref.shape and got.stride(1) == 0
    assert torch.equal(got, ref)
    assert torch.equal(ref_next, got_next)


# ------------------------------------------------------------------------------------------ embedding
@pytest.mark.parametrize("B,S,D", [(4, 365, 48), (3, 364, 200), (2, 365, 576), (1, 7, 336),
                                   (421, 365, 48)])  # the last one is large enough for the 256-token CTAs
def test_embed_fwd(B, S, D):
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.randn(B, S, 31, device="cuda", generator=g)
    mask = torch.rand(B, S, 31, device="cuda", generator=g) < 0.3
    year = 1984.0 + torch.rand(B, S, device="cuda", generator=g) * 14
    coords = torch.stack([torch.rand(B, device="cuda", generator=g) * 120 - 60,
                          torch.rand(B, device="cuda", generator=g) * 360 - 180], 1)
    w_in = torch.randn(D, 34, device="cuda", generator=g) * 0.2
    b_in = torch.randn(D, device="cuda", generator=g) * 0.1
    pe = torch.randn(365, D, device="cuda", generator=g)
    out, xin = ops.embed_fwd(w, mask, year, coords, w_in, b_in, pe, want_xin=True)
    x = torch.cat([w * (~mask), ((year - 1970) / 100.0).unsqueeze(2),
                   (coords[:, 0] / 360).view(B, 1, 1).expand(B, S, 1), (coords[:, 1] / 180).view(B, 1, 1).expand(B, S, 1)], 2)
    ref = (x.double() @ w_in.double().t() + b_in.double() + pe[:S].double().unsqueeze(0)).view(B * S, D)
    _cmp("embed", out, ref, 1e-2, 1e-2)
    _cmp("xin", xin[:, :34], x.view(B * S, 34), 1e-2, 1e-3)
    assert (xin[:, 34:] == 0).all()
    # stride-0 (expanded) feature mask
    fm = (torch.rand(B, 31, device="cuda", generator=g) < 0.3).unsqueeze(1).expand(-1, S, -1)
    out2 = ops.embed_fwd(w, fm, year, coords, w_in, b_in, pe)
    x2 = torch.cat([w * (~fm), x[:, :, 31:]], 2)
    ref2 = (x2.double() @ w_in.double().t() + b_in.double() + pe[:S].double().unsqueeze(0)).view(B * S, D)
    _cmp("embed expand-mask", out2, ref2, 1e-2, 1e-2)


# ------------------------------------------------------------------------------------------ layernorm
@pytest.mark.parametrize("M,D", [(1000, 48), (777, 200), (4099, 336), (2920, 576), (46727, 576), (7, 48)])
def test_layernorm_fwd_bwd(M, D):
    x = _bf(M, D, scale=2.0, seed=12)
    dy = _bf(M, D, seed=13)
    gamma = torch.randn(D, device="cuda") * 0.5 + 1.0
    beta = torch.randn(D, device="cuda") * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = beta.double().requires_grad_(True)
    yref = torch.nn.functional.layer_norm(xd, (D,), gd, bd, 1e-5)
    _cmp("ln y", y, yref.detach(), 1e-2, 1e-2)
    _cmp("ln mean", mean, xd.detach().mean(1), 1e-5, 1e-5)
    _cmp("ln rstd", rstd, 1.0 / torch.sqrt(xd.detach().var(1, unbiased=False) + 1e-5), 1e-4, 1e-5)
    yref.backward(dy.double())
    dx, dxd, dg, db, dbias = ops.layernorm_bwd(dy, x, gamma, mean, rstd)
    assert dxd is None
    _cmp("ln dx", dx, xd.grad, 1e-2, 1e-2)
    _cmp("ln dgamma", dg, gd.grad, 2e-3, 2e-3 * math.sqrt(M))
    _cmp("ln dbeta", db, bd.grad, 2e-3, 2e-3 * math.sqrt(M))
    _cmp("ln dbias", dbias, dx.double().sum(0), 1e-4, 1e-3)
    # dropout variant: dx_dropped = mask * dx * scale with the SAME mask gemm_tn draws for (seed, stream)
    p = 0.1
    dx2, dxd2, _, _, dbias2 = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dropout_p=p, seed=5, stream_id=9)
    assert torch.equal(dx2, dx)
    if D % 16 == 0:
        ones_a = torch.zeros(M, 16, device="cuda", dtype=torch.bfloat16)
        ones_b = torch.zeros(D, 16, device="cuda", dtype=torch.bfloat16)
        probe = ops.gemm_tn(ones_a, ones_b, bias=torch.ones(D, device="cuda"), dropout_p=p, seed=5, stream_id=9,
                            out_fp32=True)
        keep = probe != 0
        scale = 32768.0 / (32768 - ((int(p * 65536 + 0.5) + 1) >> 1))
        _cmp("ln dx_dropped", dxd2, torch.where(keep, dx.float() * scale, torch.zeros_like(probe)), 1e-2, 1e-3)
    _cmp("ln dbias dropped", dbias2, dxd2.double().sum(0), 1e-4, 1e-3)
    # the encoder's instantiation (no bias gradient, 15 row warps per SM) must give the same results bit for bit
    dx3, dxd3, dg3, db3, none = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dropout_p=p, seed=5, stream_id=9,
                                                  want_bias_grad=False)
    assert none is None and torch.equal(dx3, dx2) and torch.equal(dxd3, dxd2)
    _cmp("ln dgamma (lean)", dg3, gd.grad, 2e-3, 2e-3 * math.sqrt(M))
    _cmp("ln dbeta (lean)", db3, bd.grad, 2e-3, 2e-3 * math.sqrt(M))


def test_colsum():
    for M, N in [(1000, 600), (46720, 2304), (5, 48), (777, 1728)]:
        x = _bf(M, N, seed=14)
        _cmp(f"colsum {M}x{N}", ops.colsum(x), x.double().sum(0), 1e-4, 1e-3 * math.sqrt(M))


# ------------------------------------------------------------------------------------------ losses
def test_loss_bert():
    M, F = 8 * 365, 31
    g = torch.Generator(device="cuda").manual_seed(3)
    y = torch.randn(M, 32, device="cuda", generator=g)
    w = torch.randn(M, F, device="cuda", generator=g)
    mask = torch.rand(M, F, device="cuda", generator=g) < 0.15
    out, dy = ops.loss_bert(y, w, mask)
    yd = y[:, :F].double().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(w.double()[mask], yd[mask])
    ref.backward()
    assert abs(out[0].item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert out[1].item() == mask.sum().item()
    _cmp("bert dy", dy[:, :F], yd.grad, 1e-2, 1e-9)
    assert (dy[:, F:] == 0).all()
    # the autograd split: value now, gradient later, scaled by a DEVICE scalar (the upstream gradient of the loss)
    out2, scratch = ops.loss_bert_value(y, w, mask)
    assert torch.equal(out2, out)
    assert torch.equal(ops.loss_bert_grad(y, w, mask, scratch, None, ld_grad=32), dy)
    gs = torch.tensor(2.5, device="cuda")
    _cmp("bert dy * 2.5", ops.loss_bert_grad(y, w, mask, scratch, gs, ld_grad=32)[:, :F], 2.5 * yd.grad, 1e-2, 1e-9)


@pytest.mark.parametrize("expand", [False, True])
def test_loss_former(expand):
    B, S, F, beta = 8, 365, 31, 0.5
    g = torch.Generator(device="cuda").manual_seed(4)
    y = torch.randn(B * S, 64, device="cuda", generator=g)
    y[:, F:2 * F] = y[:, F:2 * F] * 3 - 2  # exercise both clamp sides of exp(logvar)
    w = torch.randn(B, S, F, device="cuda", generator=g)
    if expand:
        mask = (torch.rand(B, F, device="cuda", generator=g) < 0.3).unsqueeze(1).expand(-1, S, -1)
    else:
        mask = torch.rand(B, S, F, device="cuda", generator=g) < 0.3
    out, dy, mu, var = ops.loss_former(y, w, mask, beta, want_mu_var=True)
    yd = y.double().requires_grad_(True)
    mu_r = yd[:, :F].view(B, S, F)
    var_r = torch.clamp(torch.exp(yd[:, F:2 * F].view(B, S, F)), min=1e-6, max=1)
    wd = w.double()
    n_bar = mask.sum(dim=(1, 2)).double().mean()
    ll = (-0.5 * torch.log(2 * torch.pi * var_r) - 0.5 * (wd - mu_r) ** 2 / var_r) * mask
    recon = (-ll.sum(dim=(1, 2)) / n_bar).mean()
    klpd = 0.5 * (torch.log(1.0 / var_r) + var_r + mu_r ** 2 - 1.0) * mask
    kl = beta * klpd.sum(dim=(1, 2)).mean() / n_bar
    total = recon + kl
    total.backward()
    for got, ref, nm in [(out[0], total, "total"), (out[1], recon, "recon"), (out[2], kl, "kl")]:
        assert abs(got.item() - ref.item()) <= 2e-5 * abs(ref.item()) + 1e-7, f"{nm}: {got.item()} vs {ref.item()}"
    _cmp("former mu", mu, mu_r.detach(), 0, 0)
    _cmp("former var", var, var_r.detach(), 1e-6, 1e-12)
    _cmp("former dy", dy[:, :2 * F], yd.grad[:, :2 * F], 1e-2, 1e-9)
    assert (dy[:, 2 * F:] == 0).all()
    out2, scratch = ops.loss_former_value(y, w, mask, beta)
    assert torch.equal(out2, out)
    assert torch.equal(ops.loss_former_grad(y, w, mask, beta, scratch, None, ld_grad=64), dy)
    gs = torch.tensor(-0.75, device="cuda")
    _cmp("former dy * -0.75", ops.loss_former_grad(y, w, mask, beta, scratch, gs, ld_grad=64)[:, :2 * F],
         -0.75 * yd.grad[:, :2 * F], 1e-2, 1e-9)


def test_adam_matches_torch():
    n = 100003
    g = torch.Generator(device="cuda").manual_seed(5)
    p0 = torch.randn(n, device="cuda", generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=5e-4)
    p = p0.clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 6):
        grad = torch.randn(n, device="cuda", generator=g) * (0.1 if step % 2 else 3.0)
        ref_p.grad = grad.clone()
        opt.step()
        ops.adam_fused(p, grad, m, v, step, 5e-4, shadow=shadow)
    _cmp("adam p", p, ref_p.detach(), 1e-5, 1e-6)
    _cmp("adam m", m, opt.state[ref_p]["exp_avg"], 1e-5, 2e-6)
    _cmp("adam v", v, opt.state[ref_p]["exp_avg_sq"], 1e-5, 1e-6)
    assert torch.equal(shadow, p.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, B, S, H, dh):
    D = H * dh
    x = qkv.double().view(B, S, 3, H, dh).permute(2, 0, 3, 1, 4)  # [3,B,H,S,dh]
    q, k, v = x[0], x[1], x[2]
    s = q @ k.transpose(-1, -2) / math.sqrt(dh)
    p = torch.softmax(s, dim=-1)
    o = p @ v
    lse = torch.logsumexp(s, dim=-1)
    return o.permute(0, 2, 1, 3).reshape(B * S, D), lse.reshape(B * H, S)


@pytest.mark.parametrize("B,S,H,dh", [(2, 365, 4, 12), (2, 365, 10, 20), (1, 364, 12, 28), (3, 365, 16, 36),
                                      (2, 128, 2, 36), (1, 70, 3, 16), (1, 384, 2, 48),
                                      # ragged tiles / every chunk count: S around the 64 / 96 / 128-row boundaries
                                      (1, 5, 2, 12), (2, 64, 2, 24), (1, 97, 4, 32), (1, 129, 2, 40), (1, 257, 2, 44),
                                      (3, 300, 6, 20), (5, 193, 2, 36)])
def test_attention_fwd_bwd(B, S, H, dh):
    D = H * dh
    qkv = _bf(B * S, 3 * D, seed=20)
    dctx = _bf(B * S, D, seed=21)
    ctx, lse = ops.attn_fwd(qkv, B, S, H, dh)
    qd = qkv.double().requires_grad_(True)
    oref, lref = _attn_ref(qd, B, S, H, dh)
    _cmp("attn ctx", ctx, oref.detach(), 2e-2, 2e-2)
    _cmp("attn lse", lse, lref.detach(), 1e-3, 1e-3)
    oref.backward(dctx.double())
    dqkv = ops.attn_bwd(qkv, ctx, dctx, lse, B, S, H, dh)
    ref = qd.grad
    rel = ((dqkv.double() - ref).norm() / ref.norm()).item()
    for nm, sl in [("dQ", slice(0, D)), ("dK", slice(D, 2 * D)), ("dV", slice(2 * D, 3 * D))]:
        r = ((dqkv[:, sl].double() - ref[:, sl]).norm() / ref[:, sl].norm()).item()
        assert r < 2e-2, f"attn {nm} rel fro err {r:.4g} (all {rel:.4g})"
    _cmp("attn dqkv", dqkv, ref, 5e-2, 5e-2)


def _decode_keep_bits(words, B, S, H):
    """keep[b, h, q, k] from the word buffer attn_fwd writes: one uint32 per (query row, 32-key slice), laid out
    [item][key tile (3)][query block of 64 (6)][slice in tile (4)][query row in block (64)]; key k of the slice sits
    at bit (k >> 1) + 16 * (k & 1) (csrc/wm_attn.cu: attn_keep_bit)."""
    w = words.view(B * H, 3, 6, 4, 64).to(torch.int64) & 0xFFFFFFFF
    q = torch.arange(S, device=words.device)
    k = torch.arange(S, device=words.device)
    sel = w[:, (k >> 7)[None, :], (q >> 6)[:, None], ((k >> 5) & 3)[None, :], (q & 63)[:, None]]  # [BH, S(q), S(k)]
    kk = k & 31
    keep = (sel >> ((kk >> 1) + 16 * (kk & 1))[None, None, :]) & 1
    return keep.bool().view(B, H, S, S)


@pytest.mark.parametrize("B,S,H,dh", [(2, 365, 4, 12), (1, 150, 2, 36), (2, 365, 16, 36)])
def test_attention_dropout_matches_torch_with_the_same_mask(B, S, H, dh):
    """Forward and backward with dropout against autograd of softmax(QK^T/sqrt(dh)) * keep / (1 - p_eff) @ V, with
    `keep` decoded from the bits the forward kernel hands to the backward kernel (p_eff = round(32768 p) / 32768)."""
    D = H * dh
    p = 0.1
    qkv = _bf(B * S, 3 * D, seed=30)
    dctx = _bf(B * S, D, seed=31)
    ctx, lse = ops.attn_fwd(qkv, B, S, H, dh, dropout_p=p, seed=5, stream_id=2)
    keep = _decode_keep_bits(ctx.drop_words, B, S, H)
    t15 = (int(p * 65536 + 0.5) + 1) >> 1
    frac = keep.float().mean().item()
    assert abs(frac - (1 - t15 / 32768)) < 2e-3, f"keep fraction {frac}"
    qd = qkv.double().requires_grad_(True)
    x = qd.view(B, S, 3, H, dh).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    s = q @ k.transpose(-1, -2) / math.sqrt(dh)
    pr = torch.softmax(s, dim=-1) * keep.double() * (32768.0 / (32768 - t15))
    o = (pr @ v).permute(0, 2, 1, 3).reshape(B * S, D)
    _cmp("attn ctx (dropout)", ctx, o.detach(), 2e-2, 2e-2)
    _cmp("attn lse (dropout)", lse, torch.logsumexp(s, dim=-1).reshape(B * H, S).detach(), 1e-3, 1e-3)
    o.backward(dctx.double())
    dqkv = ops.attn_bwd(qkv, ctx, dctx, lse, B, S, H, dh, dropout_p=p)
    ref = qd.grad
    for nm, sl in [("dQ", slice(0, D)), ("dK", slice(D, 2 * D)), ("dV", slice(2 * D, 3 * D))]:
        r = ((dqkv[:, sl].double() - ref[:, sl]).norm() / ref[:, sl].norm()).item()
        assert r < 2e-2, f"attn {nm} (dropout) rel fro err {r:.4g}"


@pytest.mark.parametrize("B,S,H,dh", [(2, 365, 4, 12), (1, 200, 2, 36), (3, 365, 16, 36)])
def test_attention_dropout_consistency(B, S, H, dh):
    """With dropout the kernels cannot be compared with torch's RNG; check the algebra instead:
    fwd is linear in V for a fixed mask, E[ctx] ~ no-dropout ctx, and bwd == finite differences of fwd."""
    D = H * dh
    p = 0.1
    qkv = _bf(B * S, 3 * D, seed=22)
    c0, _ = ops.attn_fwd(qkv, B, S, H, dh)
    c1, lse1 = ops.attn_fwd(qkv, B, S, H, dh, dropout_p=p, seed=7, stream_id=3)
    c2, _ = ops.attn_fwd(qkv, B, S, H, dh, dropout_p=p, seed=7, stream_id=3)
    c3, _ = ops.attn_fwd(qkv, B, S, H, dh, dropout_p=p, seed=7, stream_id=4)
    assert torch.equal(c1, c2)
    assert not torch.equal(c1, c3)
    # dropout noise on a 365-key average is small and unbiased
    bias = (c1.double() - c0.double()).mean().abs().item()
    assert bias < 5e-3, f"dropout looks biased: mean diff {bias}"
    # directional derivative check of bwd against fwd (same mask): <dctx, J dv> == <dqkv, dv>
    dctx = _bf(B * S, D, seed=23)
    dqkv = ops.attn_bwd(qkv, c1, dctx, lse1, B, S, H, dh, dropout_p=p)  # keep bits travel with c1
    eps_dir = torch.zeros_like(qkv, dtype=torch.float32)
    g = torch.Generator(device="cuda").manual_seed(9)
    eps_dir[:, 2 * D:] = torch.randn(B * S, D, device="cuda", generator=g)  # perturb V only: fwd is exactly linear in V
    qkv_p = (qkv.float() + eps_dir).to(torch.bfloat16)
    real_dir = qkv_p.float() - qkv.float()
    cp, _ = ops.attn_fwd(qkv_p, B, S, H, dh, dropout_p=p, seed=7, stream_id=3)
    lhs = ((cp.double() - c1.double()) * dctx.double()).sum().item()
    rhs = (dqkv.double() * real_dir.double()).sum().item()
    assert abs(lhs - rhs) <= 3e-2 * max(abs(lhs), abs(rhs)) + 1e-2, f"<dctx, dO> {lhs} vs <dV, dv> {rhs}"
