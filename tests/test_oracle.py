"""Pin the CPU oracle (oracle/wm_oracle.py) against golden vectors produced by the unmodified reference
(oracle/make_golden.py, run in the build container) and against published known answers. CPU only."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import wm_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, name)))


def _sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_cpu_generator_rand_known_answer():
    g = _load("masks_cpu.npz")
    assert np.array_equal(O.torch_cpu_rand(1234, 5), g["rand5_seed1234"])


@pytest.mark.parametrize("name,kind,arg", [("bert_p15", "bert", 0.15), ("bert_p30", "bert", 0.30),
                                           ("former_n10", "former", 10), ("former_n1", "former", 1)])
def test_cpu_masks_match_reference(name, kind, arg):
    """SURVEY.md 8(c) golden vectors: torch.manual_seed(1234); masking_function(365, 31, 64) on CPU."""
    g = _load("masks_cpu.npz")
    if kind == "bert":
        m = O.weatherbert_mask(O.torch_cpu_rand(1234, 64 * 365 * 31).reshape(64, 365, 31), arg)
        packed = np.packbits(m)
    else:
        m = O.weatherformer_mask(O.torch_cpu_rand(1234, 64 * 31).reshape(64, 31), arg, 365)
        packed = np.packbits(m[:, 0, :])
    assert int(m.sum()) == int(g[name + "_sum"][0])
    assert _sha16(np.ascontiguousarray(m)) == bytes(g[name + "_sha16"]).decode()
    assert np.array_equal(m[0, 0], g[name + "_row00"])
    assert np.array_equal(packed, g[name + "_packed"])


def test_philox_known_answer():
    """Random123 philox4x32-10 KATs (counter, key) -> output."""
    r = O.philox4x32_10(0, np.array([0], dtype=np.uint64), np.array([0], dtype=np.uint64))[0]
    assert [hex(int(x)) for x in r] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # counter = ffffffff x4, key = ffffffff x2
    r = O.philox4x32_10(0xFFFFFFFFFFFFFFFF, np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64),
                        np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
    assert [hex(int(x)) for x in r] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_cuda_rand_structure():
    """Without a GPU only structural properties can be checked here (bit-exactness vs torch.rand on the
    CUDA generator is a -m gpu test): range, determinism, offset bookkeeping."""
    n = 64 * 31
    gx = O.rand_grid_x(n)
    a = O.torch_cuda_rand(1234, 0, gx, n)
    assert a.dtype == np.float32 and (a >= 0).all() and (a < 1).all()
    assert np.array_equal(a, O.torch_cuda_rand(1234, 0, gx, n))
    assert not np.array_equal(a, O.torch_cuda_rand(1234, 4, gx, n))
    assert O.rand_offset_increment(n, gx) == 4
    big = 4096 * 365 * 31
    assert O.rand_grid_x(big) == 1184 and O.rand_offset_increment(big, 1184) == ((big - 1) // (256 * 1184 * 4) + 1) * 4


@pytest.mark.parametrize("fname,model", [("weatherbert_mini_b8.npz", "weatherbert"),
                                         ("weatherformer_mini_b8.npz", "weatherformer")])
def test_encoder_loss_and_grads_match_reference(fname, model):
    g = _load(fname)
    state = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    losses, Y, grads = O.train_step_grads(state, int(g["num_heads"][0]), model, g["weather"], g["coords"], g["year"],
                                          g["interval"], g["mask"], beta=float(g["beta"][0]))
    if model == "weatherbert":
        assert np.allclose(Y, g["y"], rtol=1e-4, atol=2e-5)
        assert abs(losses["total_loss"] - g["loss"][0]) < 1e-5 * abs(g["loss"][0])
    else:
        mu, var = O.weatherformer_head(Y, 31)
        assert np.allclose(mu, g["mu"], rtol=1e-4, atol=2e-5)
        assert np.allclose(var, g["var"], rtol=1e-4, atol=2e-6)
        for got, ref in zip([losses["total_loss"], losses["reconstruction"], losses["kl_term"]], g["loss"]):
            assert abs(got - ref) < 1e-5 * abs(ref)
    for k, v in g.items():
        if not k.startswith("grad/"):
            continue
        got = grads[k[len("grad/"):]]
        rel = np.linalg.norm(got - v) / (np.linalg.norm(v) + 1e-30)
        assert rel < 1e-3, f"{k}: relative error {rel}"  # fp32 reference vs fp64 oracle noise floor is ~4e-4


def test_bf16_storage_oracle_is_the_same_algorithm():
    """EncoderOracleBf16 with the rounding switched off IS EncoderOracle (so it inherits the golden pinning); with
    it on, it predicts the size of the bf16 gap to the fp32 reference (what the B200 path is then measured against in
    tests/test_gpu_model.py)."""
    assert np.array_equal(O.bf16_round(np.array([1.0, 1.00390625, 1.001953125, 1.005859375, -3.14159, 0.0])),
                          np.array([1.0, 1.0, 1.0, 1.0078125, -3.140625, 0.0]))  # ties to even; 2^-8 spacing at 1
    g = _load("weatherformer_mini_b8.npz")
    state = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    args = (g["weather"], g["coords"], g["year"], g["interval"], g["mask"])
    l0, Y0, g0 = O.train_step_grads(state, 4, "weatherformer", *args, beta=0.5)
    enc = O.EncoderOracleBf16(state, 4, quant=False)
    Y = enc.forward(*args)
    _, dY = O.former_elbo(Y, g["weather"], g["mask"], 0.5, 31)
    gi = enc.backward(dY)
    assert np.allclose(Y, Y0, rtol=0, atol=1e-6)  # (year / coords are normalised in float32 as the kernel does)
    for k, v in g0.items():
        assert np.linalg.norm(gi[k] - v) <= 1e-6 * np.linalg.norm(v), k
    lq, _, gq = O.train_step_grads(state, 4, "weatherformer", *args, beta=0.5, storage="bf16")
    assert abs(lq["total_loss"] - l0["total_loss"]) < 1e-3 * l0["total_loss"]
    rels = [np.linalg.norm(gq[k] - v) / np.linalg.norm(v) for k, v in g0.items()]
    assert 2e-3 < max(rels) < 3e-2, max(rels)  # the bf16 operand-rounding gap: ~1e-2 on the worst tensor


def test_known_answers_from_survey():
    """loss and gradient norms of the reference (CPU fp32) for the mini models."""
    g = _load("weatherbert_mini_b8.npz")
    assert int(g["mask"].sum()) == 13435
    state = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    losses, _, grads = O.train_step_grads(state, 4, "weatherbert", g["weather"], g["coords"], g["year"], g["interval"], g["mask"])
    assert abs(losses["total_loss"] - 1.33503056) < 2e-5
    assert abs(np.linalg.norm(grads["in_proj.weight"]) - 0.08772561) < 5e-6
    assert abs(np.linalg.norm(grads["out_proj.weight"]) - 0.65765929) < 1e-5
    g = _load("weatherformer_mini_b8.npz")
    state = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    losses, _, grads = O.train_step_grads(state, 4, "weatherformer", g["weather"], g["coords"], g["year"], g["interval"],
                                          g["mask"], beta=0.5)
    assert abs(losses["total_loss"] - 1.75079405) < 2e-5
    assert abs(losses["reconstruction"] - 1.66437364) < 2e-5
    assert abs(losses["kl_term"] - 0.08642045) < 2e-6
    assert abs(np.linalg.norm(grads["in_proj.weight"]) - 0.07980558) < 5e-6
    assert abs(np.linalg.norm(grads["out_proj.weight"]) - 0.70168072) < 1e-5


def test_adam_matches_reference_step():
    g = _load("weatherbert_mini_b8.npz")
    for name in ["in_proj.weight", "out_proj.bias"]:
        p0, gr = g["param/" + name], g["grad/" + name]
        p1, _, _ = O.adam_step(p0, gr, np.zeros_like(p0), np.zeros_like(p0), 1, 5e-4)
        assert np.allclose(p1, g["adam/" + name], rtol=1e-6, atol=1e-9)


def test_schedules_sizes_and_pe():
    g = _load("schedules.npz")
    for nm, warm, total, decay in [("exp", 10.0, 100, 0.99), ("cos", 5, 50, None), ("nowarm", 0, 20, 0.9)]:
        lrs = np.array([5e-4 * O.lr_lambda(e, warm, total, decay) for e in range(total)])
        assert np.allclose(lrs, g[nm], rtol=1e-12, atol=0)
    assert g["exp"][0] == 0.0  # lr is 0 during epoch 0 when warm-up > 0
    assert np.allclose(O.positional_encoding(365, 48), g["pe_mini"], rtol=0, atol=5e-5)  # fp32 sin/exp differ by ulps between numpy and torch
    assert [O.n_masked_features_schedule(e, 10) for e in (None, 0, 4, 5, 37, 99)] == [10, 10, 10, 12, 24, 25]
    with pytest.raises(ValueError):
        O.get_model_params("huge")
    assert int(g["params_weatherformer_large"][0]) == 31966334 and int(g["params_weatherbert_mini"][0]) == 59743


def test_years_and_normalisation():
    idx = np.array([0.0, 1.0], dtype=np.float32)
    yrs = O.years_from_index(idx, np.array([7.0, 7.0], dtype=np.float32), 365)
    assert yrs.dtype == np.float32 and yrs[0, 0] == 1984.0
    assert abs(float(yrs[1, 364]) - (1984 + ((365 + 364) * 7) / 365)) < 1e-3 and yrs.max() < 2002.0
    y, i, c = O.normalize_year_interval_coords(yrs.astype(np.float64), np.array([[7.0], [7.0]]), np.array([[36.0, -90.0], [10.0, 20.0]]))
    assert np.allclose(c, [[0.1, -0.5], [10 / 360, 20 / 180]]) and np.allclose(i, 7 / 30) and np.allclose(y[0, 0], 0.14)


def test_dropout_hash_statistics():
    """The counter hash behind every dropout mask (csrc/wm_common.cuh: drop_hash64), restated in numpy: avalanche,
    uniformity, the keep rate of the 15-bit compare for the reference's p = 0.1 and (absence of) correlation between
    the four elements of a counter, neighbouring counters and streams."""
    m32 = np.uint64(0xFFFFFFFF)

    def fold(v):
        return ((v >> np.uint64(32)) ^ (v & m32)) & m32

    def h64(x, k0, k1, k2):
        x = x.astype(np.uint64)
        hl = fold(((x ^ np.uint64(k0)) & m32) * np.uint64(0x9E3779B1))
        a = fold(((hl ^ np.uint64(k1)) & m32) * np.uint64(0x85EBCA77))
        b = fold(((hl ^ np.uint64(k2)) & m32) * np.uint64(0xC2B2AE3D))
        return a.astype(np.uint32), b.astype(np.uint32)

    n = 1 << 18
    x = np.arange(n, dtype=np.uint32) + np.uint32(777)
    k0, k1, k2 = 0xA5A5F00D, 0x1234ABCD, 0x77AA1357
    a, b = h64(x, k0, k1, k2)
    for bit in range(32):  # flipping any input bit flips every output bit about half of the time
        a2, b2 = h64(x ^ np.uint32(1 << bit), k0, k1, k2)
        for r, r2 in ((a, a2), (b, b2)):
            frac = np.unpackbits((r ^ r2).view(np.uint8)).reshape(n, 32).mean(0)
            assert frac.min() > 0.48 and frac.max() < 0.52, (bit, frac.min(), frac.max())
    by = np.concatenate([a, b]).view(np.uint8)
    hist = np.bincount(by, minlength=256)
    chi2 = ((hist - hist.mean()) ** 2 / hist.mean()).sum()
    assert chi2 < 340, chi2  # 255 degrees of freedom: 340 is the 0.9997 quantile
    # the kernels' test: keep iff (half & 0x7FFF) >= thresh15, thresh15 = (round(65536 p) + 1) >> 1
    t15 = (int(0.1 * 65536 + 0.5) + 1) >> 1
    assert abs(t15 / 32768 - 0.1) < 1e-5  # the reference's nn.Dropout(0.1) to 1e-5 (round 1 trained at 13/128 = 0.1016)
    fields = np.stack([a & 0x7FFF, (a >> 16) & 0x7FFF, b & 0x7FFF, (b >> 16) & 0x7FFF], 1).astype(np.int64)
    keep = (fields >= t15).astype(np.float64)
    assert np.abs(keep.mean(0) - (1 - t15 / 32768)).max() < 2.5e-3
    assert np.abs(np.corrcoef(keep.T) - np.eye(4)).max() < 8e-3  # the four elements of one counter
    flat = keep.reshape(-1)
    kc = flat - flat.mean()
    for lag in (1, 2, 4, 16, 64):
        assert abs((kc[:-lag] * kc[lag:]).mean() / kc.var()) < 6e-3, lag
    oa, ob = h64(x, k0 ^ 0x9E3779B9, (k1 + 0x7F4A7C15) & 0xFFFFFFFF, (k2 + 0x5851F42D) & 0xFFFFFFFF)
    other = (np.stack([oa & 0x7FFF, (oa >> 16) & 0x7FFF, ob & 0x7FFF, (ob >> 16) & 0x7FFF], 1) >= t15).astype(np.float64).reshape(-1)
    assert abs(((other - other.mean()) * kc).mean() / np.sqrt(other.var() * kc.var())) < 6e-3
