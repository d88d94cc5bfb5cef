"""CPU tests: the oracle against the golden vectors produced from the reference itself
(tests/golden/make_cpu_golden.py, tests/golden/make_gpu_golden.py) and against independent
numpy formulations of the same arithmetic."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"


def test_packer_matches_reference_to_quantized_weight(oracle):
    """oracle.pack_linear_weights == the reference's own to_quantized_weight / scale / split lines."""
    z = np.load(GOLD / "quantized_weight.npz")
    p = oracle.pack_linear_weights(z["W"], z["act_scale"])
    assert np.array_equal(p["scale_b"].view(np.uint16), z["scale_b"].view(np.uint16))
    assert np.array_equal(p["W8"], z["W8"])
    assert np.array_equal(p["fp_weight"].view(np.uint16), z["fp_weight"].view(np.uint16))
    assert np.array_equal(p["ind"], z["ind"])
    assert (p["W8"][:, p["ind"]] == 0).all()          # outlier columns zeroed before quantising
    assert p["W8"].min() >= -128 and p["W8"].max() <= 127


def test_outlier_selection_on_real_act_scales(oracle):
    """argsort restatement vs torch.sort on the reference's act_scales fixtures (layer 0).
    torch.sort is unstable, so equal scales may come out in another order: the selected VALUES
    must agree everywhere, and the index SET wherever the 128th/129th largest do not tie."""
    a = np.load(GOLD / "act_scales_l0.npz")
    keys = [k for k in a.files if k.endswith("/ind")]
    assert len(keys) == 15
    for k in keys:
        s = a[k[:-4]]
        ind = np.argsort(s, kind="stable")[-128:]
        assert np.array_equal(s[ind], s[a[k]]), k
        srt = np.sort(s)
        if srt[-128] != srt[-129]:
            assert set(ind.tolist()) == set(a[k].tolist()), k


def test_plugin_tensor_containers(oracle):
    """int8 / int32 payloads travel as fp16-typed containers (plugin.py:99-111)."""
    lin = oracle.synth_linear(64, 256)
    t = oracle.as_plugin_tensors(lin)
    assert t["weight"].shape == (64, 128) and t["weight"].dtype == np.float16
    assert t["fp_ind"].shape == (256,) and t["fp_ind"].dtype == np.float16
    assert np.array_equal(t["weight"].view(np.int8).reshape(64, 256), lin["W8"])
    assert np.array_equal(t["fp_ind"].view(np.int32), lin["ind"])


def _np_quant_ieee(A):
    """Independent numpy formulation with IEEE division (what __hdiv computes up to the rcp ulp)."""
    A32 = A.astype(np.float32)
    mx = np.abs(A).max(axis=1)
    sa = (mx.astype(np.float32) * np.float32(1.0 / 127.0)).astype(np.float16)  # fa * rcp(127)
    with np.errstate(divide="ignore", invalid="ignore"):
        q16 = (A32 * (np.float32(1.0) / sa.astype(np.float32))[:, None]).astype(np.float16)
    return np.rint(q16.astype(np.float32)), sa


def test_quant_without_table_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    A = (rng.standard_normal((37, 512)) * 3).astype(np.float16)
    q, sa = oracle.quant(A, use_table=False)
    qn, san = _np_quant_ieee(A)
    assert np.array_equal(sa.view(np.uint16), san.view(np.uint16))
    assert np.array_equal(q.astype(np.float32), qn)
    assert q.min() >= -127 and q.max() <= 127
    assert (np.abs(q).max(axis=1) == 127).all()      # the row max always maps to +-127


def test_quant_edge_cases(oracle):
    A = np.zeros((6, 256), dtype=np.float16)
    A[1, 5] = 1.0
    A[2, :] = np.float16(6e-8)          # smallest subnormal: scale underflows to 0
    A[3, 0] = np.float16(65504.0)       # fp16 max
    A[3, 1] = np.float16(-65504.0)
    A[4, 7] = np.float16(np.inf)
    A[5, 9] = np.float16(np.nan)
    A[5, 10] = 2.0
    q, sa = oracle.quant(A, use_table=False)
    assert (q[0] == 0).all() and sa[0] == 0                       # all-zero row: 0/0 -> NaN -> 0
    assert q[1, 5] == 127 and (np.delete(q[1], 5) == 0).all()
    assert sa[2] == 0 and (q[2] == -1).all()                      # x/0 = +inf -> INT_MAX -> int8 0xFF
    assert q[3, 0] == 127 and q[3, 1] == -127
    assert np.isinf(sa[4].astype(np.float32)) and q[4, 7] == 0    # inf/inf = NaN -> 0
    assert q[5, 9] == 0 and q[5, 10] == 127                       # NaN skipped by __hmax, quantises to 0


def test_mask_mode(oracle):
    """MIXQ_FLAG_MASK_OUTLIERS restates MixQ/src (cult.cu:1588): outlier columns are zero for amax and codes."""
    rng = np.random.default_rng(1)
    A = rng.standard_normal((16, 384)).astype(np.float16)
    ind = np.array([3, 77, 200], dtype=np.int32)
    A[:, ind] *= 50
    q, sa = oracle.quant(A, ind=ind, mask=True, use_table=False)
    A2 = A.copy()
    A2[:, ind] = 0
    q2, sa2 = oracle.quant(A2, use_table=False)
    assert np.array_equal(q, q2) and np.array_equal(sa.view(np.uint16), sa2.view(np.uint16))
    assert (q[:, ind] == 0).all()
    q3, sa3 = oracle.quant(A, use_table=False)                    # plugin mode: scale dominated by outliers
    assert (sa3.astype(np.float32) > sa.astype(np.float32)).all()


def test_igemm_and_epilogue_against_numpy(oracle):
    rng = np.random.default_rng(2)
    M, N, K = 19, 40, 1024
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    w = rng.integers(-128, 128, (N, K), dtype=np.int8)
    acc = oracle.igemm(q, w)
    assert np.array_equal(acc, q.astype(np.int64) @ w.astype(np.int64).T)
    sa = rng.random(M).astype(np.float16)
    sb = (rng.random(N) * 0.01).astype(np.float16)
    out0 = rng.standard_normal((M, N)).astype(np.float16)
    out = oracle.epilogue(acc, sa, sb, out0)
    # fma(float(acc), sb*sa, out0) evaluated exactly in float64, one rounding to f32, one to f16
    p = (sb.astype(np.float32)[None, :] * sa.astype(np.float32)[:, None])
    ref = (acc.astype(np.float32).astype(np.float64) * p.astype(np.float64) + out0.astype(np.float64))
    ref16 = ref.astype(np.float32).astype(np.float16)
    mism = (out.view(np.uint16) != ref16.view(np.uint16)).mean()
    assert mism < 1e-3   # only double-rounding ties of the f64->f32->f16 shortcut may differ
    assert np.array_equal(oracle.epilogue(acc, sa, sb, None), oracle.epilogue(acc, sa, sb, np.zeros_like(out0)))


def test_forward_composition_and_accuracy(oracle):
    """forward() == its steps chained; and the mixed path tracks the fp16 dense product."""
    a = np.load(GOLD / "act_scales_l0.npz")
    lin = oracle.synth_linear(256, 4096, a["Llama-2-7b/self_attn.q_proj"])
    A = oracle.synth_activations(8, lin["act_scale"])
    r = oracle.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], return_parts=True)
    fpA = oracle.gather(A, lin["ind"])
    assert np.array_equal(fpA, A[:, lin["ind"]]) and np.array_equal(fpA, r["fp_A"])
    out0 = oracle.outlier_gemm(fpA, lin["fp_weight"])
    q, sa = oracle.quant(A)
    out = oracle.epilogue(oracle.igemm(q, lin["W8"]), sa, lin["scale_b"], out0)
    assert np.array_equal(out.view(np.uint16), r["out"].view(np.uint16))
    dense = A.astype(np.float64) @ lin["W"].astype(np.float64).T
    rel = np.linalg.norm(out.astype(np.float64) - dense) / np.linalg.norm(dense)
    assert rel < 0.1, rel      # W8A8 quantisation error, plugin mode (outliers inflate the token scale)
    truth = oracle.forward_f64(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], q, sa)
    rel2 = np.linalg.norm(out.astype(np.float64) - truth) / np.linalg.norm(truth)
    assert rel2 < 1e-3, rel2   # fp16 output rounding only


def test_rmsnorm_against_numpy(oracle):
    rng = np.random.default_rng(5)
    X = (rng.standard_normal((9, 640)) * 2).astype(np.float16)
    X[3] = 0
    X[4, 7] = 60000.0
    g = (1 + 0.1 * rng.standard_normal(640)).astype(np.float16)
    Y = oracle.rmsnorm(X, g, 1e-5)
    x32 = X.astype(np.float32)
    s = (1.0 / np.sqrt((x32.astype(np.float64) ** 2).sum(1).astype(np.float32) / np.float32(640) + np.float32(1e-5))).astype(np.float32)
    ref = np.clip((x32 * s[:, None]) * g.astype(np.float32)[None, :], -64504, 64504).astype(np.float16)
    assert np.array_equal(Y.view(np.uint16), ref.view(np.uint16))
    assert (Y[3] == 0).all() and np.isfinite(Y.astype(np.float32)).all()


def test_rcp_table_fixture(oracle):
    """The captured rcp.approx table (when present) is within 1 ulp of the exact reciprocal."""
    t = oracle.rcp_table()
    if t is None:
        pytest.skip("rcp_approx_f16.bin not captured yet (needs one GPU run of make_gpu_golden.py)")
    h = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    ok = np.isfinite(h) & (h != 0)
    with np.errstate(divide="ignore"):
        exact = (np.float32(1.0) / h[ok])
    got = t.view(np.float32)[ok]
    ulp = np.abs(got.view(np.int32).astype(np.int64) - exact.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
    assert np.isinf(t.view(np.float32)[0]) and t.view(np.float32)[0x7C00] == 0


def test_oracle_matches_reference_kernels_fixture(oracle):
    """Outputs of the reference's own CUDA kernels (FindRowScaleKernel, the gather, the whole
    enqueue) captured on a B200 by make_gpu_golden.py must be reproduced by the oracle:
    bit-exact for the int8 codes / scales / gather, within 1 fp16 ulp for the outputs
    (cuBLAS accumulation order is the only freedom)."""
    p = GOLD / "ref_kernels_b200.npz"
    if not p.exists():
        pytest.skip("ref_kernels_b200.npz not captured yet (needs one GPU run of make_gpu_golden.py)")
    z = np.load(p)
    A, ind = z["A"], z["ind"]
    q, sa = oracle.quant(A)
    assert np.array_equal(sa.view(np.uint16), z["ref_sa"].view(np.uint16))
    assert np.array_equal(q, z["ref_q"])
    assert np.array_equal(oracle.gather(A, ind).view(np.uint16), z["ref_fpA"].view(np.uint16))
    out = oracle.forward(A, z["W8"], z["scale_b"], z["fp_weight"], ind)
    ref = z["ref_out"]
    nan_o, nan_r = np.isnan(out.astype(np.float32)), np.isnan(ref.astype(np.float32))
    assert np.array_equal(nan_o, nan_r) and nan_r.any()          # the Inf token poisons its row in both
    ok = ~nan_r
    d = np.abs(out.astype(np.float32) - ref.astype(np.float32))[ok]
    ulp = np.spacing(np.abs(ref).astype(np.float16)).astype(np.float32)[ok]
    assert (d <= ulp).all()
    assert (out.view(np.uint16) != ref.view(np.uint16))[ok].mean() < 0.02


# ------------------------------------------------------------------ M <= 4 branch (weight-only GEMV)
def test_eetq_layout_matches_reference_packer(oracle):
    """eetq_preprocess / eetq_quant_weights vs the reference's own cutlass_preprocessors.cc (compiled unmodified,
    tests/golden/make_gemv_golden.py cpu): processed bytes, plain codes and scales bit-exact, incl. an all-zero
    channel (0/0 -> code 127 through std::min/std::max) and an exact zero."""
    g = np.load(GOLD / "eetq_layout.npz")
    assert np.array_equal(oracle.eetq_preprocess(g["q_kn"]), g["processed"])
    assert np.array_equal(oracle.eetq_unprocess(g["processed"]), g["q_kn"])
    qw, sc = oracle.eetq_quant_weights(g["W_t"])
    assert np.array_equal(sc.view(np.uint16), g["scales"].view(np.uint16))
    assert np.array_equal(oracle.eetq_unprocess(qw), g["codes"])
    assert np.array_equal(qw, g["processed_codes"])


def test_gemv_oracle_matches_reference_kernel_fixture(oracle):
    """mixq_oracle_gemv_w8a16 vs outputs of the reference's weight_only_batched_gemv kernels run on a B200
    (tests/golden/ref_gemv_b200.npz, make_gemv_golden.py gpu): bit-exact (fp16 chains + fp32 tree restated)."""
    import sys
    sys.path.insert(0, str(GOLD))
    import make_gemv_golden as G
    p = GOLD / "ref_gemv_b200.npz"
    if not p.exists():
        pytest.skip("ref_gemv_b200.npz not captured yet")
    g = np.load(p)
    for i, (M, N, K, seed) in enumerate(G.GEMV_CASES):
        A, W_t = G.gemv_case(M, N, K, seed)
        qw, sc = oracle.eetq_quant_weights(W_t)
        assert G.crc(A, qw, sc) == g[f"crc{i}"], "inputs regenerated from the seed differ from the captured ones"
        out = oracle.gemv_w8a16(A, qw, sc)
        assert np.array_equal(out.view(np.uint16), g[f"out{i}"].view(np.uint16)), (i, M, N, K)


def test_gemv_oracle_close_to_float64(oracle):
    rng = np.random.default_rng(3)
    K, N, M = 2048, 32, 4
    W_t = (rng.standard_normal((K, N)) * 0.02).astype(np.float16)
    A = rng.standard_normal((M, K)).astype(np.float16)
    qw, sc = oracle.eetq_quant_weights(W_t)
    out = oracle.gemv_w8a16(A, qw, sc).astype(np.float64)
    ref = A.astype(np.float64) @ (oracle.eetq_unprocess(qw).astype(np.float64) * sc.astype(np.float64)[None, :])
    assert np.abs(out - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-3


def test_epilogue_ex_reduces_to_epilogue(oracle):
    """epilogue_ex without bias / activation is the reference epilogue (the C restatement) bit for bit."""
    rng = np.random.default_rng(2)
    acc = rng.integers(-3_000_000, 3_000_000, (16, 64), dtype=np.int32)
    sa = (rng.random(16) * 0.02 + 1e-3).astype(np.float16)
    sb = (rng.random(64) * 2e-3 + 1e-4).astype(np.float16)
    out0 = rng.standard_normal((16, 64)).astype(np.float16)
    a = oracle.epilogue(acc, sa, sb, out0)
    b = oracle.epilogue_ex(acc, sa, sb, out0)
    assert np.array_equal(a.view(np.uint16), b.view(np.uint16))
    bias = rng.standard_normal(64).astype(np.float16)
    c = oracle.epilogue_ex(acc, sa, sb, out0, bias=bias)
    assert np.array_equal(c.view(np.uint16), (a.astype(np.float32) + bias.astype(np.float32)[None, :]).astype(np.float16).view(np.uint16))
    d = oracle.epilogue_ex(acc, sa, sb, out0, silu=True).astype(np.float32)
    x = a.astype(np.float32)
    assert np.allclose(d, x / (1 + np.exp(-x)), rtol=2e-3, atol=2e-3)


def test_emulated_hfma_is_the_exactly_rounded_fma(oracle):
    """dev_hfma (the fp16 FMA chain of the weight-only GEMV restatement) against exact rational arithmetic: a*b+c
    computed with fractions and rounded ONCE to fp16, round-to-nearest-even, over random and adversarial triples
    (ties created by a tiny addend next to a 22-bit product, subnormals, overflow to inf, signed zeros)."""
    from fractions import Fraction
    L = oracle.lib()

    def f16(bits):
        return np.array([bits], dtype=np.uint16).view(np.float16)[0]

    def round_f16(x: Fraction) -> int:
        """exact RN-even rounding of a rational to fp16 bits"""
        if x == 0:
            return 0
        sign = 0x8000 if x < 0 else 0
        x = abs(x)
        e = -24                                     # find e with 2^e <= ulp grid: value = m * 2^e, m < 2^11 for normals
        import math
        ex = math.floor(math.log2(float(x))) if x >= Fraction(1, 2 ** 30) else -40
        while Fraction(2) ** ex > x:
            ex -= 1
        while Fraction(2) ** (ex + 1) <= x:
            ex += 1
        q = max(ex - 10, -24)                       # quantum exponent (subnormals share 2^-24)
        m = x / Fraction(2) ** q
        fl = m.numerator // m.denominator
        rem = m - fl
        if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (fl & 1)):
            fl += 1
        val = Fraction(fl) * Fraction(2) ** q
        if val >= Fraction(65520):
            return sign | 0x7C00
        return sign | int(np.array([float(val)], dtype=np.float16).view(np.uint16)[0])

    rng = np.random.default_rng(0)
    triples = []
    for _ in range(4000):
        a, b, c = (int(v) for v in rng.integers(0, 0x7C00, 3))       # finite, positive magnitudes ...
        s = rng.integers(0, 2, 3) * 0x8000                            # ... random signs
        triples.append((a | int(s[0]), b | int(s[1]), c | int(s[2])))
    # adversarial: product exactly on an fp16 tie, addend far below -> the sticky information decides
    for pa, pb in ((0x3C01, 0x3C01), (0x4401, 0x3C03), (0x3E01, 0x3A01), (0x7801, 0x3401)):
        for cbits in (0x0001, 0x8001, 0x0000, 0x8000, 0x0400, 0x8400):
            triples.append((pa, pb, cbits))
    triples += [(0x7BFF, 0x7BFF, 0x0001), (0x0001, 0x0001, 0x0001), (0x0001, 0x3C00, 0x8001), (0x8000, 0x3C00, 0x0000)]
    for a, b, c in triples:
        exact = Fraction(float(f16(a))) * Fraction(float(f16(b))) + Fraction(float(f16(c)))
        want = round_f16(exact)
        got = int(L.mixq_oracle_hfma(a, b, c))
        if exact == 0:                              # signed zero: IEEE sum rule, not representable in a Fraction
            assert got & 0x7FFF == 0
            continue
        assert got == want, (hex(a), hex(b), hex(c), hex(got), hex(want))


def test_mixsrc_forward_dynamic_outliers(oracle):
    """configs[0] as BASELINE.json words it ("via MixQ/src torch reference"): weights quantised without outlier handling,
    outlier columns discovered at run time.  Checked against an independent numpy restatement of linear.py:163-286 that
    uses float64 for the products (the INT8 part is exact; the outlier product differs by accumulation order only)."""
    rng = np.random.default_rng(5)
    N, K = 64, 256
    W = (rng.standard_normal((N, K)) * 0.02).astype(np.float16)
    st = oracle.mixsrc_init(W)
    assert st["ind"].size == 0 and st["q_weight"].dtype == np.int8 and np.abs(st["q_weight"]).max() <= 127
    assert np.array_equal(st["scale_col"], (np.abs(W).astype(np.float32).max(1) / 127).astype(np.float16))
    x1 = rng.standard_normal((3, K)).astype(np.float16)                # nothing above sigma = 6: no outliers yet
    y1 = oracle.mixsrc_forward(st, x1)
    assert st["ind"].size == 0 and st["cnt"] == 1 and st["add_outliers"]
    q, sa = oracle.quant(x1)
    assert np.array_equal(y1.view(np.uint16), oracle.epilogue(oracle.igemm(q, st["q_weight"]), sa, st["scale_col"], None).view(np.uint16))
    x2 = rng.standard_normal((3, K)).astype(np.float16)
    x2[0, 7], x2[2, 100], x2[1, 7] = 40.0, -25.0, 9.0                  # columns 7 and 100 exceed sigma
    y2 = oracle.mixsrc_forward(st, x2)
    assert st["ind"].tolist() == [7, 100] and st["weight_cache"].shape == (N, 2) and not st["add_outliers"]   # stop = 2 calls
    assert np.array_equal(st["weight_cache"], (st["q_weight"][:, [7, 100]].astype(np.float16) * st["scale_col"][:, None]))
    xm = x2.copy(); xm[:, [7, 100]] = 0
    q, sa = oracle.quant(xm)
    want = (q.astype(np.float64) @ st["q_weight"].astype(np.float64).T) * (sa.astype(np.float64)[:, None] * st["scale_col"].astype(np.float64)[None, :]) \
        + x2[:, [7, 100]].astype(np.float64) @ st["weight_cache"].astype(np.float64).T
    assert np.abs(y2.astype(np.float64) - want).max() <= 2e-3 * np.abs(want).max() + 1e-3
    x3 = x2.copy(); x3[1, 33] = 50.0                                    # growth has stopped: column 33 stays in the INT8 part
    y3 = oracle.mixsrc_forward(st, x3)
    assert st["ind"].tolist() == [7, 100] and np.isfinite(y3.astype(np.float32)).all()
