"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI
(libmixq_b200.so via ctypes); the checkers are the CPU oracle (oracle/) and, when present, the
reference's own CUDA kernels compiled from /root/reference/kernel/i8gemm.cu (oracle/_ref).

Tolerances:
  * INT8 codes, per-token scales, outlier gather: bit-exact vs the reference kernels and vs the
    oracle (which emulates device __hdiv with the captured rcp.approx table);
  * int32 accumulators: exact -- with the outlier slab disabled the fp16 output must equal the
    oracle bit for bit;
  * full mixed output: <= 2 fp16 ulp element-wise and rel-Frobenius <= 1e-3 vs the reference
    kernels / oracle (the only freedom is the accumulation order of the 128-term fp16 outlier
    dot product inside cuBLAS / the tensor core).
"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import refgpu

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
DEV = "cuda"
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def B(lib):
    from mixq_tensorrt_llm_b200 import binding
    binding.require_device()
    return binding


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


_GEMM_WS = {}


def _gemm_ws(lib, M=1024, N=12288):
    """split-K scratch for mixq_gemm_dequant_ws/_opt (decode-kernel size for up to 1024 x 28672), deliberately filled
    with garbage once: the library must clear / re-arm its flags itself."""
    need = lib.mixq_decode_workspace_size(min(M, 1024), N)
    if "ws" not in _GEMM_WS or _GEMM_WS["ws"].numel() < need:
        _GEMM_WS["ws"] = torch.randint(0, 255, (max(need, lib.mixq_decode_workspace_size(1024, 28672)),), dtype=torch.uint8, device=DEV)
    return _GEMM_WS["ws"]


def _assert_mixed_close(got, ref, out0, what="", mag=None):
    """Tolerance of the mixed output: the fp16-rounded outlier product `out0` may
    differ from the checker's by one fp16 ulp *of out0* (accumulation order of its 128-term dot
    product inside cuBLAS / the tensor core / the oracle), and the final rounding adds at most one
    ulp of the result.  When the int8 part cancels the outlier part the result is much smaller
    than out0, so the bound has to be expressed in ulps of out0, not of the result."""
    g32, r32 = got.astype(np.float32), ref.astype(np.float32)
    fin = np.isfinite(r32)
    assert np.array_equal(np.isfinite(g32), fin), what
    bound = np.spacing(np.abs(out0).astype(np.float16)).astype(np.float32) + np.spacing(np.abs(ref).astype(np.float16)).astype(np.float32)
    if mag is not None:   # fp32 accumulation-order error of the 128-term outlier dot product, <= 2^-20 * sum|a_j w_j|
        bound = bound + (mag * 2.0 ** -20).astype(np.float32)
    d = np.abs(g32 - r32)
    bad = fin & (d > bound)
    assert not bad.any(), f"{what}: {int(bad.sum())} elements beyond 1 ulp(out0) + 1 ulp(out); worst {float((d / bound)[fin].max()):.2f}x"
    rel = np.linalg.norm((g32 - r32)[fin].astype(np.float64)) / max(np.linalg.norm(r32[fin].astype(np.float64)), 1e-30)
    assert rel <= 1e-3, f"{what}: rel-Frobenius {rel:.2e}"
    return float((got.view(np.uint16) != ref.view(np.uint16))[fin].mean())


def _ulp_diff(a, b):
    """distance in fp16 ulps (of b) between two fp16 arrays"""
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    ulp = np.spacing(np.abs(b).astype(np.float16)).astype(np.float32)
    with np.errstate(invalid="ignore"):
        return np.abs(a32 - b32) / ulp


def _edge_activations(O, M, K, act_scale=None, seed=0):
    if act_scale is None:
        act_scale = O.synth_act_scale(K, seed)
    A = O.synth_activations(M, act_scale, seed=seed + 100)
    if M >= 10:
        A[1] = 0
        A[2, :] = np.float16(6e-8)
        A[3, K // 2] = np.float16(65504.0)
        A[4, 3] = np.float16(np.inf)
        A[5, 7] = np.float16(np.nan)
        A[6] = (np.arange(K) % 255 - 127).astype(np.float16) * np.float16(0.5)
        A[7, :] = np.float16(-1.0)
    return A


# ----------------------------------------------------------------------------- stage 1
@pytest.mark.parametrize("M,K", [(1, 256), (5, 256), (48, 512), (33, 4096), (512, 4096), (64, 11008), (3, 28672),
                                 (1500, 1024)])
@pytest.mark.parametrize("mask", [False, True])
def test_quant_extract_bit_exact(B, oracle, M, K, mask):
    A = _edge_activations(oracle, M, K, seed=M + K)
    rng = np.random.default_rng(K)
    ind = rng.choice(K, size=128, replace=False).astype(np.int32)
    tA, tind = _t(A), _t(ind)
    A8 = torch.full((M, K), 99, dtype=torch.int8, device=DEV)
    sa = torch.zeros(M, dtype=torch.float16, device=DEV)
    fpA = torch.zeros(M, 128, dtype=torch.float16, device=DEV)
    B.quant_extract(tA, tind, A8, sa, fpA, flags=B.FLAG_MASK_OUTLIERS if mask else 0)
    torch.cuda.synchronize()
    assert torch.equal(tA.view(torch.int16), _t(A).view(torch.int16))          # input untouched
    # gather: exact copy
    assert np.array_equal(fpA.cpu().numpy().view(np.uint16), A[:, ind].view(np.uint16))
    # oracle (device-__hdiv emulation needs the captured rcp table for bit-exactness)
    oq, osa = oracle.quant(A, ind=ind, mask=mask)
    gq, gsa = A8.cpu().numpy(), sa.cpu().numpy()
    assert np.array_equal(gsa.view(np.uint16), osa.view(np.uint16))
    if oracle.rcp_table() is not None:
        assert np.array_equal(gq, oq)
    else:
        assert (gq != oq).mean() < 1e-4 and np.abs(gq.astype(int) - oq.astype(int)).max() <= 1
    # the reference's own kernels (plugin semantics = no mask)
    if refgpu.available() and not mask:
        rq, rsa = refgpu.int8quant(tA)
        rfp = refgpu.extract(tA, tind)
        torch.cuda.synchronize()
        assert torch.equal(rsa.view(torch.int16), sa.view(torch.int16))
        assert torch.equal(rq, A8)
        assert torch.equal(rfp.view(torch.int16), fpA.view(torch.int16))


def test_quant_only_no_outliers(B, oracle):
    A = _edge_activations(oracle, 40, 768, seed=3)
    tA = _t(A)
    A8 = torch.empty(40, 768, dtype=torch.int8, device=DEV)
    sa = torch.empty(40, dtype=torch.float16, device=DEV)
    B.quant_extract(tA, None, A8, sa, None)
    torch.cuda.synchronize()
    oq, osa = oracle.quant(A)
    assert np.array_equal(sa.cpu().numpy().view(np.uint16), osa.view(np.uint16))
    assert (A8.cpu().numpy() != oq).mean() < (1e-4 if oracle.rcp_table() is None else 1e-12)


@pytest.mark.parametrize("M,K", [(7, 256), (64, 4096), (300, 3584), (40, 11008)])
@pytest.mark.parametrize("mask", [True, False])
def test_rmsnorm_quant_extract(B, oracle, M, K, mask):
    """'next #1': RMSNorm fused into stage 1 (layernorm.cu:121-198).  The normalised row may differ from the
    CPU restatement in the last fp16 bit (fp32 tree sum + MUFU.RSQ on the device); everything downstream of
    it (gather, scale, INT8 codes) must be bit-exact given the device's own normalised row."""
    rng = np.random.default_rng(M + K)
    X = (rng.standard_normal((M, K)) * 1.5).astype(np.float16)
    X[:, rng.choice(K, 16, replace=False)] *= 30
    if M > 4:
        X[1] = 0
    gamma = (1 + 0.2 * rng.standard_normal(K)).astype(np.float16)
    ind = rng.choice(K, size=128, replace=False).astype(np.int32)
    A8 = torch.empty(M, K, dtype=torch.int8, device=DEV)
    sa = torch.empty(M, dtype=torch.float16, device=DEV)
    fpA = torch.empty(M, 128, dtype=torch.float16, device=DEV)
    Y = torch.empty(M, K, dtype=torch.float16, device=DEV)
    B.rmsnorm_quant_extract(_t(X), _t(gamma), 1e-5, _t(ind), A8, sa, fpA, Y, flags=B.FLAG_MASK_OUTLIERS if mask else 0)
    torch.cuda.synchronize()
    y = Y.cpu().numpy()
    yo = oracle.rmsnorm(X, gamma, 1e-5)
    assert _ulp_diff(y, yo).max() <= 1.0
    assert (y.view(np.uint16) != yo.view(np.uint16)).mean() < 0.02
    oq, osa = oracle.quant(y, ind=ind, mask=mask)
    assert np.array_equal(fpA.cpu().numpy().view(np.uint16), y[:, ind].view(np.uint16))
    assert np.array_equal(sa.cpu().numpy().view(np.uint16), osa.view(np.uint16))
    assert np.array_equal(A8.cpu().numpy(), oq)
    # and it equals the unfused sequence on the device: mixq_quant_extract applied to the same y
    A8b, sab, fpAb = torch.empty_like(A8), torch.empty_like(sa), torch.empty_like(fpA)
    B.quant_extract(Y, _t(ind), A8b, sab, fpAb, flags=B.FLAG_MASK_OUTLIERS if mask else 0)
    torch.cuda.synchronize()
    assert torch.equal(A8, A8b) and torch.equal(sa.view(torch.int16), sab.view(torch.int16))


@pytest.mark.skipif(not refgpu.mixsrc_available(), reason="oracle/_ref/libref_mixsrc.so not built")
@pytest.mark.parametrize("M,K", [(7, 1024), (64, 4096), (300, 3584), (40, 11008), (512, 4096)])   # the reference's max-reduction assumes >= 256 threads, i.e. K >= 512
def test_rmsnorm_quant_extract_vs_reference_kernel(B, oracle, M, K):
    """SURVEY 8f #1 pinned to the reference's own kernel: generalT5LayerNorm_extract_outliers (layernorm.cu:121-198)
    compiled unmodified (oracle/_ref/libref_mixsrc.so) and run on the same inputs.  The kernel always zeroes the outlier
    columns (= MIXQ_FLAG_MASK_OUTLIERS).  Its sum of squares is a different fp32 tree from ours, so a normalised value may
    differ in the last fp16 bit; every token whose normalised row agrees must agree bit for bit in scale and INT8 codes,
    and the oracle's restatement must sit within the same 1 ulp of the reference kernel."""
    rng = np.random.default_rng(M * 7 + K)
    X = (rng.standard_normal((M, K)) * 1.5).astype(np.float16)
    X[:, rng.choice(K, 16, replace=False)] *= 30
    gamma = (1 + 0.2 * rng.standard_normal(K)).astype(np.float16)
    ind = np.sort(rng.choice(K, size=128, replace=False)).astype(np.int32)
    tX, tg, ti = _t(X), _t(gamma), _t(ind)
    r_out, r_outl, r_q, r_sc = refgpu.rmsnorm_extract_quant(tX, tg, 1e-5, ti)
    A8 = torch.empty(M, K, dtype=torch.int8, device=DEV)
    sa = torch.empty(M, dtype=torch.float16, device=DEV)
    fpA = torch.empty(M, 128, dtype=torch.float16, device=DEV)
    Y = torch.empty(M, K, dtype=torch.float16, device=DEV)
    B.rmsnorm_quant_extract(tX, tg, 1e-5, ti, A8, sa, fpA, Y, flags=B.FLAG_MASK_OUTLIERS)
    torch.cuda.synchronize()
    y = Y.cpu().numpy().copy()
    ref_y = r_out.cpu().numpy()
    ref_full = ref_y.copy()
    ref_full[:, ind] = r_outl.cpu().numpy()                     # the reference's normalised row before the zeroing
    assert _ulp_diff(y, ref_full).max() <= 1.0
    assert _ulp_diff(oracle.rmsnorm(X, gamma, 1e-5), ref_full).max() <= 1.0
    same_row = (y.view(np.uint16) == ref_full.view(np.uint16)).all(axis=1)
    assert same_row.mean() >= 0.5, same_row.mean()
    assert np.array_equal(fpA.cpu().numpy().view(np.uint16)[same_row], r_outl.cpu().numpy().view(np.uint16)[same_row])
    assert np.array_equal(sa.cpu().numpy().view(np.uint16)[same_row], r_sc.cpu().numpy().view(np.uint16)[same_row])
    assert np.array_equal(A8.cpu().numpy()[same_row], r_q.cpu().numpy()[same_row])
    # rows that differ by an ulp somewhere: codes within 1 of the reference's
    assert np.abs(A8.cpu().numpy().astype(np.int32) - r_q.cpu().numpy().astype(np.int32)).max() <= 1


@pytest.mark.skipif(not refgpu.mixsrc_available(), reason="oracle/_ref/libref_mixsrc.so not built")
@pytest.mark.parametrize("cfg", [0, 5, 9, 13])
@pytest.mark.parametrize("M,N,K", [(96, 512, 1024), (512, 1376, 4096), (300, 2048, 3584)])
def test_silu_epilogue_vs_reference_kernel(B, lib, oracle, cfg, M, N, K):
    """SURVEY 8f #4 pinned to the reference's own kernel: GemmDequantSilu instantiated as int8FusedDequantizeSiluCUDA does
    (cult.cu:2248-2273; epilogue functor linear_combination_dequant.h:167-272), compiled unmodified with the reference's
    --use_fast_math, on the same INT8 operands, scales and fp16 addend.  Same op sequence (one FMA, fast-math SiLU, one
    rounding), so the outputs must agree bit for bit -- without the outlier slab (addend 0, the reference's cache.zeros path,
    linear.py:321-323) and with it (the addend given to the reference is this library's own fp16 outlier product)."""
    if cfg == 13 and M <= 128:
        pytest.skip("fat tiles serve 128 < M <= 1024")
    rng = np.random.default_rng(M + N + K)
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    w = rng.integers(-127, 128, (N, K), dtype=np.int8)
    sa = (rng.random(M) * 2e-3 + 1e-4).astype(np.float16)
    sb = (rng.random(N) * 2e-3 + 1e-4).astype(np.float16)
    fpA = (rng.standard_normal((M, 128)) * 2).astype(np.float16)
    fpW = (rng.standard_normal((N, 128)) * 0.05).astype(np.float16)
    tq, tw, tsa, tsb, tfa, tfw = _t(q), _t(w), _t(sa), _t(sb), _t(fpA), _t(fpW)
    ws = _gemm_ws(lib, M, N)
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    zeros = torch.zeros(M, N, dtype=torch.float16, device=DEV)
    B.gemm_dequant(tq, tw, tsa, tsb, None, None, out, workspace=ws, activation=B.ACT_SILU, config=cfg)
    ref = refgpu.int8_fused_dequant_silu(tq, tw, tsa, tsb, zeros)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int16), ref.view(torch.int16)), int((out.view(torch.int16) != ref.view(torch.int16)).sum())
    # with the outlier product: out0 = this library's fp16 product (a zero INT8 operand leaves fp16(fma(0, s, out0)) = out0)
    out0 = torch.empty(M, N, dtype=torch.float16, device=DEV)
    B.gemm_dequant(torch.zeros_like(tq), tw, tsa, tsb, tfa, tfw, out0, workspace=ws, config=cfg)
    B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, out, workspace=ws, activation=B.ACT_SILU, config=cfg)
    ref = refgpu.int8_fused_dequant_silu(tq, tw, tsa, tsb, out0)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int16), ref.view(torch.int16)), int((out.view(torch.int16) != ref.view(torch.int16)).sum())


# ----------------------------------------------------------------------------- stage 2
GEMM_SHAPES = [(128, 128, 128), (128, 128, 256), (256, 256, 512), (1, 8, 16), (100, 136, 144), (5, 4096, 4096),
               (130, 264, 4096), (512, 1024, 4096), (300, 512, 11008), (257, 1280, 8192), (32, 12288, 4096),
               (512, 256, 28672)]


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_dequant_int_path_bit_exact(B, lib, oracle, cfg, M, N, K):
    """No outlier slab: int32 accumulation is exact, the epilogue is one fma + one rounding, so
    the fp16 output must equal the oracle bit for bit."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    w = rng.integers(-128, 128, (N, K), dtype=np.int8)
    sa = (rng.random(M) * 0.05 + 1e-3).astype(np.float16)
    sb = (rng.random(N) * 0.002 + 1e-4).astype(np.float16)
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    B.gemm_dequant(_t(q), _t(w), _t(sa), _t(sb), None, None, out, workspace=_gemm_ws(lib, M, N), config=cfg)
    torch.cuda.synchronize()
    ref = oracle.epilogue(oracle.igemm(q, w), sa, sb, None)
    got = out.cpu().numpy()
    bad = np.argwhere(got.view(np.uint16) != ref.view(np.uint16))
    assert bad.size == 0, f"{len(bad)} mismatches, first at {bad[:5].tolist()}: got {got[tuple(bad[0])]} want {ref[tuple(bad[0])]}"


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (100, 136, 144), (300, 520, 1040), (64, 512, 4096), (512, 1024, 4096)])
def test_gemm_dequant_with_outlier_slab(B, lib, oracle, cfg, M, N, K):
    rng = np.random.default_rng(M + N + K)
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    w = rng.integers(-128, 128, (N, K), dtype=np.int8)
    sa = (rng.random(M) * 0.05 + 1e-3).astype(np.float16)
    sb = (rng.random(N) * 0.002 + 1e-4).astype(np.float16)
    fpA = (rng.standard_normal((M, 128)) * 4).astype(np.float16)
    fpW = (rng.standard_normal((N, 128)) * 0.05).astype(np.float16)
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    B.gemm_dequant(_t(q), _t(w), _t(sa), _t(sb), _t(fpA), _t(fpW), out, workspace=_gemm_ws(lib, M, N), config=cfg)
    torch.cuda.synchronize()
    out0 = oracle.outlier_gemm(fpA, fpW)
    ref = oracle.epilogue(oracle.igemm(q, w), sa, sb, out0)
    got = out.cpu().numpy()
    assert np.isfinite(got.astype(np.float32)).all()
    mag = np.abs(fpA).astype(np.float64) @ np.abs(fpW).astype(np.float64).T
    _assert_mixed_close(got, ref, out0, "gemm+outlier vs oracle", mag)
    # pure outlier product (int part zeroed): isolates the kind::f16 accumulator
    out2 = torch.empty_like(out)
    B.gemm_dequant(_t(np.zeros_like(q)), _t(w), _t(sa), _t(sb), _t(fpA), _t(fpW), out2)
    torch.cuda.synchronize()
    ref2 = oracle.outlier_gemm(fpA, fpW)
    # fp32 accumulation of 128 exact products: error <= ~2^-20 of sum|a_j w_j| (order / tensor-core
    # alignment), plus the one fp16 rounding of the result
    mag = np.abs(fpA).astype(np.float64) @ np.abs(fpW).astype(np.float64).T
    bound = np.spacing(np.abs(ref2)).astype(np.float64) + mag * 2.0 ** -20
    assert (np.abs(out2.cpu().numpy().astype(np.float64) - ref2.astype(np.float64)) <= bound).all()


DECODE_SHAPES = [(512, 10240, 8192), (129, 1536, 4096), (512, 4096, 4096), (512, 12288, 4096), (512, 4096, 11008), (1024, 4096, 4096), (700, 11008, 4096),
                 (512, 2056, 8192), (260, 3584, 3584), (512, 8192, 1024), (1024, 28672, 1024)]


@pytest.mark.parametrize("M,N,K", DECODE_SHAPES)
def test_decode_kernel_split_schedules(B, lib, oracle, M, N, K):
    """The decode kernel's two-phase schedule (remainder tiles cut along K, int32 partial sums through the workspace,
    fp16 outlier product through the L2 scratch) on shapes that do split: the integer path is bit-exact vs the oracle,
    the mixed output is bit-identical to the whole-tile kernel (config 5, itself pinned to the oracle) and to the
    decode kernel with the split disabled (config 12); the flags re-arm themselves (second launch, garbage-filled scratch
    only cleared by the library)."""
    rng = np.random.default_rng(M + 3 * N + 5 * K)
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    w = rng.integers(-128, 128, (N, K), dtype=np.int8)
    sa = (rng.random(M) * 0.05 + 1e-3).astype(np.float16)
    sb = (rng.random(N) * 0.002 + 1e-4).astype(np.float16)
    fpA = (rng.standard_normal((M, 128)) * 4).astype(np.float16)
    fpW = (rng.standard_normal((N, 128)) * 0.05).astype(np.float16)
    tq, tw, tsa, tsb, tfa, tfw = _t(q), _t(w), _t(sa), _t(sb), _t(fpA), _t(fpW)
    ws = _gemm_ws(lib, M, N)
    outs = {}
    for cfg in (11, 11, 12, 5, 13, 0):
        o = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
        B.gemm_dequant(tq, tw, tsa, tsb, None, None, o, workspace=ws, config=cfg)
        torch.cuda.synchronize()
        outs[cfg] = o
    ref = oracle.epilogue(oracle.igemm(q, w), sa, sb, None)
    for cfg, o in outs.items():
        got = o.cpu().numpy()
        bad = np.argwhere(got.view(np.uint16) != ref.view(np.uint16))
        assert bad.size == 0, f"cfg {cfg}: {len(bad)} mismatches vs oracle, first at {bad[:5].tolist()}"
    mixed = {}
    for cfg in (11, 11, 12, 5, 13, 0):
        o = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
        B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, o, workspace=ws, config=cfg)
        torch.cuda.synchronize()
        mixed[cfg] = o
    for cfg in (12, 5, 13, 0):   # same K order of the outlier MMAs in every kernel: the mixed output is bit-identical across them
        assert torch.equal(mixed[11].view(torch.int16), mixed[cfg].view(torch.int16)), f"config {cfg} differs from config 11"
    out0 = oracle.outlier_gemm(fpA, fpW)
    mag = np.abs(fpA).astype(np.float64) @ np.abs(fpW).astype(np.float64).T
    _assert_mixed_close(mixed[11].cpu().numpy(), oracle.epilogue(oracle.igemm(q, w), sa, sb, out0), out0, "decode kernel vs oracle", mag)


SPLITK_SHAPES = [(512, 4096, 4096), (512, 4096, 11008), (1024, 4096, 4096), (300, 3584, 3584), (512, 3584, 18944), (257, 1280, 8192),
                 (512, 8192, 1024), (512, 8192, 3584), (200, 1000, 2064), (130, 520, 1040)]


@pytest.mark.parametrize("M,N,K", SPLITK_SHAPES)
def test_fat_splitk_cluster(B, lib, oracle, M, N, K):
    """Config 14: two CTA pairs of a cluster of 4 reduce half of K each and the partial sums cross shared memory
    (gemm_fat.cuh, st.async).  int32 addition is associative, so the result must equal the oracle bit for bit without the
    outlier slab, and the unsplit fat-tile kernel (config 13) bit for bit with it -- also with the fused bias / SiLU epilogue.
    The row-parallel shapes of the benchmarked models (o, down) must be served, not refused."""
    rng = np.random.default_rng(M * 3 + N + K)
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    w = rng.integers(-127, 128, (N, K), dtype=np.int8)
    sa = (rng.random(M) * 0.05 + 1e-3).astype(np.float16)
    sb = (rng.random(N) * 0.002 + 1e-4).astype(np.float16)
    fpA = (rng.standard_normal((M, 128)) * 4).astype(np.float16)
    fpW = (rng.standard_normal((N, 128)) * 0.05).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float16)
    tq, tw, tsa, tsb, tfa, tfw = _t(q), _t(w), _t(sa), _t(sb), _t(fpA), _t(fpW)
    o = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    try:
        B.gemm_dequant(tq, tw, tsa, tsb, None, None, o, config=14)
    except Exception as e:
        assert "split-K" in str(e)
        assert M > 512 or (N, K) not in ((4096, 4096), (4096, 11008), (3584, 3584), (3584, 18944)), f"split-K refused a benchmarked shape: {e}"
        pytest.skip(f"shape not served by the split-K schedule: {e}")
    torch.cuda.synchronize()
    ref = oracle.epilogue(oracle.igemm(q, w), sa, sb, None)
    bad = np.argwhere(o.cpu().numpy().view(np.uint16) != ref.view(np.uint16))
    assert bad.size == 0, f"{len(bad)} mismatches vs oracle, first at {bad[:5].tolist()}"
    a, b = torch.empty_like(o), torch.empty_like(o)
    for kw in ({}, {"bias": _t(bias)}, {"activation": B.ACT_SILU}):
        B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, a, config=14, **kw)
        B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, b, config=13, **kw)
        torch.cuda.synchronize()
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), f"{kw}: split-K differs from the unsplit kernel"
    # back-to-back launches
    for cfg in (14, 14):
        B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, a, config=cfg)
    torch.cuda.synchronize()
    B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, b, config=13)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))


# ----------------------------------------------------------------------------- whole path
def _packed(oracle, name, N, K, seed=1234):
    a = np.load(GOLD / "act_scales_l0.npz")
    scale = a[name] if name in a.files and a[name].shape[0] == K else None
    return oracle.synth_linear(N, K, scale, seed=seed)


ENQ_CASES = [  # (act-scale fixture, M, N, K)   N kept moderate so the CPU oracle stays fast
    ("Llama-2-7b/self_attn.q_proj", 1, 512, 4096),       # config 0: bs = 1 (mixed path forced for M <= 4)
    ("Llama-2-7b/self_attn.q_proj", 32, 1024, 4096),
    ("Llama-2-7b/self_attn.q_proj", 512, 768, 4096),
    ("Llama-2-7b/mlp.down_proj", 77, 512, 11008),
    ("qwen2-7b-instruct/self_attn.q_proj", 200, 512, 3584),
    ("Llama-2-70b/mlp.down_proj", 33, 256, 28672),
    ("synthetic", 130, 264, 400),
]


@pytest.mark.parametrize("name,M,N,K", ENQ_CASES)
def test_enqueue_matches_oracle_and_reference_plugin(B, oracle, name, M, N, K):
    lin = _packed(oracle, name, N, K)
    A = oracle.synth_activations(M, lin["act_scale"], seed=M)
    tA, tW, tsb, tfw, tind = _t(A), _t(lin["W8"]), _t(lin["scale_b"]), _t(lin["fp_weight"]), _t(lin["ind"])
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
    B.enqueue(tA, tW, tsb, tfw, tind, out, ws)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    r = oracle.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], return_parts=True)
    # workspace carries the stage-1 results in the reference's order: A8 | scale_a | fp_A
    a8 = ws[: M * K].view(torch.int8).view(M, K).cpu().numpy()
    if oracle.rcp_table() is not None:
        assert np.array_equal(a8, r["q"])
    mag = np.abs(r["fp_A"]).astype(np.float64) @ np.abs(lin["fp_weight"]).astype(np.float64).T
    _assert_mixed_close(got, r["out"], r["out0"], "enqueue vs oracle", mag)
    if refgpu.available() and M > 4 and K >= 256:
        ref = refgpu.enqueue(tA, tW, tsb, tfw, tind)
        torch.cuda.synchronize()
        _assert_mixed_close(got, ref.cpu().numpy(), r["out0"], "enqueue vs reference kernels", mag)


BENCH_SHAPES = [  # (act-scale fixture, M, N, K, bias): the shapes bench.py times (BASELINE.json configs[1..4]), full size
    ("Llama-2-7b/self_attn.q_proj", 65536, 12288, 4096, False),        # configs[1]: bs 32 x seq 2048, fused qkv (1.6 GB of output)
    ("Llama-2-7b/self_attn.q_proj", 512, 12288, 4096, False),          # configs[2]: decode bs 512
    ("Llama-2-7b/self_attn.o_proj", 512, 4096, 4096, False),
    ("Llama-2-7b/mlp.gate_proj", 512, 11008, 4096, False),
    ("Llama-2-7b/mlp.down_proj", 512, 4096, 11008, False),
    ("qwen2-7b-instruct/self_attn.q_proj", 65536, 4608, 3584, True),   # configs[3]: Qwen2 qkv carries a bias
    ("qwen2-7b-instruct/mlp.gate_proj", 8192, 18944, 3584, False),
    ("qwen2-7b-instruct/mlp.down_proj", 8192, 3584, 18944, False),
    ("Llama-2-70b/self_attn.q_proj", 512, 1280, 8192, False),          # configs[4]: 70B, TP 8 shards (column: N / 8)
    ("Llama-2-70b/self_attn.o_proj", 512, 8192, 1024, False),          #   row: K / 8
    ("Llama-2-70b/mlp.gate_proj", 512, 3584, 8192, False),
    ("Llama-2-70b/mlp.down_proj", 512, 8192, 3584, False),
]


@pytest.mark.parametrize("name,M,N,K,with_bias", BENCH_SHAPES)
def test_benchmarked_shapes_through_enqueue(B, oracle, name, M, N, K, with_bias):
    """mixq_enqueue at the FULL benchmarked shapes (TsinghuaMixQPlugin.cpp:518-532): quantise / extract bit-exact against the
    reference kernels over the whole batch (byte offsets beyond 2^31 at M = 65536), the whole output against the
    reference kernels on the GPU, and rows sampled from the first, a middle and the last 128-row block against the CPU
    oracle -- each within the mixed-path tolerance.  Weights are packed on the device by bench.py's packer (checked against
    the product's host packer in test_bench_host.py)."""
    sys.path.insert(0, str(ROOT))
    import bench
    z = np.load(GOLD / "act_scales_l0.npz")
    sc = z[name].astype(np.float32) if name in z.files and z[name].shape[0] == K else oracle.synth_act_scale(K, seed=7)
    if sc.shape[0] != K:      # a TP row shard reads a K slice of the activation
        sc = sc[:K]
    g = torch.Generator(device=DEV).manual_seed(1234 + N + K)
    sct = torch.from_numpy(sc).to(DEV)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.02).half()
    W8, sb, fw, ind = bench.pack_gpu(torch, W, sct)
    del W
    A = (torch.randn(M, K, device=DEV, generator=g) * (sct[None, :] / 3.0)).half()
    bias = (torch.randn(N, device=DEV, generator=g)).half() if with_bias else None
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
    B.enqueue(A, W8, sb, fw, ind, out, ws, bias=bias)
    torch.cuda.synchronize()
    # stage 1 against the reference kernels, whole batch
    if refgpu.available():
        A8 = torch.empty(M, K, dtype=torch.int8, device=DEV)
        sa = torch.empty(M, dtype=torch.float16, device=DEV)
        fpA = torch.empty(M, 128, dtype=torch.float16, device=DEV)
        B.quant_extract(A, ind, A8, sa, fpA)
        rq, rsa = refgpu.int8quant(A)
        assert torch.equal(rq, A8) and torch.equal(rsa.view(torch.int16), sa.view(torch.int16))
        assert torch.equal(refgpu.extract(A, ind).view(torch.int16), fpA.view(torch.int16))
        del rq, A8
        ref = refgpu.enqueue(A, W8, sb, fw, ind)
        if bias is not None:
            ref = (ref.float() + bias.float()[None, :]).half()      # plugin.py:158-160: bias added after the plugin
        for lo in range(0, M, 8192):                                # the bound needs M x N floats: walk row bands
            hi = min(M, lo + 8192)
            if bias is None:
                ok, worst, rel = bench.mixed_close(torch, out[lo:hi], ref[lo:hi], A[lo:hi], fw, ind)
                assert ok, (lo, worst, rel)
            else:   # the bias add rounds once more: compare the biased outputs within one more ulp
                d = (out[lo:hi].float() - ref[lo:hi].float()).abs()
                ulp = torch.exp2(torch.floor(torch.log2(ref[lo:hi].float().abs().clamp_min(2.0 ** -14))) - 10)
                fa = A[lo:hi][:, ind.long()].float()
                o0 = (fa @ fw.float().t()).half().float().abs().clamp_min(2.0 ** -14)
                bound = 2 * ulp + torch.exp2(torch.floor(torch.log2(o0)) - 10) + (fa.abs() @ fw.float().abs().t()) * 2.0 ** -20
                assert bool((d <= bound).all()), float((d / bound).max())
        del ref
    # sampled rows against the CPU oracle
    rows = np.unique(np.concatenate([np.arange(0, 8), np.arange(M // 2 // 128 * 128, M // 2 // 128 * 128 + 8) % M, np.arange(M - 8, M)]))
    rt = torch.from_numpy(rows).to(DEV)
    A_s = A[rt].cpu().numpy()
    r = oracle.forward(A_s, W8.cpu().numpy(), sb.cpu().numpy(), fw.cpu().numpy(), ind.cpu().numpy(), return_parts=True)
    want = r["out"] if bias is None else oracle.epilogue_ex(r["acc"], r["sa"], sb.cpu().numpy(), r["out0"], bias=bias.cpu().numpy())
    got = out[rt].cpu().numpy()
    if bias is None:
        mag = np.abs(r["fp_A"]).astype(np.float64) @ np.abs(fw.cpu().numpy()).astype(np.float64).T
        _assert_mixed_close(got, want, r["out0"], f"{name} {M}x{N}x{K} sampled rows vs oracle", mag)
    else:
        ulp = lambda x: np.spacing(np.abs(x).astype(np.float16)).astype(np.float32)
        bound = 2 * ulp(want.astype(np.float32)) + ulp(r["out0"].astype(np.float32)) + ulp(r["out"].astype(np.float32))
        assert (np.abs(got.astype(np.float32) - want.astype(np.float32)) <= bound).all()


def test_enqueue_workspace_and_errors(B, lib, oracle):
    lin = _packed(oracle, "synthetic", 64, 256)
    A = oracle.synth_activations(16, lin["act_scale"])
    tA, tW, tsb, tfw, tind = _t(A), _t(lin["W8"]), _t(lin["scale_b"]), _t(lin["fp_weight"]), _t(lin["ind"])
    out = torch.empty(16, 64, dtype=torch.float16, device=DEV)
    small = torch.empty(64, dtype=torch.uint8, device=DEV)
    with pytest.raises(B.MixQError, match="workspace"):
        B.enqueue(tA, tW, tsb, tfw, tind, out, small)
    # unaligned workspace base is realigned like nextWorkspacePtr (TsinghuaMixQPlugin.cpp:206-215)
    ws = torch.empty(B.workspace_size(16, 64, 256) + 64, dtype=torch.uint8, device=DEV)
    B.enqueue(tA, tW, tsb, tfw, tind, out, ws[3:])
    torch.cuda.synchronize()
    r = oracle.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], return_parts=True)
    _assert_mixed_close(out.cpu().numpy(), r["out"], r["out0"], "unaligned workspace")


def test_enqueue_is_graph_capturable_and_async(B, oracle):
    """enqueue must not sync or allocate: capture it in a CUDA graph and replay (SURVEY 8b threading/stream row)."""
    lin = _packed(oracle, "synthetic", 256, 512)
    A = oracle.synth_activations(96, lin["act_scale"])
    tA, tW, tsb, tfw, tind = _t(A), _t(lin["W8"]), _t(lin["scale_b"]), _t(lin["fp_weight"]), _t(lin["ind"])
    out = torch.zeros(96, 256, dtype=torch.float16, device=DEV)
    ws = torch.empty(B.workspace_size(96, 256, 512), dtype=torch.uint8, device=DEV)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        B.enqueue(tA, tW, tsb, tfw, tind, out, ws, stream=s)   # warm-up (first-use attribute sets)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        out.zero_()
        with torch.cuda.graph(g, stream=s):
            B.enqueue(tA, tW, tsb, tfw, tind, out, ws, stream=s)
        g.replay()
        g.replay()
    torch.cuda.synchronize()
    r = oracle.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], return_parts=True)
    _assert_mixed_close(out.cpu().numpy(), r["out"], r["out0"], "graph replay")


@pytest.mark.parametrize("M", [7, 2048, 5000])
def test_linear_host_pipeline(B, lib, oracle, M):
    """mixq_linear_host (host buffers in, host buffers out; row slabs pipelined over three streams) must return
    exactly what mixq_enqueue writes for the same activations."""
    N, K = 264, 400
    lin = _packed(oracle, "synthetic", N, K)
    A = oracle.synth_activations(M, lin["act_scale"], seed=M)
    tW, tsb, tfw, tind = _t(lin["W8"]), _t(lin["scale_b"]), _t(lin["fp_weight"]), _t(lin["ind"])
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
    B.enqueue(_t(A), tW, tsb, tfw, tind, out, ws)
    hA = torch.from_numpy(A).pin_memory()
    hO = torch.full((M, N), float("nan"), dtype=torch.float16).pin_memory()
    scratch = torch.empty(lib.mixq_host_scratch_size(M, N, K), dtype=torch.uint8, device=DEV)
    t = B.make_tensors(None, tW, tsb, tfw, tind, None)
    for _ in range(2):       # second call reuses the pipe's streams/events
        hO.fill_(float("nan"))
        B.check(lib.mixq_linear_host(ctypes.byref(t), hA.data_ptr(), hO.data_ptr(), M, N, K, scratch.data_ptr(),
                                     scratch.numel(), 0, torch.cuda.current_stream().cuda_stream), "mixq_linear_host")
        assert torch.equal(hO.view(torch.int16), out.cpu().view(torch.int16))
    small = torch.empty(1024, dtype=torch.uint8, device=DEV)
    assert lib.mixq_linear_host(ctypes.byref(t), hA.data_ptr(), hO.data_ptr(), M, N, K, small.data_ptr(), 1024, 0, None) == -2


@pytest.mark.parametrize("parts", [1, 2, 4])
def test_host_calls_async_sequence(B, lib, oracle, parts):
    """MIXQ_FLAG_HOST_ASYNC: a chain of host-buffer calls of different shapes (plain, shared-input, gated) queued without
    waiting, one scratch for all of them (single size: a call waits for the previous download; k times the size: the
    calls take the k parts in turn), mixq_host_drain at the end.  Every result must equal the synchronous call's, bit for bit; a
    synchronous call issued while asynchronous ones are pending drains by itself."""
    # long downloads next to short ones and activations of different widths, so that a call's upload or kernels running
    # before the previous user of its scratch half has finished would be seen
    M = 2048
    shapes = [(4096, 256), (264, 1024), (2048, 512)]
    N0, K0 = shapes[0]
    lins = [_packed(oracle, "synthetic", N, K, seed=11 + i) for i, (N, K) in enumerate(shapes)]
    dev = [tuple(_t(l[k]) for k in ("W8", "scale_b", "fp_weight", "ind")) for l in lins]
    tabs = [B.make_tensors(None, *d, None) for d in dev]
    up = _packed(oracle, "synthetic", N0, K0, seed=29)       # second projection of the gated call: same shape as lins[0]
    up["ind"] = lins[0]["ind"]
    dup = tuple(_t(up[k]) for k in ("W8", "scale_b", "fp_weight")) + (dev[0][3],)
    tab_up = B.make_tensors(None, *dup, None)
    hA = [torch.from_numpy(oracle.synth_activations(M, l["act_scale"], seed=70 + i)).pin_memory() for i, l in enumerate(lins)]
    need = max([B.linears_host_scratch_size(M, [N], K) for N, K in shapes] + [B.gated_host_scratch_size(M, N0, K0),
                                                                             B.linears_host_scratch_size(M, [N0, N0], K0)])
    scratch = torch.empty(parts * need + (512 if parts > 1 else 0), dtype=torch.uint8, device=DEV)

    def run(flags, reps):
        outs = []
        for rep in range(reps):
            for i in range(3):
                o = torch.full((M, shapes[i][0]), float("nan"), dtype=torch.float16).pin_memory()
                B.linears_host([tabs[i]], hA[i], [o], scratch, flags=flags)
                outs.append(o)
            o1 = torch.full((M, N0), float("nan"), dtype=torch.float16).pin_memory()
            o2 = torch.full((M, N0), float("nan"), dtype=torch.float16).pin_memory()
            B.linears_host([tabs[0], tab_up], hA[0], [o1, o2], scratch, flags=flags)
            og = torch.full((M, N0), float("nan"), dtype=torch.float16).pin_memory()
            B.gated_host(tabs[0], tab_up, hA[0], og, scratch, flags=flags)
            outs += [o1, o2, og]
        return outs

    want = run(0, 1)
    got = run(B.FLAG_HOST_ASYNC, 3)
    B.host_drain()
    for j, o in enumerate(got):
        assert torch.equal(o.view(torch.int16), want[j % len(want)].view(torch.int16)), j
    # pending asynchronous calls followed by a synchronous one on the same scratch
    a = run(B.FLAG_HOST_ASYNC, 1)
    o = torch.full((M, shapes[1][0]), float("nan"), dtype=torch.float16).pin_memory()
    B.linears_host([tabs[1]], hA[1], [o], scratch)
    assert torch.equal(o.view(torch.int16), want[1].view(torch.int16))
    B.host_drain()
    for j, x in enumerate(a):
        assert torch.equal(x.view(torch.int16), want[j].view(torch.int16)), j


def test_python_plugin_mirror(B, oracle):
    """MixQLinear / mixgemm (reference plugin.py) end to end, 3-D activations, bias outside the plugin."""
    from mixq_tensorrt_llm_b200.plugin import MixQLinear
    lin = _packed(oracle, "Llama-2-7b/self_attn.q_proj", 384, 4096)
    mod = MixQLinear(4096, 384, bias=True, device=DEV)
    bias = (np.arange(384) % 7 - 3).astype(np.float16)
    mod.load_packed(_t(lin["W8"]), _t(lin["scale_b"]), _t(lin["fp_weight"]), _t(lin["ind"]), _t(bias))
    assert mod.weight.shape == (384, 2048) and mod.fp_ind.shape == (256,) and mod.weight.dtype == torch.float16
    A = oracle.synth_activations(6 * 5, lin["act_scale"]).reshape(6, 5, 4096)
    y = mod(_t(A))
    torch.cuda.synchronize()
    assert y.shape == (6, 5, 384)
    r = oracle.forward(A.reshape(30, 4096), lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], return_parts=True)
    mod.bias = None                    # bias is added outside the plugin, in fp16 (plugin.py:158-160)
    y0 = mod(_t(A))
    torch.cuda.synchronize()
    _assert_mixed_close(y0.cpu().numpy().reshape(30, 384), r["out"], r["out0"], "MixQLinear")
    assert torch.equal(y, y0 + _t(bias))


# ----------------------------------------------------------------------------- full-size properties
def test_full_size_properties(B, oracle):
    """Llama-2-7B qkv shape (N=12288, K=4096) at M=4096: size-independent properties + sampled rows vs oracle."""
    M, N, K = 4096, 12288, 4096
    lin = _packed(oracle, "Llama-2-7b/self_attn.q_proj", N, K)
    A = oracle.synth_activations(M, lin["act_scale"], seed=5)
    tA, tW, tsb, tfw, tind = _t(A), _t(lin["W8"]), _t(lin["scale_b"]), _t(lin["fp_weight"]), _t(lin["ind"])
    ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    B.enqueue(tA, tW, tsb, tfw, tind, out, ws)
    torch.cuda.synchronize()
    # (1) sampled rows against the oracle
    rows = np.random.default_rng(0).choice(M, 24, replace=False)
    r = oracle.forward(A[rows], lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], return_parts=True)
    got = out[_t(rows)].cpu().numpy()
    _assert_mixed_close(got, r["out"], r["out0"], "sampled rows")
    # (2) token permutation equivariance, bit-exact (per-token quantisation, no cross-row coupling)
    perm = torch.randperm(M, device=DEV)
    out_p = torch.empty_like(out)
    B.enqueue(tA[perm].contiguous(), tW, tsb, tfw, tind, out_p, ws)
    torch.cuda.synchronize()
    assert torch.equal(out_p.view(torch.int16), out[perm].view(torch.int16))
    # (3) output-channel slicing: a column-parallel shard reproduces its slice bit for bit
    n0, n1 = 4096, 4096 + 1280
    out_s = torch.empty(M, n1 - n0, dtype=torch.float16, device=DEV)
    B.enqueue(tA, tW[n0:n1].contiguous(), tsb[n0:n1].contiguous(), tfw[n0:n1].contiguous(), tind, out_s, ws)
    torch.cuda.synchronize()
    assert torch.equal(out_s.view(torch.int16), out[:, n0:n1].contiguous().view(torch.int16))
    # (4) scaling the activations by a power of two scales the per-token scale, not the codes
    out_2 = torch.empty_like(out)
    B.enqueue((tA * 2).contiguous(), tW, tsb, tfw, tind, out_2, ws)
    torch.cuda.synchronize()
    # (exact wherever the fp16 result is a normal number: subnormal rounding is not scale invariant; the
    #  same holds for the rare outlier products that are themselves fp16-subnormal)
    normal = out.abs() >= 2.0 ** -14
    neq = (out_2.view(torch.int16) != (out * 2).view(torch.int16)) & normal
    assert neq.float().mean().item() < 1e-5
    # (5) deterministic: same call, same bits
    out_r = torch.empty_like(out)
    B.enqueue(tA, tW, tsb, tfw, tind, out_r, ws)
    torch.cuda.synchronize()
    assert torch.equal(out_r.view(torch.int16), out.view(torch.int16))


# ------------------------------------------------------------------ fused row-parallel GEMM + all-reduce
@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4])
def test_fused_allreduce_virtual_ranks(world):
    """mixq_gemm_dequant_allreduce, protocol + arithmetic on one GPU: `world` concurrent launches (virtual ranks, each
    confined to 148/world SMs, peer pointers = local pointers).  Every rank's Out must equal
    fp16(sum_r fp32(partial_r)) in rank order BIT FOR BIT, across repeated launches (counter re-arming, epoch parity)
    and shapes with M/N edges.  Runs in a subprocess: a protocol bug traps that context instead of hanging pytest."""
    import subprocess
    import sys as _sys
    import os
    for env in ({}, {"MIXQ_PULL_TWO_SHOT_MAX_MB": "32"}):        # default selection; then with the opt-in two-shot pull enabled
        r = subprocess.run([_sys.executable, str(ROOT / "tests" / "gpu_ar_virtual.py"), str(world),
                            "512x4096x1024", "300x1000x512", "2048x4096x2048", "40x256x256"],
                           capture_output=True, text=True, timeout=300, env={**os.environ, **env})
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NVLink peer memory)")
def test_row_parallel_allreduce_on_real_gpus():
    """The same check over real peer memory when the box has more than one GPU (reference site: plugin.py:155-156): both
    exchange paths against fp16(sum_r fp32(partial_r)) bit for bit on min(4, #GPUs) ranks, the MixQLinear row-parallel module
    against plugin + NCCL, and a CUDA-graph replay (tests/gpu_tp_fused.py prints PASS)."""
    import subprocess
    import sys as _sys
    n = min(4, torch.cuda.device_count())
    n = 2 if n == 3 else n
    r = subprocess.run([_sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(ROOT / "tests" / "gpu_tp_fused.py"), "512x4096x4096", "32x4096x4096", "2048x4096x2048"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


# ------------------------------------------------------------------ M <= 4 branch: weight-only GEMV
def _gemv_cases():
    import sys
    sys.path.insert(0, str(GOLD))
    import make_gemv_golden as G
    return G, list(G.GEMV_CASES) + [(4, 4096, 4096, 21), (1, 12288, 4096, 22), (3, 4096, 11008, 23)]


@pytest.mark.parametrize("case", range(10))
def test_gemv_w8a16_bit_exact(B, oracle, case):
    """mixq_gemv_w8a16 vs the CPU oracle (which is pinned to the reference kernel's outputs, tests/golden/
    ref_gemv_b200.npz) and, when oracle/_ref/libref_gemv.so is on the box, vs the reference kernel itself:
    fp16 output BIT-EXACT (the kernel keeps the reference's fp16 chains and fp32 reduction tree)."""
    G, cases = _gemv_cases()
    M, N, K, seed = cases[case]
    A, W_t = G.gemv_case(M, N, K, seed)
    qw, sc = oracle.eetq_quant_weights(W_t)
    dA, dq, ds = _t(A), _t(qw), _t(sc)
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    B.gemv_w8a16(dA, dq, ds, out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    if N * K <= 4096 * 4096:                    # the scalar oracle takes seconds beyond that
        want = oracle.gemv_w8a16(A, qw, sc)
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), (M, N, K)
    if refgpu.gemv_available():
        ref = refgpu.gemv(dA, dq, ds).cpu().numpy()
        assert np.array_equal(got.view(np.uint16), ref.view(np.uint16)), (M, N, K)
    # sanity against float64 on the dequantised weights
    f64 = A.astype(np.float64) @ (oracle.eetq_unprocess(qw).astype(np.float64) * sc.astype(np.float64)[None, :])
    assert np.abs(got.astype(np.float64) - f64).max() <= 4e-3 * np.abs(f64).max() + 2e-3


def test_enqueue_takes_weight_only_branch_for_small_M(B, oracle):
    """TsinghuaMixQPlugin.cpp:472: M <= 4 -> weight-only GEMV over q_weight with input 6 as the scales; M > 4, a NULL
    q_weight or MIXQ_FLAG_FORCE_MIXED -> the mixed path."""
    N, K = 512, 4096
    lin = oracle.synth_linear(N, K, oracle.load_act_scales("Llama-2-7b/self_attn.q_proj"))
    rng = np.random.default_rng(9)
    W_t = (rng.standard_normal((K, N)) * 0.02).astype(np.float16)
    qw, sc = oracle.eetq_quant_weights(W_t)
    dq, ds = _t(qw), _t(sc)
    W8, sb, fw, ind = (_t(lin[k]) for k in ("W8", "scale_b", "fp_weight", "ind"))
    for M in (1, 4, 5):
        A = oracle.synth_activations(M, lin["act_scale"], seed=30 + M)
        dA = _t(A)
        ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
        mixed = torch.empty(M, N, dtype=torch.float16, device=DEV)
        B.enqueue(dA, W8, sb, fw, ind, mixed, ws)                                   # q_weight NULL -> mixed
        out = torch.empty(M, N, dtype=torch.float16, device=DEV)
        B.enqueue(dA, W8, sb, fw, ind, out, ws, q_weight=dq, scaling_factors=ds)
        forced = torch.empty(M, N, dtype=torch.float16, device=DEV)
        B.enqueue(dA, W8, sb, fw, ind, forced, ws, flags=B.FLAG_FORCE_MIXED, q_weight=dq, scaling_factors=ds)
        torch.cuda.synchronize()
        assert torch.equal(forced.view(torch.int16), mixed.view(torch.int16))
        if M <= 4:
            want = oracle.gemv_w8a16(A, qw, sc)
            assert np.array_equal(out.cpu().numpy().view(np.uint16), want.view(np.uint16))
        else:
            assert torch.equal(out.view(torch.int16), mixed.view(torch.int16))


def test_gemv_w8a16_rejects_bad_shapes(B, lib):
    a = torch.zeros(5, 64, dtype=torch.float16, device=DEV)
    q = torch.zeros(64, 8, dtype=torch.int8, device=DEV)
    s = torch.zeros(8, dtype=torch.float16, device=DEV)
    o = torch.zeros(5, 8, dtype=torch.float16, device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.mixq_gemv_w8a16(p(a), p(q), p(s), p(o), 5, 8, 64, None) == -4     # M > 4
    assert lib.mixq_gemv_w8a16(p(a), p(q), p(s), p(o), 2, 6, 64, None) == -4     # N % 4
    assert lib.mixq_gemv_w8a16(p(a), p(q), p(s), p(o), 2, 8, 32, None) == -4     # K % 64
    assert lib.mixq_gemv_w8a16(None, p(q), p(s), p(o), 2, 8, 64, None) == -1
    assert lib.mixq_gemv_w8a16(p(a), p(q), p(s), p(o), 0, 8, 64, None) == 0


def test_plugin_module_takes_weight_only_branch_with_qweight(B, oracle):
    """MixQLinear -> C plugin handle -> MixQPlugin::enqueue: with qweight loaded a call with M <= 4 goes through the
    weight-only GEMV (scaled by weights_scaling_factor = plugin input 6, as plugin.py:149 wires it); M = 5 is mixed."""
    from mixq_tensorrt_llm_b200.plugin import MixQLinear
    N, K = 512, 4096
    lin = oracle.synth_linear(N, K, oracle.load_act_scales("Llama-2-7b/self_attn.q_proj"))
    rng = np.random.default_rng(4)
    W_t = (rng.standard_normal((K, N)) * 0.02).astype(np.float16)
    qw, _ = oracle.eetq_quant_weights(W_t)
    mod = MixQLinear(K, N, device=DEV)
    mod.load_packed(*(_t(lin[k]) for k in ("W8", "scale_b", "fp_weight", "ind")), qweight=_t(qw))
    A = oracle.synth_activations(2, lin["act_scale"], seed=77)
    y = mod(_t(A))
    torch.cuda.synchronize()
    want = oracle.gemv_w8a16(A, qw, lin["scale_b"])
    assert np.array_equal(y.cpu().numpy().view(np.uint16), want.view(np.uint16))
    A5 = oracle.synth_activations(5, lin["act_scale"], seed=78)
    y5 = mod(_t(A5)).cpu().numpy()
    r = oracle.forward(A5, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], return_parts=True)
    _assert_mixed_close(y5, r["out"], r["out0"], "M=5 mixed")


# ------------------------------------------------------------------ fused epilogue: bias / SiLU
@pytest.mark.parametrize("cfg", [0, 1, 5, 6, 9, 10, 11, 13, 16])
def test_fused_bias_epilogue_bit_exact(B, lib, oracle, cfg):
    """mixq_gemm_dequant_ex with a bias and no outlier slab: Out == fp16(float(fp16(fma)) + bias[n]) bit for bit
    (int32 exact, one FMA, two roundings) for every kernel family."""
    M, N, K = 300, 520, 1024
    rng = np.random.default_rng(11)
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    W8 = rng.integers(-127, 128, (N, K), dtype=np.int8)
    sa = (rng.random(M) * 0.02 + 1e-3).astype(np.float16)
    sb = (rng.random(N) * 2e-3 + 1e-4).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float16)
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    B.gemm_dequant(_t(q), _t(W8), _t(sa), _t(sb), None, None, out, bias=_t(bias), config=cfg,
                   workspace=_gemm_ws(lib, M, N) if cfg in (11, 12) else None)
    torch.cuda.synchronize()
    want = oracle.epilogue_ex(oracle.igemm(q, W8), sa, sb, None, bias=bias)
    assert np.array_equal(out.cpu().numpy().view(np.uint16), want.view(np.uint16))


@pytest.mark.parametrize("M", [3, 64, 700])
def test_fused_silu_bias_epilogue(B, oracle, M):
    """mixq_enqueue_ex (SiLU + bias) against the oracle: the activation runs in fp32 before the rounding with the
    fast-math intrinsics the reference build uses, so the tolerance is 2 fp16 ulp of the result on top of the
    mixed-path tolerance (1 ulp of the outlier product; silu' <= 1.1); M = 3 with q_weight also covers the GEMV."""
    N, K = 512, 4096
    lin = oracle.synth_linear(N, K, oracle.load_act_scales("Llama-2-7b/self_attn.q_proj"))
    A = oracle.synth_activations(M, lin["act_scale"], seed=50 + M)
    bias = np.random.default_rng(1).standard_normal(N).astype(np.float16)
    ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    W8, sb, fw, ind = (_t(lin[k]) for k in ("W8", "scale_b", "fp_weight", "ind"))
    B.enqueue(_t(A), W8, sb, fw, ind, out, ws, bias=_t(bias), activation=B.ACT_SILU)
    torch.cuda.synchronize()
    r = oracle.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"], return_parts=True)
    want = oracle.epilogue_ex(r["acc"], r["sa"], lin["scale_b"], r["out0"], bias=bias, silu=True)
    got = out.cpu().numpy().astype(np.float32)
    w32 = want.astype(np.float32)
    ulp = lambda x: np.spacing(np.abs(x).astype(np.float16)).astype(np.float32)
    bound = 2 * ulp(w32) + 1.2 * ulp(r["out0"].astype(np.float32)) + 2 * ulp((want.astype(np.float32) - bias.astype(np.float32)[None, :]))
    assert (np.abs(got - w32) <= bound).all(), float((np.abs(got - w32) / bound).max())
    if M <= 4:
        Wt = (np.random.default_rng(6).standard_normal((K, N)) * 0.02).astype(np.float16)
        qw, sc = oracle.eetq_quant_weights(Wt)
        B.enqueue(_t(A), W8, sb, fw, ind, out, ws, q_weight=_t(qw), scaling_factors=_t(sc), bias=_t(bias), activation=B.ACT_SILU)
        torch.cuda.synchronize()
        plain = torch.empty(M, N, dtype=torch.float16, device=DEV)
        B.enqueue(_t(A), W8, sb, fw, ind, plain, ws, q_weight=_t(qw), scaling_factors=_t(sc))
        torch.cuda.synchronize()
        x = plain.cpu().numpy().astype(np.float64)
        ref = ((x / (1 + np.exp(-x))).astype(np.float16).astype(np.float32) + bias.astype(np.float32)[None, :])
        assert np.abs(out.cpu().numpy().astype(np.float32) - ref).max() <= 3e-2


# ------------------------------------------------------------------ gated MLP input half: gate/up fusion
@pytest.mark.parametrize("M,N,K", [(512, 11008, 4096), (32, 1376, 4096), (129, 2752, 1024), (1024, 3584, 8192), (300, 18944, 3584),
                                   (77, 264, 512), (1100, 1376, 4096)])
def test_gated_mlp_half(B, oracle, M, N, K):
    """mixq_enqueue_gated == fp16(silu(gate(x))) * fp16(up(x)) (MixLlamaMLP.forward, fused/mlp.py:57-70):
    bit-identical to the composition of the library's own two calls (mixq_enqueue_ex with SiLU, mixq_enqueue) and an fp16
    multiply, and within the propagated mixed-path tolerance of the oracle's restatement."""
    sc = oracle.synth_act_scale(K, seed=K + N)
    gate, up = oracle.synth_linear(N, K, sc, seed=11), oracle.synth_linear(N, K, sc, seed=12)
    assert np.array_equal(gate["ind"], up["ind"])
    A = oracle.synth_activations(M, sc, seed=M)
    tA, ind = _t(A), _t(gate["ind"])
    tg = tuple(_t(gate[k]) for k in ("W8", "scale_b", "fp_weight"))
    tu = tuple(_t(up[k]) for k in ("W8", "scale_b", "fp_weight"))
    ws = torch.empty(B.gated_workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    n0 = B.load().mixq_launch_count()
    B.enqueue_gated(tA, tg, tu, ind, out, ws)
    torch.cuda.synchronize()
    assert B.load().mixq_launch_count() - n0 == (2 if M <= 1024 else 4)
    # the library's own unfused sequence: same kernels' arithmetic, so the fused call must agree bit for bit
    g = torch.empty(M, N, dtype=torch.float16, device=DEV)
    u = torch.empty(M, N, dtype=torch.float16, device=DEV)
    B.enqueue(tA, *tg, ind, g, ws, activation=B.ACT_SILU)
    B.enqueue(tA, *tu, ind, u, ws)
    comp = g * u
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int16), comp.view(torch.int16)), int((out.view(torch.int16) != comp.view(torch.int16)).sum())
    unf = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    B.enqueue_gated(tA, tg, tu, ind, unf, ws, config=100)      # the two-GEMM + multiply composition inside the library
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int16), unf.view(torch.int16))
    if M <= 1024:                                              # the fat tile with 12 epilogue warps (default for the gated call: 8)
        unf.fill_(float("nan"))
        B.enqueue_gated(tA, tg, tu, ind, unf, ws, config=13)
        torch.cuda.synchronize()
        assert torch.equal(out.view(torch.int16), unf.view(torch.int16))
    # oracle
    rg = oracle.forward(A, gate["W8"], gate["scale_b"], gate["fp_weight"], gate["ind"], return_parts=True)
    ru = oracle.forward(A, up["W8"], up["scale_b"], up["fp_weight"], up["ind"], return_parts=True)
    want = oracle.gated_mlp_half(A, gate, up)
    gs = oracle.epilogue_ex(rg["acc"], rg["sa"], gate["scale_b"], rg["out0"], silu=True).astype(np.float32)
    uo = ru["out"].astype(np.float32)
    ulp = lambda x: np.spacing(np.abs(x).astype(np.float16)).astype(np.float32)
    bg = 2 * ulp(gs) + 1.2 * ulp(rg["out0"].astype(np.float32))          # SiLU path (test_fused_silu_bias_epilogue)
    bu = ulp(uo) + ulp(ru["out0"].astype(np.float32))                     # plain mixed path
    bound = np.abs(uo) * bg + np.abs(gs) * bu + bg * bu + ulp(want.astype(np.float32))
    got = out.cpu().numpy().astype(np.float32)
    d = np.abs(got - want.astype(np.float32))
    fin = np.isfinite(want.astype(np.float32))
    assert np.array_equal(np.isfinite(got), fin)
    assert (d[fin] <= bound[fin]).all(), float((d[fin] / bound[fin]).max())
    rel = np.linalg.norm(d[fin].astype(np.float64)) / max(np.linalg.norm(want.astype(np.float64)[fin]), 1e-30)
    assert rel <= 2e-3, rel


def test_gated_stage2_no_outliers_bit_exact(B, lib, oracle):
    """mixq_gemm_dequant_gated without the outlier slabs: int32 exact, one FMA per projection, so the only freedom left is
    the fast-math SiLU -- compare with the oracle on the up projection's side exactly (gate weights chosen so silu(g) = 1
    is not available; instead check against the library's unfused stage-2 calls bit for bit)."""
    rng = np.random.default_rng(3)
    M, N, K = 384, 1200, 2048
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    wg, wu = rng.integers(-127, 128, (N, K), dtype=np.int8), rng.integers(-127, 128, (N, K), dtype=np.int8)
    sa = (rng.random(M) * 0.01 + 1e-3).astype(np.float16)
    sg, su = (rng.random(N) * 1e-3 + 1e-4).astype(np.float16), (rng.random(N) * 1e-3 + 1e-4).astype(np.float16)
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    B.gemm_dequant_gated(_t(q), _t(sa), None, (_t(wg), _t(sg), None), (_t(wu), _t(su), None), out)
    g = torch.empty(M, N, dtype=torch.float16, device=DEV)
    u = torch.empty(M, N, dtype=torch.float16, device=DEV)
    B.gemm_dequant(_t(q), _t(wg), _t(sa), _t(sg), None, None, g, activation=B.ACT_SILU)
    B.gemm_dequant(_t(q), _t(wu), _t(sa), _t(su), None, None, u)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int16), (g * u).view(torch.int16))
    want_u = oracle.epilogue(oracle.igemm(q, wu), sa, su, None)
    assert np.array_equal(u.cpu().numpy().view(np.uint16), want_u.view(np.uint16))
    x = oracle.epilogue_ex(oracle.igemm(q, wg), sa, sg, None, silu=True).astype(np.float32)
    want = (x * want_u.astype(np.float32)).astype(np.float16).astype(np.float32)
    got = out.cpu().numpy().astype(np.float32)
    assert (np.abs(got - want) <= 2 * np.spacing(np.abs(x).astype(np.float16)).astype(np.float32) * np.abs(want_u.astype(np.float32))
            + np.spacing(np.abs(want).astype(np.float16)).astype(np.float32)).all()


def test_module_fused_bias_equals_unfused(B, oracle):
    """MixQLinear(bias=True): forward(fuse_bias=True) == forward() (plugin, then x + bias as plugin.py:158-160) bit for bit."""
    from mixq_tensorrt_llm_b200.plugin import MixQLinear
    N, K, M = 512, 4096, 96
    lin = oracle.synth_linear(N, K, oracle.load_act_scales("Llama-2-7b/self_attn.q_proj"))
    bias = torch.randn(N, device=DEV).half()
    mod = MixQLinear(K, N, bias=True, device=DEV)
    mod.load_packed(*(_t(lin[k]) for k in ("W8", "scale_b", "fp_weight", "ind")), bias=bias)
    A = _t(oracle.synth_activations(M, lin["act_scale"], seed=5))
    a, b = mod(A), mod(A, fuse_bias=True)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))


# ------------------------------------------------------------------ configs[0]: the MixQ/src torch path (dynamic outliers)
@pytest.mark.parametrize("M", [1, 32])
def test_config0_mixsrc_forward(B, oracle, M):
    """BASELINE.json configs[0]: one 4096 x 4096 Llama-7B linear at bs = 1 "via MixQ/src torch reference": weights quantised
    WITHOUT outlier handling (linear.py:110-118), outlier columns found at run time above sigma = 6 (:197-223), masked in the
    activations (MIXQ_FLAG_MASK_OUTLIERS) and multiplied with the dequantised codes.  MixQSrcLinear (the product's mirror of
    MixLinear_GEMM) against the oracle's restatement over three calls: before, while and after the outlier set grows."""
    from mixq_tensorrt_llm_b200.plugin import MixQSrcLinear
    N = K = 4096
    rng = np.random.default_rng(40 + M)
    W = (rng.standard_normal((N, K)) * 0.02).astype(np.float16)
    sc = oracle.load_act_scales("Llama-2-7b/self_attn.q_proj")
    st = oracle.mixsrc_init(W)
    mod = MixQSrcLinear(_t(W))
    assert np.array_equal(mod.q_weight.cpu().numpy(), st["q_weight"]) and np.array_equal(mod.scale_col.cpu().numpy().view(np.uint16), st["scale_col"].view(np.uint16))
    calm = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)                     # nothing above sigma
    # layer-0 channel statistics, scaled so that a dozen (bs 1) to a few dozen (bs 32) columns hold values above sigma = 6
    hot = (oracle.synth_activations(M, sc, seed=77 + M).astype(np.float32) * 6).astype(np.float16)
    hot2 = (oracle.synth_activations(M, sc, seed=99 + M).astype(np.float32) * 6).astype(np.float16)
    for step, x in enumerate((calm, hot, hot2)):
        want = oracle.mixsrc_forward(st, x)
        got = mod(_t(x)).cpu().numpy()
        assert mod.ind.cpu().numpy().tolist() == st["ind"].tolist(), step
        assert mod.add_outliers == st["add_outliers"]
        n = st["ind"].size
        if n == 0:
            assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), step     # INT8 part only: bit-exact
        else:
            assert 0 < n <= 128
            acts = x[:, st["ind"]]
            out0 = oracle.outlier_gemm(acts, st["weight_cache"])
            mag = np.abs(acts).astype(np.float64) @ np.abs(st["weight_cache"]).astype(np.float64).T
            _assert_mixed_close(got, want, out0, f"MixQ/src forward, call {step}", mag)
    assert st["ind"].size > 0 and not st["add_outliers"]


def test_llama_mlp_module(B, oracle):
    """MixQLlamaMLP (reference fused/mlp.py:36-70) == down(silu(gate(x)) * up(x)) built from the three MixQLinear modules'
    own forward calls, bit for bit."""
    from mixq_tensorrt_llm_b200.plugin import MixQLinear, MixQLlamaMLP
    H, F, M = 1024, 2752, 200
    sc_h, sc_f = oracle.synth_act_scale(H, seed=1), oracle.synth_act_scale(F, seed=2)
    mods = {}
    for name, (n, k, sc, seed) in dict(gate=(F, H, sc_h, 3), up=(F, H, sc_h, 4), down=(H, F, sc_f, 5)).items():
        lin = oracle.synth_linear(n, k, sc, seed=seed)
        mods[name] = MixQLinear(k, n, device=DEV).load_packed(*(_t(lin[x]) for x in ("W8", "scale_b", "fp_weight", "ind")))
    mlp = MixQLlamaMLP(mods["gate"], mods["down"], mods["up"])
    x = _t(oracle.synth_activations(M, sc_h, seed=9)).view(2, M // 2, H)
    y = mlp(x)
    want = mods["down"](mods["gate"](x, activation="silu") * mods["up"](x))
    torch.cuda.synchronize()
    assert y.shape == (2, M // 2, H) and torch.equal(y.view(torch.int16), want.view(torch.int16))


# ------------------------------------------------------------------ K-split of one-CTA tiles (one row-block of tokens)
@pytest.mark.parametrize("M,N,K", [(32, 4096, 4096), (32, 4096, 11008), (17, 1280, 8192), (64, 8192, 1024), (100, 3584, 3584),
                                   (128, 4096, 4096), (1, 4096, 4096), (32, 12288, 4096), (5, 264, 1040)])
def test_small_batch_ksplit(B, lib, oracle, M, N, K):
    """gemm_config = tile id + 100 x split: 2 / 4 / 8 CTAs of a cluster share one 128 x {128, 64, 32} tile along K and the partial
    sums of the rows that exist cross shared memory (st.async).  int32 sums are associative: without the outlier slab the
    output equals the oracle bit for bit, with it (and with the fused bias / SiLU epilogue) the unsplit kernel's; auto (0)
    must agree too whatever it picks."""
    rng = np.random.default_rng(M + N + K)
    q = rng.integers(-127, 128, (M, K), dtype=np.int8)
    w = rng.integers(-127, 128, (N, K), dtype=np.int8)
    sa = (rng.random(M) * 0.05 + 1e-3).astype(np.float16)
    sb = (rng.random(N) * 0.002 + 1e-4).astype(np.float16)
    fpA = (rng.standard_normal((M, 128)) * 4).astype(np.float16)
    fpW = (rng.standard_normal((N, 128)) * 0.05).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float16)
    tq, tw, tsa, tsb, tfa, tfw, tb = _t(q), _t(w), _t(sa), _t(sb), _t(fpA), _t(fpW), _t(bias)
    ref = oracle.epilogue(oracle.igemm(q, w), sa, sb, None)
    base = torch.empty(M, N, dtype=torch.float16, device=DEV)
    B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, base, config=1)
    base_b = torch.empty_like(base)
    B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, base_b, config=1, bias=tb, activation=B.ACT_SILU)
    served = 0
    for tile in (1, 3, 15):
        for split in (2, 4, 8):
            cfg = split * 100 + tile
            o = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
            try:
                B.gemm_dequant(tq, tw, tsa, tsb, None, None, o, config=cfg)
            except Exception as e:
                assert "K-split" in str(e), e
                continue
            served += 1
            torch.cuda.synchronize()
            bad = np.argwhere(o.cpu().numpy().view(np.uint16) != ref.view(np.uint16))
            assert bad.size == 0, f"cfg {cfg}: {len(bad)} mismatches vs oracle, first at {bad[:5].tolist()}"
            for _ in range(2):       # back to back: barriers are single-use per launch
                B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, o, config=cfg)
            torch.cuda.synchronize()
            assert torch.equal(o.view(torch.int16), base.view(torch.int16)), f"cfg {cfg}: mixed output differs from the unsplit kernel"
            B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, o, config=cfg, bias=tb, activation=B.ACT_SILU)
            torch.cuda.synchronize()
            assert torch.equal(o.view(torch.int16), base_b.view(torch.int16)), f"cfg {cfg}: fused epilogue differs"
    assert served >= 3, served
    o = torch.empty_like(base)
    B.gemm_dequant(tq, tw, tsa, tsb, tfa, tfw, o)      # auto
    torch.cuda.synchronize()
    assert torch.equal(o.view(torch.int16), base.view(torch.int16))
