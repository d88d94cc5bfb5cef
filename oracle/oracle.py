"""CPU oracle for the W8A8O16 mixed-precision linear -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product (mixq_tensorrt_llm_b200/) never does and fails
loudly when its CUDA library is missing.

The arithmetic lives in oracle/mixq_oracle.c (each function cites the reference lines it
restates); this file is the numpy/ctypes front end plus the numpy restatement of the
checkpoint packer ``pack_linear_weights`` (reference
modelopt/torch/export/model_config_utils.py:378-466).

Parity status: the reference has no tests or golden vectors for this path.  The oracle is
pinned by (a) the reference's ``to_quantized_weight`` executed from its own source file
(tests/golden/make_cpu_golden.py), (b) outputs of the reference's CUDA kernels compiled
from /root/reference/kernel/i8gemm.cu and run on a B200 (tests/golden/make_gpu_golden.py),
both committed under tests/golden/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libmixq_oracle.so"
_GOLDEN = _HERE.parent / "tests" / "golden"
NUM_OUTLIERS = 128  # fp_features, model_config_utils.py:446; num_ind, TsinghuaMixQPlugin.cpp:518

_lib = None


def build(force: bool = False) -> Path:
    """Compile oracle/mixq_oracle.c with the system gcc (see oracle/Makefile)."""
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < (_HERE / "mixq_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "libmixq_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _P(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        i64, i32 = ctypes.c_int64, ctypes.c_int
        u16p, i8p = ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_int8)
        i32p, u32p = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint32)
        L.mixq_oracle_gather.argtypes = [u16p, i64, i64, i32p, i32, u16p]
        L.mixq_oracle_quant.argtypes = [u16p, i64, i64, u32p, i32p, i32, i32, i8p, u16p]
        L.mixq_oracle_outlier_gemm.argtypes = [u16p, u16p, i64, i64, i32, u16p]
        L.mixq_oracle_igemm.argtypes = [i8p, i8p, i64, i64, i64, i32p]
        L.mixq_oracle_epilogue.argtypes = [i32p, u16p, u16p, u16p, i64, i64, u16p]
        L.mixq_oracle_forward.argtypes = [u16p, i8p, u16p, u16p, i32p, i32, i64, i64, i64, u32p, i32,
                                          u16p, u16p, i8p, u16p, i32p, u16p]
        L.mixq_oracle_rmsnorm.argtypes = [u16p, u16p, ctypes.c_float, i64, i64, u16p]
        L.mixq_oracle_rmsnorm.restype = None
        L.mixq_oracle_gemv_w8a16.argtypes = [u16p, ctypes.POINTER(ctypes.c_uint8), u16p, i64, i64, i64, u16p]
        L.mixq_oracle_gemv_w8a16.restype = None
        L.mixq_oracle_hfma.argtypes = [ctypes.c_uint16] * 3
        L.mixq_oracle_hfma.restype = ctypes.c_uint16
        L.mixq_oracle_num_threads.restype = i32
        L.mixq_oracle_set_threads.argtypes = [i32]
        L.mixq_oracle_set_threads.restype = None
        for f in ("gather", "quant", "outlier_gemm", "igemm", "epilogue", "forward"):
            getattr(L, "mixq_oracle_" + f).restype = None
        _lib = L
    return _lib


_rcp_table = None
_rcp_loaded = False


def rcp_table():
    """rcp.approx.ftz.f32 for every fp16 input, captured on a B200 (None until captured)."""
    global _rcp_table, _rcp_loaded
    if not _rcp_loaded:
        p = _GOLDEN / "rcp_approx_f16.bin"
        if p.exists():
            t = np.fromfile(p, dtype=np.uint32)
            assert t.shape == (65536,)
            _rcp_table = np.ascontiguousarray(t)
        _rcp_loaded = True
    return _rcp_table


def _rcp_ptr(use_table: bool):
    t = rcp_table() if use_table else None
    return _P(t, ctypes.c_uint32) if t is not None else None


def _u16(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.float16, a.dtype
    return a.view(np.uint16)


def num_threads() -> int:
    return int(lib().mixq_oracle_num_threads())


def set_threads(n: int | None = None) -> int:
    """Use `n` OpenMP threads (default: every host core; torchrun pins OMP_NUM_THREADS=1)."""
    lib().mixq_oracle_set_threads(int(n or os.cpu_count() or 1))
    return num_threads()


# ----------------------------------------------------------------------------- steps
def gather(A: np.ndarray, ind: np.ndarray) -> np.ndarray:
    """kernel/i8gemm.cu:198-224 -- fp_A[i, j] = A[i, ind[j]]."""
    M, K = A.shape
    ind = np.ascontiguousarray(ind, dtype=np.int32)
    out = np.empty((M, ind.size), dtype=np.float16)
    lib().mixq_oracle_gather(_P(_u16(A), ctypes.c_uint16), M, K, _P(ind, ctypes.c_int32), ind.size,
                             _P(out.view(np.uint16), ctypes.c_uint16))
    return out


def quant(A: np.ndarray, ind: np.ndarray | None = None, mask: bool = False, use_table: bool = True):
    """kernel/i8gemm.cu:66-107 -- per-token scale + int8 codes. Returns (q int8 [M,K], sa fp16 [M])."""
    M, K = A.shape
    q = np.empty((M, K), dtype=np.int8)
    sa = np.empty((M,), dtype=np.float16)
    if ind is None:
        ind = np.zeros((0,), dtype=np.int32)
    ind = np.ascontiguousarray(ind, dtype=np.int32)
    lib().mixq_oracle_quant(_P(_u16(A), ctypes.c_uint16), M, K, _rcp_ptr(use_table),
                            _P(ind, ctypes.c_int32), ind.size, int(mask), _P(q, ctypes.c_int8),
                            _P(sa.view(np.uint16), ctypes.c_uint16))
    return q, sa


def rmsnorm(X: np.ndarray, gamma: np.ndarray, eps: float) -> np.ndarray:
    """layernorm.cu:121-157 -- T5-style RMSNorm, fp32 math, fp16 result (the producer of the linear's input)."""
    M, K = X.shape
    Y = np.empty((M, K), dtype=np.float16)
    lib().mixq_oracle_rmsnorm(_P(_u16(X), ctypes.c_uint16), _P(_u16(gamma), ctypes.c_uint16), float(eps), M, K,
                              _P(Y.view(np.uint16), ctypes.c_uint16))
    return Y


def outlier_gemm(fp_A: np.ndarray, fp_weight: np.ndarray) -> np.ndarray:
    """TsinghuaMixQPlugin.cpp:122-161 -- fp16 x fp16 -> fp32 accumulate -> fp16."""
    M, kf = fp_A.shape
    N = fp_weight.shape[0]
    assert fp_weight.shape[1] == kf and kf <= 1024
    out = np.empty((M, N), dtype=np.float16)
    lib().mixq_oracle_outlier_gemm(_P(_u16(fp_A), ctypes.c_uint16), _P(_u16(fp_weight), ctypes.c_uint16),
                                   M, N, kf, _P(out.view(np.uint16), ctypes.c_uint16))
    return out


def igemm(q: np.ndarray, W8: np.ndarray) -> np.ndarray:
    """acc[m,n] = sum_k q[m,k]*W8[n,k] (exact int32)."""
    M, K = q.shape
    N = W8.shape[0]
    q = np.ascontiguousarray(q, dtype=np.int8)
    W8 = np.ascontiguousarray(W8, dtype=np.int8)
    acc = np.empty((M, N), dtype=np.int32)
    lib().mixq_oracle_igemm(_P(q, ctypes.c_int8), _P(W8, ctypes.c_int8), M, N, K, _P(acc, ctypes.c_int32))
    return acc


def epilogue(acc: np.ndarray, sa: np.ndarray, sb: np.ndarray, out0: np.ndarray | None) -> np.ndarray:
    """linear_combination_dequant.h:152-157 -- fp16(fma(float(acc), sb*sa, float(out0)))."""
    M, N = acc.shape
    acc = np.ascontiguousarray(acc, dtype=np.int32)
    out = np.empty((M, N), dtype=np.float16)
    o0 = _P(_u16(out0), ctypes.c_uint16) if out0 is not None else None
    lib().mixq_oracle_epilogue(_P(acc, ctypes.c_int32), _P(_u16(sa), ctypes.c_uint16),
                               _P(_u16(sb), ctypes.c_uint16), o0, M, N,
                               _P(out.view(np.uint16), ctypes.c_uint16))
    return out


def epilogue_ex(acc: np.ndarray, sa: np.ndarray, sb: np.ndarray, out0: np.ndarray | None, bias: np.ndarray | None = None,
                silu: bool = False) -> np.ndarray:
    """The fused epilogue of SURVEY.md 8f #4:
        y   = fp16( act( fma(float(acc), sb*sa, float(out0)) ) )   act = silu(x) = x / (1 + exp(-x)) in fp32 before the rounding
                                                                    (linear_combination_dequant.h:167-272)
        out = fp16( float(y) + float(bias[n]) )                     the bias add the reference does outside the kernel
                                                                    (plugin.py:158-160)
    The FMA is evaluated in float64 and rounded to fp32 (exact except for astronomically rare double-rounding ties);
    silu uses the exact exp (the reference extension is built with --use_fast_math: ~2 fp32 ulp away)."""
    p = (sb.astype(np.float32)[None, :] * sa.astype(np.float32)[:, None]).astype(np.float64)
    r = np.ascontiguousarray(acc, dtype=np.int32).astype(np.float32).astype(np.float64) * p
    if out0 is not None:
        r = r + out0.astype(np.float64)
    r = r.astype(np.float32)
    if silu:
        with np.errstate(over="ignore"):
            r = (r.astype(np.float64) / (1.0 + np.exp(-r.astype(np.float64)))).astype(np.float32)
    y = r.astype(np.float16)
    if bias is not None:
        y = (y.astype(np.float32) + bias.astype(np.float32)[None, :]).astype(np.float16)
    return y


def forward(A, W8, sb, fp_weight, ind, mask: bool = False, use_table: bool = True, return_parts: bool = False):
    """MixQPlugin::enqueueImpl, M>4 branch (TsinghuaMixQPlugin.cpp:518-532), whole path."""
    M, K = A.shape
    N = W8.shape[0]
    ind = np.ascontiguousarray(ind, dtype=np.int32)
    W8 = np.ascontiguousarray(W8, dtype=np.int8)
    fp_A = np.empty((M, ind.size), dtype=np.float16)
    out0 = np.empty((M, N), dtype=np.float16)
    q = np.empty((M, K), dtype=np.int8)
    sa = np.empty((M,), dtype=np.float16)
    acc = np.empty((M, N), dtype=np.int32)
    out = np.empty((M, N), dtype=np.float16)
    u16 = ctypes.c_uint16
    lib().mixq_oracle_forward(_P(_u16(A), u16), _P(W8, ctypes.c_int8), _P(_u16(sb), u16),
                              _P(_u16(fp_weight), u16), _P(ind, ctypes.c_int32), ind.size, M, N, K,
                              _rcp_ptr(use_table), int(mask), _P(fp_A.view(np.uint16), u16),
                              _P(out0.view(np.uint16), u16), _P(q, ctypes.c_int8),
                              _P(sa.view(np.uint16), u16), _P(acc, ctypes.c_int32),
                              _P(out.view(np.uint16), u16))
    if return_parts:
        return dict(out=out, fp_A=fp_A, out0=out0, q=q, sa=sa, acc=acc)
    return out


def gated_mlp_half(A, gate: dict, up: dict, mask: bool = False, use_table: bool = True) -> np.ndarray:
    """Input half of the reference's fused Llama MLP, MixLlamaMLP.forward (MixQ/src/mixquant/modules/fused/mlp.py:57-70):
        up_output   = up_proj(x, cache)                                        fp16 [M, N]
        gate_output = gate_proj.forward_without_preconditionFusedSilu(x, cache) fp16, SiLU fused in the dequant epilogue
                      (modules/linear.py:288-373 -> int8FusedDequantizeSilu, linear_combination_dequant.h:167-272)
        gate_output *= up_output                                                fp16 multiply (one rounding)
    One quantised A serves both projections (the MixGemmCache); `gate` / `up` are packed linears (W8, scale_b, fp_weight, ind)
    with identical `ind`."""
    assert np.array_equal(gate["ind"], up["ind"]), "gate and up share the activation's outlier columns"
    g = forward(A, gate["W8"], gate["scale_b"], gate["fp_weight"], gate["ind"], mask=mask, use_table=use_table, return_parts=True)
    u = forward(A, up["W8"], up["scale_b"], up["fp_weight"], up["ind"], mask=mask, use_table=use_table, return_parts=True)
    gs = epilogue_ex(g["acc"], g["sa"], gate["scale_b"], g["out0"], silu=True)
    return (gs.astype(np.float32) * u["out"].astype(np.float32)).astype(np.float16)


# ---- MixQ/src torch path (BASELINE.json configs[0]: "via MixQ/src torch reference"), 8-bit branch ---------------------
def mixsrc_init(W: np.ndarray) -> dict:
    """MixLinear_GEMM.from_linear, bit == 8 (MixQ/src/mixquant/modules/linear.py:110-118): per-output-channel scale
    max|W| / 127 in fp16, codes round(W / scale) -- the division in fp16 like the torch call, NO outlier columns at init
    (`ind` empty, :42) and none removed from the codes."""
    W = np.ascontiguousarray(W, dtype=np.float16)
    scale_col = (np.abs(W).astype(np.float32).max(axis=1) / np.float32(127)).astype(np.float16)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.rint((W / scale_col[:, None]).astype(np.float16).astype(np.float32))
    q = np.nan_to_num(q, nan=0.0).astype(np.int8)
    return dict(q_weight=q, scale_col=scale_col, ind=np.zeros((0,), dtype=np.int32), weight_cache=np.zeros((W.shape[0], 0), dtype=np.float16),
                add_outliers=True, cnt=0)


def mixsrc_forward(st: dict, x: np.ndarray, sigma: float = 6.0, stop: int = 2, use_table: bool = True) -> np.ndarray:
    """MixLinear_GEMM.forward(x, cache, unfused=True), 8-bit, non-Hopper branch (linear.py:163-286), on a [M, K] fp16 batch.
    `st` is the layer state of mixsrc_init and is UPDATED like the module (ind / weight_cache grow while add_outliers):
      1. known outlier columns are extracted and ZEROED in the activations (ExtractOutliersAndSetToZeros, cult.cu:1588),
      2. per-token scale max|x| / 127 and INT8 codes (FindRowScale),
      3. while add_outliers: if any token scale exceeds sigma / 127, the columns holding a value above sigma become outliers
         (FindOutliers, :154-159: torch.unique -> ascending), are extracted and zeroed too, their weight columns are
         DEQUANTISED from the codes (q_weight[:, ind] * scale_col, fp16, :203-204) and the batch is re-quantised (:219);
         growth stops after `stop` calls or beyond 256 columns (:222-224),
      4. outlier product in fp16 with fp32 accumulation (torch.mm) and the fused dequant epilogue (int8FusedDequantize:
         the plugin's epilogue, linear_combination_dequant.h:152-157)."""
    x = np.array(x, dtype=np.float16, copy=True)
    M, K = x.shape
    acts = np.zeros((M, 0), dtype=np.float16)
    if st["ind"].size:
        acts = x[:, st["ind"]].copy()
        x[:, st["ind"]] = 0
    q, sa = quant(x, use_table=use_table)
    if st["add_outliers"]:
        if np.float16(sa.astype(np.float32).max()) > np.float16(np.float16(sigma) / np.float16(127)):
            new = np.unique(np.where(np.abs(x) > np.float16(sigma))[1]).astype(np.int32)
            a_new = x[:, new].copy()
            x[:, new] = 0
            w_new = (st["q_weight"][:, new].astype(np.float16) * st["scale_col"][:, None]).astype(np.float16)
            acts = np.hstack([acts, a_new])
            st["weight_cache"] = np.hstack([st["weight_cache"], w_new])
            st["ind"] = np.concatenate([st["ind"], new]).astype(np.int32)
            q, sa = quant(x, use_table=use_table)
        st["cnt"] += 1
        if st["cnt"] >= stop or st["ind"].size > 256:
            st["add_outliers"] = False
    acc = igemm(q, st["q_weight"])
    out0 = outlier_gemm(acts, st["weight_cache"]) if st["ind"].size else None
    return epilogue(acc, sa, st["scale_col"], out0)


def forward_f64(A, W8, sb, fp_weight, ind, q, sa):
    """Higher-precision 'truth' for error reporting: same quantised operands, float64 math, no
    intermediate fp16 rounding of the outlier product."""
    acc = igemm(q, W8).astype(np.float64)
    o = A[:, ind].astype(np.float64) @ fp_weight.astype(np.float64).T
    return acc * (sb.astype(np.float64)[None, :] * sa.astype(np.float64)[:, None]) + o


# ----------------------------------------------------------------------------- packer
def pack_linear_weights(W: np.ndarray, act_scale: np.ndarray, fp_features: int = NUM_OUTLIERS):
    """numpy restatement of pack_linear_weights / to_quantized_weight
    (modelopt/torch/export/model_config_utils.py:429-464, 298-308) for one linear.

      sb        = fp16(max_k |W[n,k]| / 127)                computed BEFORE zeroing (:429-430)
      ind       = argsort(act_scale)[-128:]  (ascending)    (:448)
      fp_weight = W[:, ind]                                  (:452)
      W[:, ind] = 0                                          (:453)
      W8        = round_half_even(fp16(W / sb)).clamp(-128,127).int8   (:308, CPU fp16 divide)

    Returns dict(W8 int8 [N,K], scale_b fp16 [N], fp_weight fp16 [N,128], ind int32 [128]).
    The weight-only `qweight/scales` pair (EETQ layout, :437-441) feeds only the M<=4 branch
    and is not produced here.
    """
    W = np.array(W, dtype=np.float16, copy=True)
    N, K = W.shape
    # torch: (max(abs(W), dim=1) / 127).to(float16) -- fp16 tensor / python int runs in fp32, rounds to fp16
    sb = (np.abs(W).max(axis=1).astype(np.float32) / np.float32(127)).astype(np.float16)
    # torch.sort(layer_scales)[1][-fp_features:]
    ind = np.argsort(np.asarray(act_scale, dtype=np.float32), kind="stable")[-fp_features:].astype(np.int32)
    fp_weight = np.ascontiguousarray(W[:, ind])
    W[:, ind] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        quot = (W.astype(np.float32) / sb.astype(np.float32)[:, None]).astype(np.float16)
    W8 = np.clip(np.rint(quot.astype(np.float32)), -128, 127)
    W8 = np.nan_to_num(W8, nan=0.0).astype(np.int8)
    return dict(W8=W8, scale_b=sb, fp_weight=fp_weight, ind=ind)


# ----------------------------------------------------------------------------- M <= 4 branch (weight-only)
_PERM16 = np.array([0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15])   # cutlass_preprocessors.cc:133-134,174-175


def eetq_preprocess(q_kn: np.ndarray) -> np.ndarray:
    """numpy restatement of preprocess_weights_for_mixed_gemm for int8 on the Sm80 layout
    (weightonlykernel/cutlass_kernels/cutlass_preprocessors.cc:497-533; arch 80..90 -> :121-122,
    ColumnMajorTileInterleave<64, 2> + OpMultiplyAddDequantizeInterleavedBToA):
      1. permute_B_rows_for_mixed_gemm   (:137-199)  rows (K) permuted inside groups of 16
      2. subbyte_transpose               (:322-335)  [K, N] row-major -> [N, K]
      3. interleave_column_major_tensor  (:432-495)  two output channels interleaved in runs of 64 codes
      4. add_bias_and_interleave_int8s   (:337-358)  +128, then bytes 1 and 2 of every four swapped
    `q_kn` is the row-major int8 [K, N] matrix (the TRANSPOSED weight, as model_config_utils.py:437 passes it);
    returns the processed bytes with the same nominal shape [K, N] (int8 view, as EETQ returns it)."""
    q = np.ascontiguousarray(q_kn, dtype=np.int8)
    K, N = q.shape
    assert K % 64 == 0 and N % 2 == 0, "EETQ layout: K multiple of 64 (rows_per_column_tile), N even"
    q = q.reshape(K // 16, 16, N)[:, _PERM16, :].reshape(K, N)
    t = np.ascontiguousarray(q.T)                                                    # [N, K]
    t = t.reshape(N // 2, 2, K // 64, 64).transpose(0, 2, 1, 3).reshape(N // 2, 2 * K)
    u = (t.astype(np.int16) + 128).astype(np.uint8).reshape(-1, 4)[:, [0, 2, 1, 3]]
    return np.ascontiguousarray(u).reshape(K, N).view(np.int8)


def eetq_unprocess(processed: np.ndarray) -> np.ndarray:
    """Inverse of eetq_preprocess: the plain int8 [K, N] codes."""
    K, N = processed.shape
    u = processed.view(np.uint8).reshape(-1, 4)[:, [0, 2, 1, 3]]
    t = (u.astype(np.int16) - 128).astype(np.int8).reshape(N // 2, K // 64, 2, 64).transpose(0, 2, 1, 3).reshape(N, K)
    q = np.ascontiguousarray(t.T)
    inv = np.argsort(_PERM16)
    return q.reshape(K // 16, 16, N)[:, inv, :].reshape(K, N)


def eetq_quant_weights(W_t: np.ndarray):
    """numpy restatement of EETQ.quant_weights(W.T, torch.int8, False) = symmetric_quantize<half, half>
    (EETQ/csrc/cutlass_kernels/fpA_intB_gemm_wrapper.cu:28-110 -> cutlass_preprocessors.cc:581-678), as called at
    model_config_utils.py:437-438 on the UN-zeroed weight:
      per output channel n:  s = max_k |W_t[k, n]| * (1/128)  in fp32; scales[n] = fp16(s)          (:626-629)
      code = clamp(round_half_away(W_t[k, n] / s), -128, 127)   with the fp32 s, not its fp16 image   (:637-641)
    Returns (processed int8 [K, N], scales fp16 [N])."""
    W_t = np.ascontiguousarray(W_t, dtype=np.float16)
    s = np.abs(W_t.astype(np.float32)).max(axis=0) * np.float32(1.0 / 128.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        x = W_t.astype(np.float32) / s[None, :]
    r = np.where(x >= 0, np.floor(x + np.float32(0.5)), np.ceil(x - np.float32(0.5)))   # C round(): half away from zero
    # int8_t(std::max(-128.f, std::min(127.f, r))): std::min(127.f, NaN) keeps 127.f, so an all-zero channel (0/0) codes 127
    q = np.nan_to_num(np.clip(r, -128, 127), nan=127.0).astype(np.int8)
    return eetq_preprocess(q), s.astype(np.float16)


def gemv_w8a16(A: np.ndarray, qweight: np.ndarray, scales: np.ndarray) -> np.ndarray:
    """The M <= 4 branch: Out = A . dequant(qweight) with the reference kernel's exact rounding sequence
    (oracle/mixq_oracle.c: mixq_oracle_gemv_w8a16).  A fp16 [M, K], qweight processed int8 [K, N], scales fp16 [N]."""
    A = np.ascontiguousarray(A, dtype=np.float16)
    M, K = A.shape
    N = scales.shape[0]
    assert qweight.size == K * N and 1 <= M <= 4 and N % 4 == 0 and K % 64 == 0
    qw = np.ascontiguousarray(qweight).view(np.uint8).reshape(-1)
    out = np.empty((M, N), dtype=np.float16)
    lib().mixq_oracle_gemv_w8a16(_P(_u16(A), ctypes.c_uint16), _P(qw, ctypes.c_uint8), _P(_u16(scales), ctypes.c_uint16),
                                 M, N, K, _P(out.view(np.uint16), ctypes.c_uint16))
    return out


def as_plugin_tensors(packed: dict):
    """The half-typed containers the TensorRT side sees (plugin.py:99-111,
    mixlib.int8_matrix_to_half / int_to_half = raw byte reinterpretation)."""
    return dict(weight=packed["W8"].view(np.float16), fp_ind=packed["ind"].view(np.float16),
                fp_weight=packed["fp_weight"], weights_scaling_factor=packed["scale_b"])


# ----------------------------------------------------------------------------- synthetic data
def load_act_scales(name: str) -> np.ndarray | None:
    """Layer-0 activation maxima committed from /root/reference/act_scales (tests/golden/act_scales_l0.npz)."""
    p = _GOLDEN / "act_scales_l0.npz"
    if not p.exists():
        return None
    z = np.load(p)
    return z[name] if name in z.files else None


def synth_act_scale(K: int, seed: int = 1234) -> np.ndarray:
    """Fallback per-channel activation scale: ~N(0,1) channels with 128 random channels x20."""
    rng = np.random.default_rng(seed)
    s = np.abs(rng.standard_normal(K)).astype(np.float32) * 0.5 + 0.05
    hot = rng.choice(K, size=min(NUM_OUTLIERS, K), replace=False)
    s[hot] *= 20.0
    return s


def synth_linear(N: int, K: int, act_scale: np.ndarray | None = None, seed: int = 1234):
    """Synthetic packed linear: W ~ N(0, 0.02^2) fp16, packed per pack_linear_weights."""
    rng = np.random.default_rng(seed)
    W = (rng.standard_normal((N, K), dtype=np.float32) * 0.02).astype(np.float16)
    if act_scale is None:
        act_scale = synth_act_scale(K, seed)
    p = pack_linear_weights(W, act_scale)
    p["W"] = W
    p["act_scale"] = np.asarray(act_scale, dtype=np.float32)
    return p


def synth_activations(M: int, act_scale: np.ndarray, seed: int = 4321) -> np.ndarray:
    """A = randn(M,K) * act_scale/3 in fp16."""
    rng = np.random.default_rng(seed)
    K = act_scale.shape[0]
    return (rng.standard_normal((M, K), dtype=np.float32) * (act_scale[None, :] / 3.0)).astype(np.float16)
