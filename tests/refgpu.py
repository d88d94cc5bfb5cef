"""ctypes front end of oracle/_ref/libref_driver.so -- the reference's own CUDA kernels
(compiled from /root/reference/kernel/i8gemm.cu) replayed in enqueueImpl's order.
TEST / BENCH INFRASTRUCTURE ONLY; the product never loads it."""
from __future__ import annotations

import ctypes
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent.parent / "oracle" / "_ref"
_lib = None


def available() -> bool:
    return (REF_DIR / "libref_driver.so").exists() and (REF_DIR / "libref_i8gemm.so").exists()


def load():
    global _lib
    if _lib is None:
        ctypes.CDLL(str(REF_DIR / "libref_i8gemm.so"), mode=ctypes.RTLD_GLOBAL)
        L = ctypes.CDLL(str(REF_DIR / "libref_driver.so"))
        vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.ref_init.restype = ci
        L.ref_workspace_size.restype = ctypes.c_size_t
        L.ref_workspace_size.argtypes = [i64, i64, i64]
        L.ref_int8quant.argtypes = [vp, ci, ci, vp, vp, vp]
        L.ref_extract.argtypes = [vp, ci, ci, vp, ci, vp, vp]
        L.ref_fused_dequant.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, vp, vp]
        L.ref_enqueue.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, vp, vp]
        L.ref_rcp_table.argtypes = [vp, vp]
        L.ref_hdiv.argtypes = [vp, vp, vp, vp, ci, vp]
        for f in ("ref_int8quant", "ref_extract", "ref_fused_dequant", "ref_enqueue", "ref_rcp_table", "ref_hdiv"):
            getattr(L, f).restype = ci
        assert L.ref_init() == 0
        _lib = L
    return _lib


def _s():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def int8quant(A):
    import torch
    M, K = A.shape
    q = torch.empty(M, K, dtype=torch.int8, device=A.device)
    sa = torch.empty(M, dtype=torch.float16, device=A.device)
    assert load().ref_int8quant(_p(A), M, K, _p(q), _p(sa), _s()) == 0
    return q, sa


def extract(A, ind):
    import torch
    M, K = A.shape
    fpA = torch.empty(M, ind.numel(), dtype=torch.float16, device=A.device)
    assert load().ref_extract(_p(A), M, K, _p(ind), ind.numel(), _p(fpA), _s()) == 0
    return fpA


def enqueue(A, W8, sb, fp_weight, ind, out=None, ws=None):
    import torch
    M, K = A.shape
    N = W8.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float16, device=A.device)
    if ws is None:
        ws = torch.empty(load().ref_workspace_size(M, N, K), dtype=torch.uint8, device=A.device)
    rc = load().ref_enqueue(_p(A), _p(W8), _p(sb), _p(fp_weight), _p(ind), _p(out), M, N, K, _p(ws), _s())
    assert rc == 0, rc
    return out


def rcp_table():
    import torch
    t = torch.empty(65536, dtype=torch.int32, device="cuda")
    assert load().ref_rcp_table(_p(t), _s()) == 0
    torch.cuda.synchronize()
    return t.cpu().numpy().view("uint32")


def hdiv(a, b):
    import torch
    q = torch.empty_like(a)
    qi = torch.empty(a.numel(), dtype=torch.int32, device=a.device)
    assert load().ref_hdiv(_p(a), _p(b), _p(q), _p(qi), a.numel(), _s()) == 0
    return q, qi


# ---- M <= 4 branch: the reference's weight-only GEMV kernels (oracle/_ref/libref_gemv.so) and its CPU packer
_gemv = None
_pre = None


def gemv_available() -> bool:
    return (REF_DIR / "libref_gemv.so").exists()


def preprocess_available() -> bool:
    return (REF_DIR / "libref_preprocess.so").exists()


def gemv(A, qweight, scales):
    """reference weight_only_batched_gemv (Int8b, PerChannel, fp16) on torch CUDA tensors."""
    import torch
    global _gemv
    if _gemv is None:
        _gemv = ctypes.CDLL(str(REF_DIR / "libref_gemv.so"))
        _gemv.ref_w8a16_gemv.restype = ctypes.c_int
        _gemv.ref_w8a16_gemv.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    M, K = A.shape
    N = scales.numel()
    out = torch.empty(M, N, dtype=torch.float16, device=A.device)
    rc = _gemv.ref_w8a16_gemv(_p(A), _p(qweight), _p(scales), _p(out), M, N, K, _s())
    assert rc == 0, rc
    return out


def _prelib():
    global _pre
    if _pre is None:
        _pre = ctypes.CDLL(str(REF_DIR / "libref_preprocess.so"))
        _pre.ref_preprocess_int8.restype = ctypes.c_int
        _pre.ref_preprocess_int8.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
        _pre.ref_symmetric_quantize_half.restype = ctypes.c_int
        _pre.ref_symmetric_quantize_half.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_size_t] * 2
    return _pre


def preprocess_int8(q_kn, arch: int = 80):
    """reference preprocess_weights(int8 [K,N] row-major) -> processed bytes (numpy, CPU)."""
    import numpy as np
    q = np.ascontiguousarray(q_kn, dtype=np.int8)
    out = np.empty_like(q)
    assert _prelib().ref_preprocess_int8(out.ctypes.data, q.ctypes.data, q.shape[0], q.shape[1], arch) == 0
    return out


def symmetric_quantize_half(W_t):
    """reference symmetric_quantize<half, half> on W^T fp16 [K,N]: (plain int8 codes [K,N], scales fp16 [N])."""
    import numpy as np
    W = np.ascontiguousarray(W_t, dtype=np.float16)
    K, N = W.shape
    proc = np.zeros((K, N), dtype=np.int8)
    unproc = np.zeros((K, N), dtype=np.int8)
    scales = np.zeros(N, dtype=np.float16)
    _prelib().ref_symmetric_quantize_half(proc.ctypes.data, unproc.ctypes.data, scales.ctypes.data, W.ctypes.data, K, N)
    return unproc, scales


# ---- MixQ/src torch-extension pieces compiled unmodified (oracle/_ref/libref_mixsrc.so, driver oracle/ref_mixsrc_driver.cu):
# generalT5LayerNorm_extract_outliers (layernorm.cu:121-198) and GemmDequantSilu as int8FusedDequantizeSiluCUDA instantiates it
_mixsrc = None


def mixsrc_available() -> bool:
    return (REF_DIR / "libref_mixsrc.so").exists()


def _mixsrc_lib():
    global _mixsrc
    if _mixsrc is None:
        import torch  # noqa: F401  (libtorch must be in the process: layernorm.cu launches on torch's current stream)
        L = ctypes.CDLL(str(REF_DIR / "libref_mixsrc.so"))
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.ref_rmsnorm_extract_quant.restype = ci
        L.ref_rmsnorm_extract_quant.argtypes = [vp, vp, vp, ctypes.c_float, ci, ci, vp, vp, ci, vp, vp]
        L.ref_int8_fused_dequant_silu.restype = ci
        L.ref_int8_fused_dequant_silu.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]
        _mixsrc = L
    return _mixsrc


def rmsnorm_extract_quant(X, gamma, eps, ind):
    """reference generalT5LayerNorm_extract_outliers: returns (normalised rows with the outlier columns zeroed, outliers
    [M, len(ind)], INT8 codes, per-token scales)."""
    import torch
    M, K = X.shape
    out = torch.empty_like(X)
    outl = torch.zeros(M, ind.numel(), dtype=torch.float16, device=X.device)
    q = torch.zeros(M, K, dtype=torch.int8, device=X.device)
    sc = torch.empty(M, dtype=torch.float16, device=X.device)
    rc = _mixsrc_lib().ref_rmsnorm_extract_quant(_p(X), _p(gamma), _p(out), float(eps), M, K, _p(outl), _p(ind), ind.numel(), _p(q), _p(sc))
    assert rc == 0, rc
    return out, outl, q, sc


def int8_fused_dequant_silu(A8, W8, scale_row, scale_col, y):
    """reference int8FusedDequantizeSiluCUDA(A, B, scale_row [M], scale_col [N], y [M,N]) -> fp16 [M,N]."""
    import torch
    M, K = A8.shape
    N = W8.shape[0]
    D = torch.empty(M, N, dtype=torch.float16, device=A8.device)
    rc = _mixsrc_lib().ref_int8_fused_dequant_silu(_p(A8), _p(W8), _p(scale_row), _p(scale_col), _p(y), _p(D), M, N, K, _s())
    assert rc == 0, rc
    return D
