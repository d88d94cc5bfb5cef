"""ctypes binding of libmixq_b200.so (the C ABI in include/mixq_b200.h).

There is no fallback of any kind: if the shared library is missing this module raises, and
every compute entry point raises ``MixQError`` when the library reports a failure (for example
no sm_100 device).  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libmixq_b200.so"

NUM_OUTLIERS = 128
FLAG_MASK_OUTLIERS = 1
FLAG_FORCE_MIXED = 2
FLAG_HOST_ASYNC = 1 << 8

# every symbol include/mixq_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "mixq_version", "mixq_last_error", "mixq_device_ok", "mixq_workspace_size", "mixq_workspace_size_opt", "mixq_enqueue", "mixq_enqueue_ex",
    "mixq_gemm_dequant_ex", "mixq_gated_workspace_size", "mixq_enqueue_gated", "mixq_gemm_dequant_gated",
    "mixq_quant_extract", "mixq_rmsnorm_quant_extract", "mixq_gemv_w8a16", "mixq_gemm_dequant", "mixq_gemm_dequant_ws", "mixq_gemm_workspace_size",
    "mixq_host_scratch_size", "mixq_linear_host", "mixq_linears_host_scratch_size", "mixq_linears_host", "mixq_gated_host_scratch_size", "mixq_gated_host", "mixq_host_drain",
    "mixq_allreduce_staging_size", "mixq_allreduce_counter_size", "mixq_allreduce_check", "mixq_enqueue_allreduce", "mixq_gemm_dequant_allreduce",
    "mixq_enqueue_opt", "mixq_gemm_dequant_opt", "mixq_enqueue_allreduce_opt", "mixq_gemm_dequant_allreduce_opt",
    "mixq_decode_workspace_size",
    "mixq_launch_count", "mixq_debug_set_trace", "mixq_debug_fat_plan", "mixq_debug_host_part_offset", "initOpenAiTritonPlugins", "mixq_plugin_create",
    "mixq_plugin_deserialize", "mixq_plugin_clone", "mixq_plugin_destroy", "mixq_plugin_type",
    "mixq_plugin_version", "mixq_plugin_namespace", "mixq_plugin_nb_outputs",
    "mixq_plugin_serialization_size", "mixq_plugin_serialize", "mixq_plugin_supports_format",
    "mixq_plugin_workspace_size", "mixq_plugin_enqueue",
]


class MixQError(RuntimeError):
    pass


class Tensors(ctypes.Structure):
    """struct mixq_tensors"""
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("A", "W8", "scale_b", "fp_weight", "ind", "q_weight", "scaling_factors", "Out")]


MAX_RANKS = 8
ACT_NONE, ACT_SILU = 0, 1


class Epilogue(ctypes.Structure):
    """struct mixq_epilogue"""
    _fields_ = [("bias", ctypes.c_void_p), ("activation", ctypes.c_int)]



class Options(ctypes.Structure):
    """struct mixq_options: per-call tuning (tile configuration id, SM limit); never changes a result bit"""
    _fields_ = [("gemm_config", ctypes.c_int), ("sm_limit", ctypes.c_int)]


class PeerGroup(ctypes.Structure):
    """struct mixq_peer_group"""
    _fields_ = [("world", ctypes.c_int), ("rank", ctypes.c_int),
                ("out", ctypes.c_void_p * MAX_RANKS), ("staging", ctypes.c_void_p * MAX_RANKS),
                ("counters", ctypes.c_void_p * MAX_RANKS),
                ("staging_bytes", ctypes.c_size_t), ("counter_bytes", ctypes.c_size_t),
                ("out_multicast", ctypes.c_void_p)]


_lib = None


def load() -> ctypes.CDLL:
    """Load libmixq_b200.so; raises if it has not been built (python -m mixq_tensorrt_llm_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MixQError(f"{LIB_PATH} not found: build it with `python -m mixq_tensorrt_llm_b200.build` "
                        "(there is no CPU or PyTorch fallback)")
    L = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
    vp, i64, sz, u32, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int
    L.mixq_version.restype = ctypes.c_char_p
    L.mixq_last_error.restype = ctypes.c_char_p
    L.mixq_device_ok.restype = ci
    L.mixq_workspace_size.restype = sz
    L.mixq_workspace_size.argtypes = [i64, i64, i64]
    L.mixq_workspace_size_opt.restype = sz
    L.mixq_workspace_size_opt.argtypes = [i64, i64, i64, ctypes.POINTER(Options)]
    L.mixq_enqueue.restype = ci
    L.mixq_enqueue.argtypes = [ctypes.POINTER(Tensors), i64, i64, i64, vp, sz, u32, vp]
    L.mixq_enqueue_ex.restype = ci
    L.mixq_enqueue_ex.argtypes = [ctypes.POINTER(Tensors), i64, i64, i64, vp, sz, ctypes.POINTER(Epilogue), u32, vp]
    L.mixq_gemm_dequant_ex.restype = ci
    L.mixq_gemm_dequant_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, ctypes.POINTER(Epilogue), vp]
    L.mixq_gated_workspace_size.restype = sz
    L.mixq_gated_workspace_size.argtypes = [i64, i64, i64]
    L.mixq_enqueue_gated.restype = ci
    L.mixq_enqueue_gated.argtypes = [ctypes.POINTER(Tensors), ctypes.POINTER(Tensors), i64, i64, i64, vp, sz, ctypes.POINTER(Options), u32, vp]
    L.mixq_gemm_dequant_gated.restype = ci
    L.mixq_gemm_dequant_gated.argtypes = [vp] * 10 + [i64, i64, i64, ctypes.POINTER(Options), vp, sz, vp]
    L.mixq_quant_extract.restype = ci
    L.mixq_quant_extract.argtypes = [vp, i64, i64, vp, ci, vp, vp, vp, u32, vp]
    L.mixq_rmsnorm_quant_extract.restype = ci
    L.mixq_rmsnorm_quant_extract.argtypes = [vp, vp, ctypes.c_float, i64, i64, vp, ci, vp, vp, vp, vp, u32, vp]
    L.mixq_gemm_dequant.restype = ci
    L.mixq_gemm_dequant.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, vp]
    L.mixq_gemv_w8a16.restype = ci
    L.mixq_gemv_w8a16.argtypes = [vp, vp, vp, vp, i64, i64, i64, vp]
    L.mixq_gemm_workspace_size.restype = sz
    L.mixq_gemm_workspace_size.argtypes = []
    L.mixq_gemm_dequant_ws.restype = ci
    L.mixq_gemm_dequant_ws.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, vp, sz, vp]
    L.mixq_host_scratch_size.restype = sz
    L.mixq_host_scratch_size.argtypes = [i64, i64, i64]
    L.mixq_linear_host.restype = ci
    L.mixq_linear_host.argtypes = [ctypes.POINTER(Tensors), vp, vp, i64, i64, i64, vp, sz, u32, vp]
    L.mixq_linears_host_scratch_size.restype = sz
    L.mixq_linears_host_scratch_size.argtypes = [i64, ctypes.POINTER(i64), ci, i64]
    L.mixq_linears_host.restype = ci
    L.mixq_linears_host.argtypes = [ctypes.POINTER(ctypes.POINTER(Tensors)), ci, vp, ctypes.POINTER(vp), i64, ctypes.POINTER(i64), i64,
                                    vp, sz, u32, vp]
    L.mixq_host_drain.restype = ci
    L.mixq_debug_host_part_offset.restype = i64
    L.mixq_debug_host_part_offset.argtypes = [sz, sz, u32]
    L.mixq_debug_fat_plan.restype = ci
    L.mixq_debug_fat_plan.argtypes = [i64, i64, ci, ci, ci, ctypes.POINTER(ci)]
    L.mixq_host_drain.argtypes = [vp]
    L.mixq_gated_host_scratch_size.restype = sz
    L.mixq_gated_host_scratch_size.argtypes = [i64, i64, i64]
    L.mixq_gated_host.restype = ci
    L.mixq_gated_host.argtypes = [ctypes.POINTER(Tensors), ctypes.POINTER(Tensors), vp, vp, i64, i64, i64, vp, sz, u32, vp]
    L.mixq_allreduce_check.restype = ci
    L.mixq_allreduce_check.argtypes = [vp, ci, vp]
    L.mixq_allreduce_staging_size.restype = sz
    L.mixq_allreduce_staging_size.argtypes = [i64, i64, ci]
    L.mixq_allreduce_counter_size.restype = sz
    L.mixq_allreduce_counter_size.argtypes = [i64, i64, ci]
    L.mixq_enqueue_allreduce.restype = ci
    L.mixq_enqueue_allreduce.argtypes = [ctypes.POINTER(Tensors), i64, i64, i64, vp, sz, ctypes.POINTER(PeerGroup), u32, vp]
    L.mixq_gemm_dequant_allreduce.restype = ci
    L.mixq_gemm_dequant_allreduce.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, ctypes.POINTER(PeerGroup), vp]
    L.mixq_launch_count.restype = ctypes.c_uint64
    L.mixq_debug_set_trace.restype = ci
    L.mixq_debug_set_trace.argtypes = [vp]
    PE, PO, PT, PG = ctypes.POINTER(Epilogue), ctypes.POINTER(Options), ctypes.POINTER(Tensors), ctypes.POINTER(PeerGroup)
    L.mixq_enqueue_opt.restype = ci
    L.mixq_enqueue_opt.argtypes = [PT, i64, i64, i64, vp, sz, PE, PO, u32, vp]
    L.mixq_gemm_dequant_opt.restype = ci
    L.mixq_gemm_dequant_opt.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, PE, PO, vp, sz, vp]
    L.mixq_enqueue_allreduce_opt.restype = ci
    L.mixq_enqueue_allreduce_opt.argtypes = [PT, i64, i64, i64, vp, sz, PG, PO, u32, vp]
    L.mixq_gemm_dequant_allreduce_opt.restype = ci
    L.mixq_gemm_dequant_allreduce_opt.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, PG, PO, vp]
    L.mixq_decode_workspace_size.restype = sz
    L.mixq_decode_workspace_size.argtypes = [i64, i64]
    L.initOpenAiTritonPlugins.restype = ctypes.c_bool
    L.initOpenAiTritonPlugins.argtypes = [vp, ctypes.c_char_p]
    L.mixq_plugin_create.restype = vp
    L.mixq_plugin_create.argtypes = [ctypes.c_char_p, ci, ci, ci]
    L.mixq_plugin_deserialize.restype = vp
    L.mixq_plugin_deserialize.argtypes = [ctypes.c_char_p, vp, sz]
    L.mixq_plugin_clone.restype = vp
    L.mixq_plugin_clone.argtypes = [vp]
    L.mixq_plugin_destroy.restype = None
    L.mixq_plugin_destroy.argtypes = [vp]
    for f in ("type", "version", "namespace"):
        getattr(L, "mixq_plugin_" + f).restype = ctypes.c_char_p
        getattr(L, "mixq_plugin_" + f).argtypes = [vp]
    L.mixq_plugin_nb_outputs.restype = ci
    L.mixq_plugin_nb_outputs.argtypes = [vp]
    L.mixq_plugin_serialization_size.restype = sz
    L.mixq_plugin_serialization_size.argtypes = [vp]
    L.mixq_plugin_serialize.restype = None
    L.mixq_plugin_serialize.argtypes = [vp, vp]
    L.mixq_plugin_supports_format.restype = ci
    L.mixq_plugin_supports_format.argtypes = [vp, ci, ci, ci]
    L.mixq_plugin_workspace_size.restype = sz
    L.mixq_plugin_workspace_size.argtypes = [vp, ctypes.POINTER(i64), ci, i64]
    L.mixq_plugin_enqueue.restype = ci
    L.mixq_plugin_enqueue.argtypes = [vp, ctypes.POINTER(i64), ci, i64, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise MixQError(f"{what} failed with status {rc}: {load().mixq_last_error().decode()}")


def require_device() -> None:
    if not load().mixq_device_ok():
        raise MixQError("libmixq_b200 needs a CUDA device of compute capability 10.x (B200); none is usable "
                        "and there is no fallback path")


# ------------------------------------------------------------------ torch-facing helpers
def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def workspace_size(M: int, N: int, K: int, config: int = 0) -> int:
    if config:
        return int(load().mixq_workspace_size_opt(M, N, K, ctypes.byref(Options(int(config), 0))))
    return int(load().mixq_workspace_size(M, N, K))


def make_tensors(A, W8, scale_b, fp_weight, ind, Out, q_weight=None, scaling_factors=None) -> Tensors:
    t = Tensors()
    t.A, t.W8, t.scale_b, t.fp_weight, t.ind = (x.data_ptr() if x is not None else None
                                               for x in (A, W8, scale_b, fp_weight, ind))
    t.q_weight = q_weight.data_ptr() if q_weight is not None else None
    t.scaling_factors = scaling_factors.data_ptr() if scaling_factors is not None else None
    t.Out = Out.data_ptr() if Out is not None else None
    return t


def _opts(config: int, sm_limit: int):
    return ctypes.byref(Options(int(config), int(sm_limit))) if (config or sm_limit) else None


def enqueue(A, W8, scale_b, fp_weight, ind, Out, workspace, flags: int = 0, stream=None, q_weight=None,
            scaling_factors=None, bias=None, activation: int = 0, config: int = 0, sm_limit: int = 0) -> None:
    """mixq_enqueue on torch CUDA tensors (A [M,K] fp16 contiguous, Out [M,N] fp16).  With q_weight / scaling_factors
    (the EETQ pair) a call with M <= 4 takes the weight-only branch, as the reference plugin does."""
    M, K = A.shape
    N = Out.shape[-1]
    t = make_tensors(A, W8, scale_b, fp_weight, ind, Out, q_weight, scaling_factors)
    if config or sm_limit:
        e = Epilogue(bias.data_ptr() if bias is not None else None, int(activation))
        check(load().mixq_enqueue_opt(ctypes.byref(t), M, N, K, _ptr(workspace), workspace.numel() * workspace.element_size(),
                                      ctypes.byref(e), _opts(config, sm_limit), flags, _stream(stream)), "mixq_enqueue_opt")
        return
    if bias is not None or activation:
        e = Epilogue(bias.data_ptr() if bias is not None else None, int(activation))
        check(load().mixq_enqueue_ex(ctypes.byref(t), M, N, K, _ptr(workspace), workspace.numel() * workspace.element_size(),
                                     ctypes.byref(e), flags, _stream(stream)), "mixq_enqueue_ex")
        return
    check(load().mixq_enqueue(ctypes.byref(t), M, N, K, _ptr(workspace), workspace.numel() * workspace.element_size(),
                              flags, _stream(stream)), "mixq_enqueue")


def gated_workspace_size(M: int, N: int, K: int) -> int:
    return int(load().mixq_gated_workspace_size(M, N, K))


def enqueue_gated(A, gate, up, ind, Out, workspace, flags: int = 0, stream=None, config: int = 0, sm_limit: int = 0) -> None:
    """mixq_enqueue_gated: Out [M,N] = fp16(silu(gate(A))) * fp16(up(A)); ``gate`` / ``up`` are (W8, scale_b, fp_weight)
    triples of the two projections, ``ind`` their common outlier index vector."""
    M, K = A.shape
    N = Out.shape[-1]
    tg = make_tensors(A, gate[0], gate[1], gate[2], ind, Out)
    tu = make_tensors(A, up[0], up[1], up[2], ind, None)
    check(load().mixq_enqueue_gated(ctypes.byref(tg), ctypes.byref(tu), M, N, K, _ptr(workspace),
                                    workspace.numel() * workspace.element_size(), _opts(config, sm_limit), flags, _stream(stream)),
          "mixq_enqueue_gated")


def gemm_dequant_gated(A8, scale_a, fp_A, gate, up, Out, stream=None, scratch=None, config: int = 0, sm_limit: int = 0) -> None:
    """mixq_gemm_dequant_gated (stage 2 of the gated call); ``gate`` / ``up`` = (W8, scale_b, fp_weight or None)."""
    M, K = A8.shape
    N = Out.shape[-1]
    check(load().mixq_gemm_dequant_gated(_ptr(A8), _ptr(scale_a), _ptr(fp_A), _ptr(gate[0]), _ptr(gate[1]), _ptr(gate[2]),
                                         _ptr(up[0]), _ptr(up[1]), _ptr(up[2]), _ptr(Out), M, N, K, _opts(config, sm_limit),
                                         _ptr(scratch), scratch.numel() * scratch.element_size() if scratch is not None else 0,
                                         _stream(stream)), "mixq_gemm_dequant_gated")


def quant_extract(A, ind, A8, scale_a, fp_A, flags: int = 0, stream=None) -> None:
    M, K = A.shape
    n_ind = 0 if ind is None else ind.numel()
    check(load().mixq_quant_extract(_ptr(A), M, K, _ptr(ind), n_ind, _ptr(A8), _ptr(scale_a), _ptr(fp_A), flags,
                                    _stream(stream)), "mixq_quant_extract")


def rmsnorm_quant_extract(X, gamma, eps, ind, A8, scale_a, fp_A, Y=None, flags: int = 0, stream=None) -> None:
    M, K = X.shape
    n_ind = 0 if ind is None else ind.numel()
    check(load().mixq_rmsnorm_quant_extract(_ptr(X), _ptr(gamma), float(eps), M, K, _ptr(ind), n_ind, _ptr(A8),
                                            _ptr(scale_a), _ptr(fp_A), _ptr(Y), flags, _stream(stream)),
          "mixq_rmsnorm_quant_extract")


def gemm_dequant(A8, W8, scale_a, scale_b, fp_A, fp_weight, Out, stream=None, workspace=None, bias=None,
                 activation: int = 0, config: int = 0, sm_limit: int = 0) -> None:
    """Stage 2.  ``config`` / ``sm_limit`` are the per-call mixq_options (tests pin every tile configuration)."""
    M, K = A8.shape
    N = W8.shape[0]
    if config or sm_limit:
        e = Epilogue(bias.data_ptr() if bias is not None else None, int(activation))
        check(load().mixq_gemm_dequant_opt(_ptr(A8), _ptr(W8), _ptr(scale_a), _ptr(scale_b), _ptr(fp_A), _ptr(fp_weight),
                                           _ptr(Out), M, N, K, ctypes.byref(e), _opts(config, sm_limit), _ptr(workspace),
                                           workspace.numel() * workspace.element_size() if workspace is not None else 0,
                                           _stream(stream)), "mixq_gemm_dequant_opt")
        return
    if bias is not None or activation:
        e = Epilogue(bias.data_ptr() if bias is not None else None, int(activation))
        check(load().mixq_gemm_dequant_ex(_ptr(A8), _ptr(W8), _ptr(scale_a), _ptr(scale_b), _ptr(fp_A), _ptr(fp_weight),
                                          _ptr(Out), M, N, K, ctypes.byref(e), _stream(stream)), "mixq_gemm_dequant_ex")
        return
    if workspace is None:
        check(load().mixq_gemm_dequant(_ptr(A8), _ptr(W8), _ptr(scale_a), _ptr(scale_b), _ptr(fp_A), _ptr(fp_weight),
                                       _ptr(Out), M, N, K, _stream(stream)), "mixq_gemm_dequant")
    else:
        check(load().mixq_gemm_dequant_ws(_ptr(A8), _ptr(W8), _ptr(scale_a), _ptr(scale_b), _ptr(fp_A),
                                          _ptr(fp_weight), _ptr(Out), M, N, K, _ptr(workspace),
                                          workspace.numel() * workspace.element_size(), _stream(stream)),
              "mixq_gemm_dequant_ws")


def linears_host(tensor_tables, A_host, outs_host, dev_scratch, flags: int = 0, stream=None) -> None:
    """mixq_linears_host: ``tensor_tables`` are Tensors structs of linears that share the pinned-host activations A_host
    [M, K]; outs_host[i] is the pinned-host output [M, N_i] of linear i."""
    M, K = A_host.shape
    n = len(tensor_tables)
    tt = (ctypes.POINTER(Tensors) * n)(*[ctypes.pointer(t) for t in tensor_tables])
    oo = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs_host])
    nn = (ctypes.c_int64 * n)(*[o.shape[-1] for o in outs_host])
    check(load().mixq_linears_host(tt, n, ctypes.c_void_p(A_host.data_ptr()), oo, M, nn, K, _ptr(dev_scratch),
                                   dev_scratch.numel() * dev_scratch.element_size(), flags, _stream(stream)), "mixq_linears_host")


def gated_host(gate_table, up_table, A_host, out_host, dev_scratch, flags: int = 0, stream=None) -> None:
    """mixq_gated_host: pinned-host A [M, K] -> Out_host [M, N] = fp16(silu(gate(A))) * fp16(up(A))"""
    M, K = A_host.shape
    check(load().mixq_gated_host(ctypes.byref(gate_table), ctypes.byref(up_table), ctypes.c_void_p(A_host.data_ptr()),
                                 ctypes.c_void_p(out_host.data_ptr()), M, out_host.shape[-1], K, _ptr(dev_scratch),
                                 dev_scratch.numel() * dev_scratch.element_size(), flags, _stream(stream)), "mixq_gated_host")


def host_drain(stream=None) -> None:
    """mixq_host_drain: wait for every host-buffer call issued with FLAG_HOST_ASYNC by this thread."""
    check(load().mixq_host_drain(_stream(stream)), "mixq_host_drain")


def gated_host_scratch_size(M: int, N: int, K: int) -> int:
    return int(load().mixq_gated_host_scratch_size(M, N, K))


def linears_host_scratch_size(M: int, Ns, K: int) -> int:
    nn = (ctypes.c_int64 * len(Ns))(*Ns)
    return int(load().mixq_linears_host_scratch_size(M, nn, len(Ns), K))


def make_peer_group(world: int, rank: int, out_ptrs, staging_ptrs, counter_ptrs, staging_bytes: int, counter_bytes: int,
                    out_multicast: int = 0) -> PeerGroup:
    """struct mixq_peer_group from raw addresses (ints) of every rank's Out / staging / counter buffers."""
    g = PeerGroup()
    g.world, g.rank = world, rank
    g.out_multicast = out_multicast or None
    for i in range(world):
        g.out[i], g.staging[i], g.counters[i] = out_ptrs[i], staging_ptrs[i], counter_ptrs[i]
    g.staging_bytes, g.counter_bytes = staging_bytes, counter_bytes
    return g


def enqueue_allreduce(A, W8, scale_b, fp_weight, ind, workspace, group: PeerGroup, flags: int = 0, stream=None,
                      sm_limit: int = 0, config: int = 0) -> None:
    """mixq_enqueue_allreduce: the result lands in every rank's Out buffer of ``group``.  ``config=9`` keeps the one-kernel
    path where the pull path (small decode-sized results) would be taken."""
    M, K = A.shape
    N = W8.shape[0]
    t = make_tensors(A, W8, scale_b, fp_weight, ind, None)
    check(load().mixq_enqueue_allreduce_opt(ctypes.byref(t), M, N, K, _ptr(workspace), workspace.numel() * workspace.element_size(),
                                            ctypes.byref(group), _opts(config, sm_limit), flags, _stream(stream)), "mixq_enqueue_allreduce")


def gemm_dequant_allreduce(A8, W8, scale_a, scale_b, fp_A, fp_weight, group: PeerGroup, stream=None, sm_limit: int = 0,
                           config: int = 0) -> None:
    M, K = A8.shape
    N = W8.shape[0]
    check(load().mixq_gemm_dequant_allreduce_opt(_ptr(A8), _ptr(W8), _ptr(scale_a), _ptr(scale_b), _ptr(fp_A), _ptr(fp_weight),
                                                 M, N, K, ctypes.byref(group), _opts(config, sm_limit), _stream(stream)),
          "mixq_gemm_dequant_allreduce")


def gemv_w8a16(A, q_weight, scales, Out, stream=None) -> None:
    """mixq_gemv_w8a16: the M <= 4 weight-only branch alone."""
    M, K = A.shape
    N = Out.shape[-1]
    check(load().mixq_gemv_w8a16(_ptr(A), _ptr(q_weight), _ptr(scales), _ptr(Out), M, N, K, _stream(stream)), "mixq_gemv_w8a16")
