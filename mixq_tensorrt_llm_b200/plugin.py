"""Python side of the drop-in boundary: ``mixgemm`` and ``MixQLinear``.

Mirrors the reference's plugin.py (reference plugin.py:34-47 library load +
initOpenAiTritonPlugins, :52-79 ``mixgemm``, :86-162 ``MixQLinear``) with torch CUDA tensors in
place of TensorRT network tensors: same constructor arguments, same parameter names, shapes and
fp16-typed containers (so a reference ``int8_mix`` checkpoint maps 1:1), same seven plugin
inputs in the same order.  The call goes Python -> C handle API -> MixQPlugin::enqueue ->
mixq_enqueue -> two CUDA kernels.  No fallback: without the built library or without a B200
this module raises.

Differences from the reference, on purpose:
  * tensor parallelism: the reference all-reduces after *every* MixQLinear that has a tp_group
    (plugin.py:155-156), which is wrong for the column-parallel linears it wraps, and forbids
    row-parallel (quantization/quantize.py:342).  Here ``parallel_mode="column"`` needs no
    collective (optionally all-gathers when gather_output) and ``parallel_mode="row"`` does the
    single all-reduce of the partial sums.
  * the M<=4 weight-only branch (TsinghuaMixQPlugin.cpp:472) is taken when ``qweight`` has been loaded; an empty
    ``qweight`` (a NULL plugin input 5) keeps the mixed path for every M.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch

from . import binding

TRT_LLM_PLUGIN_NAMESPACE = "tensorrt_llm"
LAYER_NAME = "MixQLayer"
NUM_OUTLIERS = binding.NUM_OUTLIERS

_registered = False


def _load_plugin_lib() -> None:
    """plugin.py:34-47 -- dlopen the library and register the creator under 'tensorrt_llm'."""
    global _registered
    if _registered:
        return
    lib = binding.load()
    assert lib.initOpenAiTritonPlugins(None, TRT_LLM_PLUGIN_NAMESPACE.encode("utf-8"))
    _registered = True


class _PluginHandle:
    """Owns one MixQPlugin instance created through the registered creator ('MixQ','1',ns)."""

    def __init__(self, m: int, n: int, k: int):
        _load_plugin_lib()
        self._lib = binding.load()
        self._h = self._lib.mixq_plugin_create(TRT_LLM_PLUGIN_NAMESPACE.encode(), int(m), int(n), int(k))
        if not self._h:
            raise binding.MixQError("plugin creator ('MixQ','1','tensorrt_llm') is not registered")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.mixq_plugin_destroy(h)

    def enqueue(self, a_dims, n: int, inputs, out, workspace, stream) -> None:
        dims = (ctypes.c_int64 * len(a_dims))(*a_dims)
        in_ptrs = (ctypes.c_void_p * 7)(*[t.data_ptr() for t in inputs])
        out_ptrs = (ctypes.c_void_p * 1)(out.data_ptr())
        rc = self._lib.mixq_plugin_enqueue(self._h, dims, len(a_dims), int(n), in_ptrs, out_ptrs,
                                           ctypes.c_void_p(workspace.data_ptr()), ctypes.c_void_p(stream.cuda_stream))
        binding.check(rc, "MixQPlugin::enqueue")


_workspaces: dict = {}
_retired: list = []


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Scratch playing the role of the workspace TensorRT hands to enqueue: one buffer per (device, stream), so linears
    running on different streams never share A8 / scale_a / fp_A, grow-only, and a buffer that has been outgrown is kept
    alive (a captured CUDA graph may still hold its address) instead of being freed."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _retired.append(ws)
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def mixgemm(m: int, n: int, k: int, inputs: List[torch.Tensor], plugin: Optional[_PluginHandle] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """plugin.py:52-79.  ``inputs`` are the seven plugin inputs in the reference's order:
    [A, weight(int8 bytes as fp16 [N,K/2]), weights_scaling_factor [N], fp_weight [N,128],
     fp_ind (int32 bytes as fp16 [256]), qweight, scaling_factors]."""
    binding.require_device()
    A = inputs[0]
    if not (A.is_cuda and A.dtype == torch.float16):
        raise binding.MixQError("mixgemm: activation must be a CUDA fp16 tensor")
    if not A.is_contiguous():
        A = A.contiguous()
    if plugin is None:
        plugin = _PluginHandle(m, n, k)
    a_dims = list(A.shape)
    M = 1
    for d in a_dims[:-1]:
        M *= d
    K = a_dims[-1]
    N = inputs[1].shape[0]
    if out is None:
        out = torch.empty(*a_dims[:-1], N, dtype=torch.float16, device=A.device)
    ws = _workspace(A.device, binding.workspace_size(max(M, 1), N, K))
    plugin.enqueue(a_dims, N, [A] + list(inputs[1:7]), out, ws, torch.cuda.current_stream(A.device))
    return out


class MixQLinear(torch.nn.Module):
    """reference plugin.py:86-162, torch flavour.

    Parameters hold exactly the reference's checkpoint tensors (all typed float16):
      weight [N/tp, K/2]      int8 codes, two per fp16 slot  (model_config_utils.py:460-466)
      fp_weight [N/tp, 128]   outlier weight columns          (:452)
      fp_ind [256]            128 int32 indices as raw bytes  (:455-457)
      qweight [K, N/tp/2]     weight-only layout for M<=4     (:437-441; optional, see load_packed)
      weights_scaling_factor [N/tp]                            (:429-430)
    """

    def __init__(self, in_features: int, out_features: int, bias: bool = False, dtype=None, tp_group=None,
                 tp_size: int = 1, gather_output: bool = True, parallel_mode: str = "column", device=None):
        super().__init__()
        if parallel_mode not in ("column", "row"):
            raise ValueError("parallel_mode must be 'column' or 'row'")
        self.parallel_mode = parallel_mode
        self.tp_size = tp_size
        self.tp_group = tp_group
        self.gather_output = gather_output
        if parallel_mode == "column":
            self.in_features = in_features
            self.out_features = out_features // tp_size
        else:
            self.in_features = in_features // tp_size
            self.out_features = out_features
        h = dict(dtype=torch.float16, device=device)
        N, K = self.out_features, self.in_features
        self.register_buffer("weight", torch.zeros(N, K // 2, **h))
        self.register_buffer("fp_weight", torch.zeros(N, NUM_OUTLIERS, **h))
        self.register_buffer("fp_ind", torch.zeros(NUM_OUTLIERS * 2, **h))
        self.register_buffer("qweight", torch.zeros(0, **h))  # empty = NULL plugin input 5: mixed path for every M
        self.register_buffer("weights_scaling_factor", torch.zeros(N, **h))
        if bias:
            self.register_buffer("bias", torch.zeros(N, dtype=dtype or torch.float16, device=device))
        else:
            self.bias = None
        self._plugin = None
        self._peer = None

    def attach_peer_buffers(self, peer) -> "MixQLinear":
        """Row-parallel only: fuse the all-reduce into the GEMM kernel (mixq_enqueue_allreduce) using the symmetric
        buffers in ``peer`` (mixq_tensorrt_llm_b200.peer.PeerBuffers).  The returned tensor of forward() is then a
        view of peer's Out buffer, valid until the next fused call that uses the same buffers."""
        if self.parallel_mode != "row":
            raise ValueError("peer buffers are used by row-parallel linears only")
        self._peer = peer
        return self

    @torch.no_grad()
    def load_packed(self, W8: torch.Tensor, scale_b: torch.Tensor, fp_weight: torch.Tensor, ind: torch.Tensor,
                    bias: Optional[torch.Tensor] = None, qweight: Optional[torch.Tensor] = None) -> "MixQLinear":
        """Fill the buffers from typed tensors (int8 [N,K], fp16 [N], fp16 [N,128], int32 [128]); ``qweight`` is the
        processed EETQ tensor (int8 [K,N]): with it, calls with M <= 4 take the weight-only branch like the reference
        (TsinghuaMixQPlugin.cpp:472), scaled by weights_scaling_factor as plugin.py:149 wires it."""
        if qweight is not None:
            self.qweight = qweight.contiguous().view(torch.float16).reshape(self.in_features, self.out_features // 2).to(self.weight.device)
        self.weight.copy_(W8.contiguous().view(torch.float16))
        self.weights_scaling_factor.copy_(scale_b.reshape(-1))
        self.fp_weight.copy_(fp_weight)
        self.fp_ind.copy_(ind.to(torch.int32).contiguous().view(torch.float16))
        if bias is not None and self.bias is not None:
            self.bias.copy_(bias)
        return self

    def forward(self, A: torch.Tensor, activation: Optional[str] = None, fuse_bias: bool = False) -> torch.Tensor:
        """``activation="silu"`` and/or ``fuse_bias=True`` run the fused epilogue (mixq_enqueue_ex: SiLU in fp32 before the
        output rounding, bias added to the fp16 result) instead of separate elementwise passes; not for row-parallel
        shards, whose partial results must be reduced first (column-parallel shards are gathered afterwards)."""
        if activation is not None or fuse_bias:
            if activation not in (None, "silu"):
                raise ValueError("activation must be None or 'silu'")
            if self.tp_size > 1 and self.parallel_mode == "row":
                raise ValueError("the fused epilogue cannot run on row-parallel partial sums")
            binding.require_device()
            M = A.numel() // A.shape[-1]
            A2 = A.reshape(M, A.shape[-1]).contiguous()
            out = torch.empty(M, self.out_features, dtype=torch.float16, device=A.device)
            ws = _workspace(A.device, binding.workspace_size(max(M, 1), self.out_features, self.in_features))
            binding.enqueue(A2, self.weight.view(torch.int8).view(self.out_features, self.in_features),
                            self.weights_scaling_factor, self.fp_weight, self.fp_ind.view(torch.int32), out, ws,
                            q_weight=self.qweight if self.qweight.numel() else None,
                            scaling_factors=self.weights_scaling_factor,
                            bias=self.bias.to(torch.float16) if (fuse_bias and self.bias is not None) else None,
                            activation=binding.ACT_SILU if activation == "silu" else binding.ACT_NONE)
            x = out.view(*A.shape[:-1], self.out_features)
            if self.bias is not None and not fuse_bias:
                x = x + self.bias.to(x.dtype)
            return self._gather_columns(x)      # the bias shard was added to its own columns before the gather
        if self._peer is not None and self.tp_size > 1:
            binding.require_device()
            M = A.numel() // A.shape[-1]
            A2 = A.reshape(M, A.shape[-1]).contiguous()
            ws = _workspace(A.device, binding.workspace_size(max(M, 1), self.out_features, self.in_features))
            binding.enqueue_allreduce(A2, self.weight.view(torch.int8).view(self.out_features, self.in_features),
                                      self.weights_scaling_factor, self.fp_weight, self.fp_ind.view(torch.int32), ws,
                                      self._peer.peer_group(M, self.out_features))
            x = self._peer.out(M, self.out_features).view(*A.shape[:-1], self.out_features)
            if self.bias is not None:
                x = x + self.bias.to(x.dtype)
            return x
        if self._plugin is None:
            self._plugin = _PluginHandle(A.shape[0], self.out_features, self.in_features)
        x = mixgemm(A.shape[0], self.out_features, self.in_features,
                    [A, self.weight, self.weights_scaling_factor, self.fp_weight, self.fp_ind, self.qweight,
                     self.weights_scaling_factor], plugin=self._plugin)
        if self.tp_size > 1 and self.tp_group is not None and self.parallel_mode == "row":
            import torch.distributed as dist
            dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.tp_group)  # the one exchange step of the path
        if self.bias is not None:
            # outside the plugin, as plugin.py:158-160.  Column-parallel: the bias buffer is this rank's [N / tp] shard and
            # belongs to this rank's output columns, so it is added BEFORE the gather (TensorRT-LLM's ColumnLinear order);
            # row-parallel: the full [N] bias is added once, after the reduction.
            x = x + self.bias.to(x.dtype)
        return self._gather_columns(x)

    def _gather_columns(self, x: torch.Tensor) -> torch.Tensor:
        """column-parallel with gather_output: concatenate the ranks' output-channel shards"""
        if self.parallel_mode != "column" or self.tp_size <= 1 or self.tp_group is None or not self.gather_output:
            return x
        import torch.distributed as dist
        parts = [torch.empty_like(x) for _ in range(self.tp_size)]
        dist.all_gather(parts, x.contiguous(), group=self.tp_group)
        return torch.cat(parts, dim=-1)


class MixQSrcLinear(torch.nn.Module):
    """The reference's torch-side module, MixLinear_GEMM (MixQ/src/mixquant/modules/linear.py), 8-bit branch with
    ``unfused=True``: weights quantised WITHOUT outlier handling at construction (:110-118, ``ind`` empty :42) and outlier
    columns discovered at run time -- while ``add_outliers`` holds, a batch whose per-token scale exceeds sigma / 127 adds
    the columns holding a value above sigma (:197-223).  Known outlier columns are zeroed in the activations before the
    per-token quantisation (MIXQ_FLAG_MASK_OUTLIERS = cult.cu:1588) and their weights are the DEQUANTISED codes.
    The device work is this library's two kernels; the growth logic is host code, as it is in the reference.
    The kernels hold 128 outlier columns (the plugin's num_ind), so growth beyond that raises."""

    def __init__(self, weight: torch.Tensor, sigma: float = 6.0, stop: int = 2, bias: Optional[torch.Tensor] = None):
        super().__init__()
        W = weight.detach().to(torch.float16)
        scale = (W.abs().amax(dim=1, keepdim=True) / 127).to(torch.float16)
        self.register_buffer("scale_col", scale.reshape(-1).contiguous())
        self.register_buffer("q_weight", (W / scale).round().nan_to_num(0).to(torch.int8).contiguous())
        self.register_buffer("ind", torch.zeros(0, dtype=torch.int32, device=W.device))
        self.register_buffer("weight_cache", torch.zeros(W.shape[0], 0, dtype=torch.float16, device=W.device))
        self.bias = bias
        self.sigma = torch.tensor(sigma, dtype=torch.float16, device=W.device)
        self.stop, self.cnt, self.add_outliers = stop, 0, True
        self.out_features, self.in_features = W.shape

    def _padded(self):
        """(ind [128] int32, fp_weight [N,128]) for the kernels: unused slots repeat the first outlier column with zero weights"""
        n = self.ind.numel()
        ind = torch.cat([self.ind, self.ind[:1].expand(NUM_OUTLIERS - n)]).contiguous()
        fw = torch.zeros(self.out_features, NUM_OUTLIERS, dtype=torch.float16, device=self.q_weight.device)
        fw[:, :n] = self.weight_cache
        return ind, fw

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        binding.require_device()
        M = x.numel() // x.shape[-1]
        x2 = x.reshape(M, self.in_features).to(torch.float16).contiguous()
        dev = x2.device
        A8 = torch.empty(M, self.in_features, dtype=torch.int8, device=dev)
        sa = torch.empty(M, dtype=torch.float16, device=dev)
        fpA = torch.empty(M, NUM_OUTLIERS, dtype=torch.float16, device=dev)

        def stage1():
            if self.ind.numel():
                binding.quant_extract(x2, self._padded()[0], A8, sa, fpA, flags=binding.FLAG_MASK_OUTLIERS)
            else:
                binding.quant_extract(x2, None, A8, sa, None)
        stage1()
        if self.add_outliers:
            if bool(sa.max() > self.sigma / 127):                       # linear.py:198
                xm = x2.clone()
                if self.ind.numel():
                    xm[:, self.ind.long()] = 0
                new = torch.unique(torch.where(xm.abs() > self.sigma)[1]).to(torch.int32)      # FindOutliers, :154-159
                if self.ind.numel() + new.numel() > NUM_OUTLIERS:
                    raise binding.MixQError(f"MixQSrcLinear: {self.ind.numel() + new.numel()} outlier columns; the kernels hold {NUM_OUTLIERS}")
                w_new = (self.q_weight[:, new.long()].to(torch.float16) * self.scale_col[:, None]).to(torch.float16)   # :203-204
                self.weight_cache = torch.cat([self.weight_cache, w_new], dim=1)
                self.ind = torch.cat([self.ind, new])
                stage1()                                                 # re-quantise with the new columns zeroed, :219
            self.cnt += 1
            if self.cnt >= self.stop or self.ind.numel() > 256:
                self.add_outliers = False
        out = torch.empty(M, self.out_features, dtype=torch.float16, device=dev)
        if self.ind.numel():
            binding.gemm_dequant(A8, self.q_weight, sa, self.scale_col, fpA, self._padded()[1], out)
        else:
            binding.gemm_dequant(A8, self.q_weight, sa, self.scale_col, None, None, out)
        y = out.view(*x.shape[:-1], self.out_features)
        return y + self.bias.to(y.dtype) if self.bias is not None else y


class MixQLlamaMLP(torch.nn.Module):
    """reference MixQ/src/mixquant/modules/fused/mlp.py:36-70 (MixLlamaMLP), over three MixQLinear modules:
        y = down_proj( silu(gate_proj(x)) * up_proj(x) )
    gate and up read the same activations with the same outlier columns; they run as ONE mixq_enqueue_gated call (one
    quantise launch, one GEMM launch whose epilogue applies the SiLU and the product), then the down projection."""

    def __init__(self, gate_proj: MixQLinear, down_proj: MixQLinear, up_proj: MixQLinear):
        super().__init__()
        if (gate_proj.in_features, gate_proj.out_features) != (up_proj.in_features, up_proj.out_features):
            raise ValueError("gate_proj and up_proj must have the same shape")
        if gate_proj.bias is not None or up_proj.bias is not None:
            raise ValueError("the gated call carries no bias (Llama / Qwen2 MLPs have none)")
        self.gate_proj_, self.down_proj_, self.up_proj_ = gate_proj, down_proj, up_proj
        self.out_features = down_proj.out_features

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        binding.require_device()
        g, u = self.gate_proj_, self.up_proj_
        if not torch.equal(g.fp_ind, u.fp_ind):
            raise binding.MixQError("MixQLlamaMLP: gate_proj and up_proj must share their outlier columns (same activation statistics)")
        M = x.numel() // x.shape[-1]
        x2 = x.reshape(M, g.in_features).contiguous()
        h = torch.empty(M, g.out_features, dtype=torch.float16, device=x.device)
        ws = _workspace(x.device, binding.gated_workspace_size(max(M, 1), g.out_features, g.in_features))

        def triple(m):
            return (m.weight.view(torch.int8).view(m.out_features, m.in_features), m.weights_scaling_factor, m.fp_weight)
        binding.enqueue_gated(x2, triple(g), triple(u), g.fp_ind.view(torch.int32), h, ws)
        return self.down_proj_(h.view(*x.shape[:-1], g.out_features))
