"""Layer-stack decode harness (tokens/sec (ii)): a Llama-2-7B-shaped decoder stack (32 layers, hidden 4096,
32 heads, FFN 11008; synthetic weights, random-init) whose linears are the MixQ W8A8O16 path, stepping a batch of
`bs` tokens with a KV window of `kv` positions.  Clearly NOT the TensorRT engine (which cannot run here): attention is
torch SDPA (library), rotary embedding is skipped, the KV window has a fixed length so that the step is one CUDA graph.
What it does measure is the linears in context, with everything the library offers around the plugin path:

  ours       RMSNorm -> outlier extract -> INT8 quantise in ONE pass (mixq_rmsnorm_quant_extract), gate and up projections
             with SiLU and their product in ONE GEMM launch (mixq_gemm_dequant_gated);
  plugin     the same stack through the plugin contract only: torch RMSNorm, then one mixq_enqueue per linear;
  reference  torch RMSNorm, then the reference's own kernels per linear (oracle/_ref: gather, cuBLAS fp16, int8quant,
             CUTLASS GemmDequant -- 4 launches), when oracle/_ref is on the box.

    python tests/gpu_layer_stack.py [bs] [kv] [layers]      -> one JSON line (tokens/s = bs / median step time)
kv defaults to 2048 (SURVEY.md 8d ii) where the KV cache fits: 2 x 32 layers x 4096 x kv x bs fp16 values are 34 GB at bs 32
but 1.1 TB at bs 512, so larger batches get the longest window that keeps the cache under ~70 GB (256 positions at bs 512).
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402
import refgpu  # noqa: E402

H, HEADS, FFN, EPS = 4096, 32, 11008, 1e-5


def make_linear(N, K, g, dev):
    W8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g)
    ind = torch.randperm(K, device=dev, generator=g)[:128].int()
    W8[:, ind.long()] = 0
    sb = (torch.rand(N, device=dev, generator=g) * 1e-4 + 5e-5).half()
    fw = (torch.randn(N, 128, device=dev, generator=g) * 0.01).half()
    return dict(W8=W8, sb=sb, fw=fw, ind=ind, N=N, K=K)


def main():
    bs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n_layers = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    kv_fit = int(70e9 / (2 * n_layers * H * bs * 2))
    kv = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else min(2048, max(64, 1 << (kv_fit.bit_length() - 1)))
    dev = "cuda"
    B.require_device()
    lib = B.load()
    g = torch.Generator(device=dev).manual_seed(0)
    layers = []
    for _ in range(n_layers):
        L = dict(qkv=make_linear(3 * H, H, g, dev), o=make_linear(H, H, g, dev), gate=make_linear(FFN, H, g, dev),
                 up=make_linear(FFN, H, g, dev), down=make_linear(H, FFN, g, dev),
                 g1=(1 + 0.1 * torch.randn(H, device=dev, generator=g)).half(),
                 g2=(1 + 0.1 * torch.randn(H, device=dev, generator=g)).half(),
                 k=torch.randn(bs, HEADS, kv, H // HEADS, device=dev, generator=g).half(),
                 v=torch.randn(bs, HEADS, kv, H // HEADS, device=dev, generator=g).half())
        L["up"]["ind"] = L["gate"]["ind"]          # gate and up read the same activations: one outlier set (same act scales)
        L["up"]["W8"][:, L["gate"]["ind"].long()] = 0
        layers.append(L)
    x0 = torch.randn(bs, H, device=dev, generator=g).half()
    ws = torch.empty(B.workspace_size(bs, FFN, FFN), dtype=torch.uint8, device=dev)
    A8 = torch.empty(bs, FFN, dtype=torch.int8, device=dev)
    sa = torch.empty(bs, dtype=torch.float16, device=dev)
    fpA = torch.empty(bs, 128, dtype=torch.float16, device=dev)
    qkv = torch.empty(bs, 3 * H, dtype=torch.float16, device=dev)
    o = torch.empty(bs, H, dtype=torch.float16, device=dev)
    gate = torch.empty(bs, FFN, dtype=torch.float16, device=dev)
    up = torch.empty(bs, FFN, dtype=torch.float16, device=dev)
    down = torch.empty(bs, H, dtype=torch.float16, device=dev)
    rws = None

    def rms(x, gamma):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + EPS) * gamma.float()).half()

    def attention(L, pos):
        q, k, v = qkv.view(bs, 3, HEADS, H // HEADS).unbind(1)
        L["k"][:, :, pos] = k
        L["v"][:, :, pos] = v
        return F.scaled_dot_product_attention(q.unsqueeze(2), L["k"], L["v"]).reshape(bs, H)

    def lin_plugin(p, a, out):
        B.enqueue(a, p["W8"], p["sb"], p["fw"], p["ind"], out, ws)

    def lin_ref(p, a, out):
        refgpu.enqueue(a, p["W8"], p["sb"], p["fw"], p["ind"], out, rws)

    def step(mode, pos=kv - 1):
        x = x0
        for L in layers:
            if mode == "ours":
                a8 = A8[:, :H]
                B.rmsnorm_quant_extract(x, L["g1"], EPS, L["qkv"]["ind"], a8, sa, fpA)
                p = L["qkv"]
                B.gemm_dequant(a8, p["W8"], sa, p["sb"], fpA, p["fw"], qkv)
                lin_plugin(L["o"], attention(L, pos), o)
                x = x + o
                B.rmsnorm_quant_extract(x, L["g2"], EPS, L["gate"]["ind"], a8, sa, fpA)
                pg, pu = L["gate"], L["up"]
                B.gemm_dequant_gated(a8, sa, fpA, (pg["W8"], pg["sb"], pg["fw"]), (pu["W8"], pu["sb"], pu["fw"]), gate)   # silu(gate) * up
                lin_plugin(L["down"], gate, down)
                x = x + down
            else:
                lin = lin_plugin if mode == "plugin" else lin_ref
                lin(L["qkv"], rms(x, L["g1"]), qkv)
                lin(L["o"], attention(L, pos), o)
                x = x + o
                h = rms(x, L["g2"])
                lin(L["gate"], h, gate)
                lin(L["up"], h, up)
                lin(L["down"], F.silu(gate) * up, down)
                x = x + down
        return x

    res = {}
    outs = {}
    modes = ["ours", "plugin"] + (["reference"] if refgpu.available() else [])
    if "reference" in modes:
        rws = torch.empty(refgpu.load().ref_workspace_size(bs, FFN, FFN), dtype=torch.uint8, device=dev)
    for mode in modes:
        torch.cuda.synchronize()
        outs[mode] = step(mode).float().clone()
        gs = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.stream(gs):
            step(mode)
            gs.synchronize()
            with torch.cuda.graph(graph, stream=gs):
                step(mode)
        torch.cuda.synchronize()
        for _ in range(3):
            graph.replay()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 5)
        ms = float(np.median(ts))
        res[mode] = {"ms_per_step": round(ms, 4), "tokens_per_s": round(bs / (ms * 1e-3), 1)}
    # the three stacks compute the same function up to rounding (fused RMSNorm: <= 1 fp16 ulp per element; SiLU intrinsics)
    for m in modes[1:]:
        d = (outs[m] - outs["ours"]).norm() / outs["ours"].norm()
        res[m]["rel_diff_vs_ours"] = float(d)
    line = {"what": "Llama-2-7B-shaped decoder stack, synthetic weights; linears = MixQ W8A8O16; attention = torch SDPA over a fixed "
                    "KV window; NOT the TensorRT engine", "bs": bs, "kv_window": kv, "layers": n_layers,
            "launch": "one CUDA graph per decode step", "results": res,
            "linear_flops_per_step": 2.0 * bs * n_layers * (3 * H * H + H * H + 3 * H * FFN)}
    if "reference" in res:
        line["speedup_vs_reference_kernels"] = round(res["reference"]["ms_per_step"] / res["ours"]["ms_per_step"], 3)
    line["speedup_vs_plugin_only"] = round(res["plugin"]["ms_per_step"] / res["ours"]["ms_per_step"], 3)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
