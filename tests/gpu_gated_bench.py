"""Timing of the gated call against the two separate plugin calls it replaces (debug aid, not a test).
    python tests/gpu_gated_bench.py [M N K [layers]]
Each variant is captured in a CUDA graph that walks `layers` sets of weights (so W streams from HBM) and replayed."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 11008, 4096)
L = int(sys.argv[4]) if len(sys.argv) > 4 else 8
dev = "cuda"
B.require_device()
g = torch.Generator(device=dev).manual_seed(1)
A = torch.randn(M, K, device=dev, generator=g).half()
ind = torch.randperm(K, device=dev, generator=g)[:128].int().sort()[0].int()


def lin():
    return (torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g),
            (torch.rand(N, device=dev, generator=g) * 0.002 + 1e-4).half(),
            (torch.randn(N, 128, device=dev, generator=g) * 0.02).half())


gates, ups = [lin() for _ in range(L)], [lin() for _ in range(L)]
ws = torch.empty(B.gated_workspace_size(M, N, K), dtype=torch.uint8, device=dev)
o1 = torch.empty(M, N, dtype=torch.float16, device=dev)
o2 = torch.empty(M, N, dtype=torch.float16, device=dev)


def fused(i):
    B.enqueue_gated(A, gates[i], ups[i], ind, o1, ws)


def fused_epi12(i):
    B.enqueue_gated(A, gates[i], ups[i], ind, o1, ws, config=13)


def separate(i):
    B.enqueue(A, *gates[i], ind, o1, ws, activation=B.ACT_SILU)
    B.enqueue(A, *ups[i], ind, o2, ws)


def separate_mul(i):
    separate(i)
    torch.mul(o1, o2, out=o1)


def timed(fn):
    for i in range(L):
        fn(i)
    torch.cuda.synchronize()
    s, gr = torch.cuda.Stream(), torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            for i in range(L):
                fn(i)
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            gr.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / (4 * L))
    return best


print(f"gated MLP input half {M}x{N}x{K}, {L} weight sets, us per layer (graph replay)")
for name, fn in (("mixq_enqueue_gated (1 quant + 1 GEMM)", fused), ("the same with 12 epilogue warps (config 13)", fused_epi12), ("2 x mixq_enqueue (gate with SiLU, up)", separate),
                 ("2 x mixq_enqueue + torch.mul", separate_mul)):
    us = timed(fn)
    print(f"  {name:42s} {us:8.2f} us   {4.0 * M * N * K / us / 1e6:7.1f} TOPS")
