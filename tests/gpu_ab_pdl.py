"""A/B of the early PDL release (mixq_set_pdl_early_rows) on graph-replayed decode steps, same process, interleaved.
   python tests/gpu_ab_pdl.py M [rows thresholds...]"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402
M = int(sys.argv[1])
ths = [int(x) for x in sys.argv[2:]] or [0, 100000]
dev = "cuda"
lib = B.load()
shapes = [(12288, 4096), (4096, 4096), (11008, 4096), (11008, 4096), (4096, 11008)]
g = torch.Generator(device=dev).manual_seed(0)
lins = []
for N, K in shapes:
    W8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g)
    ind = torch.randperm(K, device=dev, generator=g)[:128].int()
    sb = (torch.rand(N, device=dev, generator=g) * 2e-4 + 1e-4).half()
    fw = (torch.randn(N, 128, device=dev, generator=g) * 0.02).half()
    lins.append((W8, sb, fw, ind, N, K))
acts = {K: torch.randn(M, K, device=dev, generator=g).half() for K in (4096, 11008)}
out = torch.empty(M * 12288, dtype=torch.float16, device=dev)
ws = torch.empty(B.workspace_size(M, 12288, 11008), dtype=torch.uint8, device=dev)


def step():
    for W8, sb, fw, ind, N, K in lins:
        B.enqueue(acts[K], W8, sb, fw, ind, out[: M * N].view(M, N), ws)


graphs = {}
for th in ths:
    lib.mixq_set_pdl_early_rows(th)
    gs = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.stream(gs):
        step(); gs.synchronize()
        with torch.cuda.graph(gr, stream=gs):
            step()
    torch.cuda.synchronize()
    graphs[th] = gr
res = {th: [] for th in ths}
for r in range(8):
    for th in ths:
        gr = graphs[th]
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        if r:
            res[th].append(e0.elapsed_time(e1) * 10)
for th in ths:
    print(f"M={M} early_rows={th}: median {np.median(res[th]):.1f} us/step  min {min(res[th]):.1f}")
