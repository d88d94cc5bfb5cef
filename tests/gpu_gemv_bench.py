"""Timing of the M <= 4 weight-only GEMV: ours vs the reference kernel (oracle/_ref/libref_gemv.so), HBM roofline.
   python tests/gpu_gemv_bench.py"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402
import refgpu  # noqa: E402

dev = "cuda"
for (N, K) in [(12288, 4096), (4096, 4096), (11008, 4096), (4096, 11008), (28672, 8192)]:
    # rotate over enough weight copies to exceed the 126 MB L2
    copies = max(2, int(300e6 // (N * K)) + 1)
    qs = [torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev) for _ in range(copies)]
    sc = (torch.rand(N, device=dev) * 1e-3 + 1e-4).half()
    for M in (1, 4):
        A = torch.randn(M, K, device=dev).half()
        out = torch.empty(M, N, dtype=torch.float16, device=dev)
        res = {}
        for name, fn in (("ours", lambda q: B.gemv_w8a16(A, q, sc, out)),
                         ("ref", (lambda q: refgpu.gemv(A, q, sc)) if refgpu.gemv_available() else None)):
            if fn is None:
                continue
            for q in qs:
                fn(q)
            torch.cuda.synchronize()
            reps = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                for q in qs:
                    fn(q)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * copies)
            res[name] = us
        gbs = {k: round(N * K / v / 1e3, 1) for k, v in res.items()}
        print(f"N={N} K={K} M={M}: us {dict((k, round(v, 2)) for k, v in res.items())}  GB/s {gbs}", flush=True)
