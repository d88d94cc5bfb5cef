"""Interleaved A/B timing of GEMM configurations (debug aid).  python tests/gpu_ab.py "6,9,8" M N K [rounds]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402

cfgs = [int(x) for x in sys.argv[1].split(",")]
M, N, K = (int(x) for x in sys.argv[2:5])
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 7
dev = "cuda"
lib = B.load()
A8 = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
W8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
sa = (torch.rand(M, device=dev) * 0.01 + 1e-3).half()
sb = (torch.rand(N, device=dev) * 0.002 + 1e-4).half()
fpA = torch.randn(M, 128, device=dev).half()
fw = (torch.randn(N, 128, device=dev) * 0.02).half()
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = torch.zeros(lib.mixq_decode_workspace_size(min(M, 1024), N), dtype=torch.uint8, device=dev)
reps = max(3, int(2e-3 / max(2.0 * M * N * K / 2.5e15, 1e-6)))   # ~2 ms of work per sample
res = {c: [] for c in cfgs}
for r in range(rounds + 1):
    for c in cfgs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        B.gemm_dequant(A8, W8, sa, sb, fpA, fw, out, workspace=ws, config=c)
        e0.record()
        for _ in range(reps):
            B.gemm_dequant(A8, W8, sa, sb, fpA, fw, out, workspace=ws, config=c)
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            res[c].append(e0.elapsed_time(e1) * 1e3 / reps)
print(f"shape {M}x{N}x{K} group_m={os.environ.get('MIXQ_GROUP_M', 'default')} reps={reps}")
for c in cfgs:
    us = np.array(res[c])
    print(f"  cfg{c}: median {np.median(us):9.1f} us  min {us.min():9.1f}  -> {2.0 * M * N * K / np.median(us) / 1e6:7.1f} TOPS")
