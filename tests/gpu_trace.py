"""Per-CTA timeline of the stream-K kernel (debug).  python tests/gpu_trace.py M N K"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402
M, N, K = (int(x) for x in sys.argv[1:4])
cfg = int(sys.argv[4]) if len(sys.argv) > 4 else 8
dev = "cuda"
lib = B.load()
A8 = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
W8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
sa = (torch.rand(M, device=dev) * 0.01 + 1e-3).half()
sb = (torch.rand(N, device=dev) * 0.002 + 1e-4).half()
import os
fpA = torch.randn(M, 128, device=dev).half() if not os.environ.get("NO_OUTLIER") else None
fw = (torch.randn(N, 128, device=dev) * 0.02).half() if not os.environ.get("NO_OUTLIER") else None
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = torch.zeros(lib.mixq_decode_workspace_size(min(M, 1024), N), dtype=torch.uint8, device=dev)
trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
for it in range(4):
    B.gemm_dequant(A8, W8, sa, sb, fpA, fw, out, workspace=ws, config=cfg)
torch.cuda.synchronize()
B.check(lib.mixq_debug_set_trace(trace.data_ptr()), "trace")
B.gemm_dequant(A8, W8, sa, sb, fpA, fw, out, workspace=ws, config=cfg)
torch.cuda.synchronize()
lib.mixq_debug_set_trace(None)
t = trace.cpu().numpy().reshape(148, 16).astype(np.float64)
t0 = t[:, 0][t[:, 0] > 0].min()
names = ["entry", "prologue", "firstTMA", "tile0issued", "mmaDone", "acc0ready", "accLast", "epiDone",
         "peersIn", "chunksDone", "loopExit", "storesDone", "fDrained", "s13", "s14", "s15"]
rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
print("shape", M, N, K, "cfg", cfg, "us relative to first CTA entry")
for i, n in enumerate(names):
    col = rel[:, i]
    ok = ~np.isnan(col)
    if ok.any():
        print(f"{n:12s} min {np.nanmin(col):7.2f} med {np.nanmedian(col):7.2f} max {np.nanmax(col):7.2f}  (n={ok.sum()})")
for c in (0, 2, 72, 146):
    print("cta", c, np.round(rel[c][:13], 2).tolist())
