"""Small launch sequence for ncu: a few quant + GEMM launches of one shape/config.
   python tests/gpu_profile_target.py CFG M N K [iters]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402

cfg, M, N, K = (int(x) for x in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = "cuda"
B.require_device()
A = (torch.randn(M, K, device=dev) * 0.5).half()
W8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
sb = (torch.rand(N, device=dev) * 0.002 + 1e-4).half()
fw = (torch.randn(N, 128, device=dev) * 0.02).half()
ind = torch.randperm(K, device=dev)[:128].int()
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = torch.empty(B.workspace_size(M, N, K, config=cfg), dtype=torch.uint8, device=dev)
for _ in range(iters):
    B.enqueue(A, W8, sb, fw, ind, out, ws, config=cfg)
torch.cuda.synchronize()
print("done", cfg, M, N, K)
