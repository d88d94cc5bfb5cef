"""ncu target: the gated call (mixq_enqueue_gated: quantise kernel + fat-tile GEMM holding a gate tile and an up tile).
   python tests/gpu_profile_gated.py M N K [iters]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = "cuda"
B.require_device()
A = (torch.randn(M, K, device=dev) * 0.5).half()
ind = torch.randperm(K, device=dev)[:128].int()


def lin():
    return (torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev), (torch.rand(N, device=dev) * 0.002 + 1e-4).half(),
            (torch.randn(N, 128, device=dev) * 0.02).half())


gate, up = lin(), lin()
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = torch.empty(B.gated_workspace_size(M, N, K), dtype=torch.uint8, device=dev)
for _ in range(iters):
    B.enqueue_gated(A, gate, up, ind, out, ws)
torch.cuda.synchronize()
print("done gated", M, N, K)
