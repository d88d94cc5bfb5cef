"""ncu target: the M <= 4 weight-only GEMV.   python tests/gpu_profile_gemv.py M N K [iters]"""
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
A = torch.randn(M, K, device=dev).half()
sc = (torch.rand(N, device=dev) * 1e-3 + 1e-4).half()
out = torch.empty(M, N, dtype=torch.float16, device=dev)
qs = [torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev) for _ in range(iters)]
for q in qs:
    B.gemv_w8a16(A, q, sc, out)
torch.cuda.synchronize()
print("done gemv", M, N, K)
