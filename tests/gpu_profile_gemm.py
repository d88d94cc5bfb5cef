"""ncu target: GEMM stage only (quantised operands prepared once), one shape/config.
   python tests/gpu_profile_gemm.py CFG M N K [iters]"""
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
lib = B.load()
A8 = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
W8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
sa = (torch.rand(M, device=dev) * 0.01 + 1e-3).half()
sb = (torch.rand(N, device=dev) * 0.002 + 1e-4).half()
fpA = torch.randn(M, 128, device=dev).half()
fw = (torch.randn(N, 128, device=dev) * 0.02).half()
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = torch.zeros(lib.mixq_decode_workspace_size(min(M, 1024), N), dtype=torch.uint8, device=dev)
for _ in range(iters):
    B.gemm_dequant(A8, W8, sa, sb, fpA, fw, out, workspace=ws, config=cfg)
torch.cuda.synchronize()
print("done", cfg, M, N, K)
