"""cuBLASLt INT8 GEMM (torch._int_mm) as an ncu target: the library ceiling on this box."""
import sys
import torch
M, N, K = (int(x) for x in sys.argv[1:4])
a = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
b = torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda")
for _ in range(3):
    c = torch._int_mm(a, b.t())
torch.cuda.synchronize()
print("done")
