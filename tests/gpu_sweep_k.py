"""Debug aid: is the time of a back-to-back launched GEMM quantised (cluster launches) or smooth?
Sweeps K for the given configs and prints us/launch for (a) plain back-to-back launches and (b) a CUDA-graph of 20 launches.
Also prints torch._int_mm (cuBLASLt INT8) at the decode shapes as a library yardstick.
    python tests/gpu_sweep_k.py "1,5,9" M N
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402

cfgs = [int(x) for x in sys.argv[1].split(",")]
M, N = int(sys.argv[2]), int(sys.argv[3])
dev = "cuda"
lib = B.load()
Kmax = 5120
A8f = torch.randint(-127, 128, (M, Kmax), dtype=torch.int8, device=dev)
W8f = torch.randint(-127, 128, (N, Kmax), dtype=torch.int8, device=dev)
sa = (torch.rand(M, device=dev) * 0.01 + 1e-3).half()
sb = (torch.rand(N, device=dev) * 0.002 + 1e-4).half()
fpA = torch.randn(M, 128, device=dev).half()
fw = (torch.randn(N, 128, device=dev) * 0.02).half()
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = torch.zeros(lib.mixq_decode_workspace_size(min(M, 1024), N), dtype=torch.uint8, device=dev)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


print(f"M={M} N={N}: us per launch, direct back-to-back | inside a 20-launch CUDA graph")
for K in range(3072, Kmax + 1, 256):
    A8 = A8f[:, :K].contiguous()
    W8 = W8f[:, :K].contiguous()
    row = [f"K={K:5d}"]
    for c in cfgs:
        f = lambda c=c: B.gemm_dequant(A8, W8, sa, sb, fpA, fw, out, workspace=ws, config=c)  # noqa: E731
        d = timed(f, 100)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            f()
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(20):
                    f()
        torch.cuda.synchronize()
        gr = timed(g.replay, 10) / 20
        row.append(f"cfg{c}: {d:6.2f} | {gr:6.2f}")
    print("  ".join(row), flush=True)
for (m, n, k) in [(512, 12288, 4096), (512, 4096, 4096), (512, 11008, 4096), (512, 4096, 11008), (512, 22016, 4096)]:
    a = torch.randint(-127, 128, (m, k), dtype=torch.int8, device=dev)
    w = torch.randint(-127, 128, (n, k), dtype=torch.int8, device=dev)
    us = timed(lambda: torch._int_mm(a, w.t()), 100)
    print(f"torch._int_mm {m}x{n}x{k}: {us:6.2f} us  {2.0 * m * n * k / us / 1e6:7.1f} TOPS")
