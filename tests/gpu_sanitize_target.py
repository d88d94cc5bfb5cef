"""compute-sanitizer target: one small call of every kernel family through the C ABI.
   compute-sanitizer --tool memcheck --error-exitcode 7 python tests/gpu_sanitize_target.py"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402

dev = "cuda"
B.require_device()
g = torch.Generator(device="cpu").manual_seed(7)


def lin(N, K):
    W8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g).to(dev)
    sb = (torch.rand(N, generator=g) * 2e-3 + 1e-4).half().to(dev)
    fw = (torch.randn(N, 128, generator=g) * 0.02).half().to(dev)
    ind = torch.randperm(K, generator=g)[:128].int().to(dev)
    return W8, sb, fw, ind


for (M, N, K) in [(64, 512, 4096), (300, 264, 400), (512, 1024, 1024), (1500, 520, 656), (32, 384, 4096)]:
    W8, sb, fw, ind = lin(N, K)
    A = torch.randn(M, K, generator=g).half().to(dev)
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=dev)
    B.enqueue(A, W8, sb, fw, ind, out, ws)                       # quantise kernel + the GEMM kernel `auto` picks
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    print("enqueue", M, N, K, "ok", flush=True)

# gated call (fat tile, two tensor maps over the weights)
M, N, K = 512, 704, 1024
W8, sb, fw, ind = lin(N, K)
W8u, sbu, fwu, _ = lin(N, K)
A = torch.randn(M, K, generator=g).half().to(dev)
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = torch.empty(B.gated_workspace_size(M, N, K), dtype=torch.uint8, device=dev)
B.enqueue_gated(A, (W8, sb, fw), (W8u, sbu, fwu), ind, out, ws)
torch.cuda.synchronize()
print("gated ok", flush=True)

# M <= 4 weight-only GEMV
for Mg in (1, 2, 3, 4):
    N, K = 512, 1024
    q = torch.randint(-128, 128, (K, N), dtype=torch.int8, generator=g).to(dev)
    sc = (torch.rand(N, generator=g) * 1e-3 + 1e-4).half().to(dev)
    A = torch.randn(Mg, K, generator=g).half().to(dev)
    out = torch.empty(Mg, N, dtype=torch.float16, device=dev)
    B.gemv_w8a16(A, q, sc, out)
torch.cuda.synchronize()
print("gemv ok", flush=True)

# queued host-buffer calls
M, N, K = 300, 264, 400
W8, sb, fw, ind = lin(N, K)
tab = B.make_tensors(None, W8, sb, fw, ind, None)
hA = torch.randn(M, K, generator=g).half().pin_memory()
outs = [torch.empty(M, N, dtype=torch.float16).pin_memory() for _ in range(5)]
scratch = torch.empty(3 * B.linears_host_scratch_size(M, [N], K) + 512, dtype=torch.uint8, device=dev)
for o in outs:
    B.linears_host([tab], hA, [o], scratch, flags=B.FLAG_HOST_ASYNC)
B.host_drain()
assert all(torch.equal(o.view(torch.int16), outs[0].view(torch.int16)) for o in outs)
print("host async ok", flush=True)
