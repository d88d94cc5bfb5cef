"""2-GPU diagnostic: does the NCCL all-reduce of row slab c overlap the GEMM of slab c+1?
   torchrun --nproc-per-node 2 tests/gpu_tp_overlap.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = B.load()
M, N, K = 65536, 4096, 2048            # o_proj shard at tp2
A = torch.randn(M, K, device=dev).half()
W8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
sb = (torch.rand(N, device=dev) * 2e-4 + 1e-4).half()
fw = (torch.randn(N, 128, device=dev) * 0.02).half()
ind = torch.randperm(K, device=dev)[:128].int()
out = torch.empty(M, N, dtype=torch.float16, device=dev)
ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=dev)
nsm = torch.cuda.get_device_properties(dev).multi_processor_count


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def compute(chunks):
    rows = M // chunks
    for c in range(chunks):
        B.enqueue(A[c * rows:(c + 1) * rows], W8, sb, fw, ind, out[c * rows:(c + 1) * rows], ws, sm_limit=SM_LIMIT)


def comm(chunks):
    rows = M // chunks
    ws_ = [dist.all_reduce(out[c * rows:(c + 1) * rows], async_op=True) for c in range(chunks)]
    for w in ws_:
        w.wait()


def both(chunks):
    rows = M // chunks
    works = []
    for c in range(chunks):
        B.enqueue(A[c * rows:(c + 1) * rows], W8, sb, fw, ind, out[c * rows:(c + 1) * rows], ws, sm_limit=SM_LIMIT)
        works.append(dist.all_reduce(out[c * rows:(c + 1) * rows], async_op=True))
    for w in works:
        w.wait()


side = torch.cuda.Stream()
SM_LIMIT = 0


def both_side_stream(chunks):
    """explicit side stream + sync all_reduce issued from it"""
    rows = M // chunks
    main = torch.cuda.current_stream()
    for c in range(chunks):
        B.enqueue(A[c * rows:(c + 1) * rows], W8, sb, fw, ind, out[c * rows:(c + 1) * rows], ws, sm_limit=SM_LIMIT)
        ev = torch.cuda.Event(); ev.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            dist.all_reduce(out[c * rows:(c + 1) * rows])
    main.wait_stream(side)


for lim in (0, 16, 40):
    SM_LIMIT = nsm - lim if lim else 0
    r = {"compute1": timeit(lambda: compute(1)), "compute4": timeit(lambda: compute(4)), "comm1": timeit(lambda: comm(1)),
         "comm4": timeit(lambda: comm(4)), "both1": timeit(lambda: both(1)), "both4": timeit(lambda: both(4)),
         "both8": timeit(lambda: both(8)), "side4": timeit(lambda: both_side_stream(4))}
    if rank == 0:
        print(f"reserved SMs {lim}: " + "  ".join(f"{k} {v:.3f}" for k, v in r.items()), flush=True)
dist.destroy_process_group()
