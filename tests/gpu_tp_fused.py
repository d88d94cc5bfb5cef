"""Multi-GPU check + timing of the fused row-parallel GEMM + all-reduce (mixq_enqueue_allreduce) against the
unfused path (mixq_enqueue, then one NCCL all-reduce).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_tp_fused.py [MxNxK ...]          (K = the UNSHARDED input width)

Parity: the fused result must equal fp16(sum_r fp32(partial_r)) (rank order) bit for bit on every rank, where the
partials are what mixq_enqueue produces on each rank (gathered with NCCL for the comparison only).
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402
from mixq_tensorrt_llm_b200.peer import PeerBuffers  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = B.load()
B.require_device()
shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[1:]] or [(512, 8192, 8192), (512, 4096, 4096), (8192, 4096, 4096)]
maxM, maxN = max(s[0] for s in shapes), max(s[1] for s in shapes)
pb = PeerBuffers(maxM, maxN, device=dev)                       # broadcast through the NVSwitch multicast mapping when there is one
pb_uc = PeerBuffers(maxM, maxN, device=dev, multicast=False) if pb.multicast_base else None   # per-rank TMA stores, for A/B
results = []


def timeit(fn, iters):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3   # us


for (M, N, K) in shapes:
    Kr = K // world
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    A = torch.randn(M, Kr, device=dev, generator=g).half()
    W8 = torch.randint(-127, 128, (N, Kr), dtype=torch.int8, device=dev, generator=g)
    sb = (torch.rand(N, device=dev, generator=g) * 2e-4 + 1e-4).half()
    fw = (torch.randn(N, 128, device=dev, generator=g) * 0.02).half()
    ind = torch.randperm(Kr, device=dev, generator=g)[:128].int()
    A[:, ind.long()] *= 20.0
    ws = torch.empty(B.workspace_size(M, N, Kr), dtype=torch.uint8, device=dev)
    part = torch.empty(M, N, dtype=torch.float16, device=dev)
    B.enqueue(A, W8, sb, fw, ind, part, ws)
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    ref = parts[0].float()
    for r in range(1, world):
        ref = ref + parts[r].float()
    ref = ref.half()
    grp = pb.peer_group(M, N)
    if pb.multicast_base:          # A/B: force the multicast broadcast here regardless of PeerBuffers' size policy
        grp.out_multicast = pb.multicast_base
    out = pb.out(M, N)
    out.fill_(float("nan"))
    torch.cuda.synchronize()
    dist.barrier()
    B.enqueue_allreduce(A, W8, sb, fw, ind, ws, grp, config=9)          # the one-kernel path
    torch.cuda.synchronize()
    bad = int((out.view(torch.int16) != ref.view(torch.int16)).sum())
    # default path: decode-sized results on 2 ranks go GEMM -> pull-reduce kernel (bit-identical arithmetic); elsewhere it IS the one above
    out.fill_(float("nan"))
    torch.cuda.synchronize()
    dist.barrier()
    for _ in range(3):                                                   # consecutive calls: epochs advance, partial buffer reused
        B.enqueue_allreduce(A, W8, sb, fw, ind, ws, grp)
    torch.cuda.synchronize()
    bad_default = int((out.view(torch.int16) != ref.view(torch.int16)).sum())
    nccl_out = part.clone()
    dist.all_reduce(nccl_out)
    nccl_diff = float((nccl_out.float() - ref.float()).abs().max())
    t_bad = torch.tensor([bad], device=dev)
    dist.all_reduce(t_bad)

    def unfused():
        B.enqueue(A, W8, sb, fw, ind, part, ws)
        dist.all_reduce(part)

    def fused():
        B.enqueue_allreduce(A, W8, sb, fw, ind, ws, grp, config=9)

    def default_path():
        B.enqueue_allreduce(A, W8, sb, fw, ind, ws, grp)

    grp_uc = pb_uc.peer_group(M, N) if pb_uc is not None else None

    def fused_unicast():
        B.enqueue_allreduce(A, W8, sb, fw, ind, ws, grp_uc)

    def gemm_only():
        B.enqueue(A, W8, sb, fw, ind, part, ws)

    def nccl_only():
        dist.all_reduce(part)

    iters = 50 if M <= 2048 else 10
    t_bad2 = torch.tensor([bad_default], device=dev)
    dist.all_reduce(t_bad2)
    res = dict(shape=[M, N, K], world=world, mismatches=int(t_bad.item()), default_path_mismatches=int(t_bad2.item()),
               default_path_us=timeit(default_path, 50 if M <= 2048 else 10), nccl_vs_fp32sum_maxabs=nccl_diff,
               unfused_us=timeit(unfused, iters), fused_us=timeit(fused, iters), gemm_only_us=timeit(gemm_only, iters),
               nccl_only_us=timeit(nccl_only, iters))
    # per-CTA timeline of one fused launch (debug stamps of the kernel), rank 0
    import numpy as np
    trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    dist.barrier()
    B.check(lib.mixq_debug_set_trace(trace.data_ptr()), "trace")
    fused()
    torch.cuda.synchronize()
    lib.mixq_debug_set_trace(None)
    tr = trace.cpu().numpy().reshape(148, 16).astype(np.float64)
    t0 = tr[:, 0][tr[:, 0] > 0].min()
    rel = np.where(tr > 0, (tr - t0) / 1e3, np.nan)
    tl = {}
    for nm, i in (("mmaDone", 4), ("epiLoopExit", 10), ("storesLanded", 11), ("phaseA_signalled", 12), ("firstUnitReady", 13),
                  ("phaseB_issued", 14), ("exit", 15)):
        col = rel[:, i]
        if (~np.isnan(col)).any():
            tl[nm] = [round(float(np.nanmin(col)), 1), round(float(np.nanmedian(col)), 1), round(float(np.nanmax(col)), 1)]
    res["timeline_us_min_med_max"] = tl
    res["multicast_broadcast"] = bool(pb.multicast_base)
    if pb_uc is not None:
        out_uc = pb_uc.out(M, N)
        out_uc.fill_(float("nan"))
        torch.cuda.synchronize()
        dist.barrier()
        fused_unicast()
        torch.cuda.synchronize()
        res["unicast_mismatches"] = int((out_uc.view(torch.int16) != ref.view(torch.int16)).sum())
        res["fused_unicast_us"] = timeit(fused_unicast, iters)
    res["speedup"] = res["unfused_us"] / res["fused_us"]
    res["allreduce_bytes"] = M * N * 2
    if rank == 0:
        print(json.dumps(res), flush=True)
    results.append(res)
# ---- the module path: MixQLinear(parallel_mode="row").attach_peer_buffers(...) == plugin + NCCL all-reduce up to the
# summation order (fp32 rank-order sum vs NCCL's fp16 ring), and a CUDA-graph replay of the fused call
from mixq_tensorrt_llm_b200.plugin import MixQLinear  # noqa: E402
M, N, K = shapes[0]
Kr = K // world
g = torch.Generator(device=dev).manual_seed(500 + rank)
mod = MixQLinear(K, N, tp_size=world, tp_group=dist.group.WORLD, parallel_mode="row", device=dev)
W8 = torch.randint(-127, 128, (N, Kr), dtype=torch.int8, device=dev, generator=g)
ind = torch.randperm(Kr, device=dev, generator=g)[:128].int()
mod.load_packed(W8, (torch.rand(N, device=dev, generator=g) * 2e-4 + 1e-4).half(),
                (torch.randn(N, 128, device=dev, generator=g) * 0.02).half(), ind)
A = torch.randn(M, Kr, device=dev, generator=g).half()
y_nccl = mod(A).clone()
mod.attach_peer_buffers(pb)
y_fused = mod(A).clone()
torch.cuda.synchronize()
rel = float((y_fused.float() - y_nccl.float()).norm() / y_nccl.float().norm())
gs = torch.cuda.Stream()
graph = torch.cuda.CUDAGraph()
dist.barrier()
torch.cuda.synchronize()
with torch.cuda.stream(gs):
    mod(A)
    gs.synchronize()
    with torch.cuda.graph(graph, stream=gs):
        y_graph = mod(A)
torch.cuda.synchronize()
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
graph_ok = bool(torch.equal(y_graph.view(torch.int16), y_fused.view(torch.int16)))
module_ok = rel < 2e-3 and graph_ok
t_ok = torch.tensor([1 if module_ok else 0], device=dev)
dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"module_path": {"rel_diff_fused_vs_nccl": rel, "graph_replay_bit_equal": graph_ok}}), flush=True)
dist.barrier()
if rank == 0:
    ok = all(r["mismatches"] == 0 and r["default_path_mismatches"] == 0 and r.get("unicast_mismatches", 0) == 0 for r in results) and bool(t_ok.item())
    print("PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
