#!/usr/bin/env python
"""bench.py -- W8A8O16 GEMM throughput on the Llama-2-7B linear shapes (BASELINE.json configs[1]).

One *step* = one pass of the hot path (MixQPlugin::enqueue -> quant/extract kernel + tcgen05
GEMM kernel) over one batch of bs=32 x seq=2048 = 65536 synthetic tokens through the five linear
shapes of a Llama-2-7B decoder layer (qkv 12288x4096, o 4096x4096, gate 11008x4096,
up 11008x4096, down 4096x11008).  Metric: W8A8O16 GEMM TFLOPS = 2*M*N*K summed over the five
linears / step time (the reference's "INT8 ops" count; the 128-column FP16 outlier GEMM and the
quantise pass are inside the time but not in the numerator).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

N > 1 (torchrun, one rank per GPU): the same batch, linears sharded tensor-parallel
(qkv/gate/up column-parallel, o/down row-parallel with ONE NCCL all-reduce each) -> strong scaling.
`--impl reference` times the reference's CPU path (the oracle port of MixQ/src + plugin
arithmetic, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (M tokens, [(linear, N, K, parallel)], description)
    "llama2-7b-linears-bs32xseq2048": (65536, [("qkv", 12288, 4096, "column"), ("o", 4096, 4096, "row"),
                                               ("gate", 11008, 4096, "column"), ("up", 11008, 4096, "column"),
                                               ("down", 4096, 11008, "row")]),
    "llama2-7b-linears-decode-bs512": (512, [("qkv", 12288, 4096, "column"), ("o", 4096, 4096, "row"),
                                             ("gate", 11008, 4096, "column"), ("up", 11008, 4096, "column"),
                                             ("down", 4096, 11008, "row")]),
    "llama2-7b-linears-decode-bs32": (32, [("qkv", 12288, 4096, "column"), ("o", 4096, 4096, "row"),
                                           ("gate", 11008, 4096, "column"), ("up", 11008, 4096, "column"),
                                           ("down", 4096, 11008, "row")]),
    "llama2-70b-linears-decode-bs512": (512, [("qkv", 10240, 8192, "column"), ("o", 8192, 8192, "row"),
                                              ("gate", 28672, 8192, "column"), ("up", 28672, 8192, "column"),
                                              ("down", 8192, 28672, "row")]),
    "qwen2-7b-linears-bs32xseq2048": (65536, [("qkv", 4608, 3584, "column"), ("o", 3584, 3584, "row"),
                                              ("gate", 18944, 3584, "column"), ("up", 18944, 3584, "column"),
                                              ("down", 3584, 18944, "row")]),
}
METRIC = "W8A8O16 GEMM TFLOPS (Llama-2-7B linears)"
UNIT = "TFLOP/s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: the poller starts before
    the warm-up (nvidia-smi needs ~100 ms to produce its first line), every line is stamped on
    arrival, and only lines that arrived inside [mark_begin, mark_end] are summarised."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for ln in self.proc.stdout:
                    self.lines.append((time.perf_counter(), ln))
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if self.t0 is not None and not (self.t0 <= ts <= (self.t1 or 1e30) + 0.03):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (qkv shape at
    M=65536), from the committed ncu --set full capture (profiles/r1_traffic.json); None if absent."""
    p = ROOT / "profiles" / "r1_traffic.json"
    try:
        d = json.loads(p.read_text())["gemm"]
        return {"bytes": d["dram_read_bytes"] + d["dram_write_bytes"], "algorithmic_bytes": d["algorithmic_bytes"],
                "shape": d["shape"], "source": "profiles/r1s2_ncu_summary.csv"}
    except (OSError, KeyError, ValueError):
        return None


def shard(N, K, mode, tp):
    if tp == 1:
        return N, K
    return (N // tp, K) if mode == "column" else (N, K // tp)


# ----------------------------------------------------------------------------------------------
def run_reference(args, M, linears):
    """The reference's CPU path: oracle port (oracle/mixq_oracle.c), all host threads, on a
    bounded sample of `sample_tokens` tokens per step through the same five linears."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:            # under torchrun rank 0 alone runs the CPU arm; the others leave without touching the build
        return
    from oracle import oracle as O
    O.build()
    O.set_threads()          # all host cores (torchrun exports OMP_NUM_THREADS=1)
    sample = args.cpu_sample_tokens
    lins, acts = [], {}
    for name, N, K, _ in linears:
        sc = O.load_act_scales({4096: "Llama-2-7b/self_attn.q_proj", 11008: "Llama-2-7b/mlp.down_proj"}.get(K, ""))
        lin = O.synth_linear(N, K, sc, seed=1234)
        lins.append((name, lin))
        if K not in acts:
            acts[K] = O.synth_activations(sample, lin["act_scale"], seed=4321)
    flops = sum(2.0 * sample * N * K for _, N, K, _ in linears)

    def step():
        for name, lin in lins:
            O.forward(acts[lin["W8"].shape[1]], lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"])
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = flops / dt / 1e12
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": args.workload, "tokens_per_step": M, "sample_tokens_per_step": sample,
                       "linears": [[n, N, K] for n, N, K, _ in linears]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                             "sample": f"{sample} of {M} tokens per step through all {len(linears)} linears "
                                       f"(oracle/mixq_oracle.c, OpenMP, {O.num_threads()} threads)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tokens_per_s": sample / dt}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def run_ours(args, M, linears):
    import torch
    import torch.distributed as dist
    from mixq_tensorrt_llm_b200 import binding as B
    from mixq_tensorrt_llm_b200.plugin import MixQLinear

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B.require_device()
    lib = B.load()
    tp = world
    pk = peaks()

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    mods, acts, outs = [], {}, {}
    max_out = 0
    for name, N, K, mode in linears:
        Ns, Ks = shard(N, K, mode, tp)
        mod = MixQLinear(K, N, tp_size=tp, tp_group=dist.group.WORLD if tp > 1 else None, parallel_mode=mode,
                         gather_output=False, device=dev)
        W8 = torch.randint(-127, 128, (Ns, Ks), dtype=torch.int8, device=dev, generator=g)
        ind = torch.randperm(Ks, device=dev, generator=g)[:128].int()
        W8[:, ind.long()] = 0
        sb = (torch.rand(Ns, device=dev, generator=g) * 2e-4 + 1e-4).half()
        fw = (torch.randn(Ns, 128, device=dev, generator=g) * 0.02).half()
        mod.load_packed(W8, sb, fw, ind)
        mods.append((name, mod, Ns, Ks, mode))
        if Ks not in acts:
            a = torch.randn(M, Ks, device=dev, generator=g).half()
            a[:, ind.long()] *= 20.0
            acts[Ks] = a
        max_out = max(max_out, Ns)
    out_buf = torch.empty(M * max_out, dtype=torch.float16, device=dev)
    ws = torch.empty(max(B.workspace_size(M, n, k) for _, _, n, k, _ in mods), dtype=torch.uint8, device=dev)
    flops_step = sum(2.0 * M * N * K for _, N, K, _ in linears)      # whole job, all ranks together
    stream = torch.cuda.current_stream()

    chunks = args.tp_chunks if (tp > 1 and M >= 4096 * args.tp_chunks) else 1
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    # Row-parallel linears: the all-reduce is fused into the GEMM kernel (partial tiles pushed to their owner over
    # NVLink peer memory, fp32 reduce, result written to every rank) -- mixq_enqueue_allreduce.  --tp-reduce nccl keeps
    # the unfused baseline (mixq_enqueue + NCCL all-reduce in overlapped row slabs) for comparison.
    peer = None
    peer_note = None
    if tp > 1 and args.tp_reduce == "fused":
        from mixq_tensorrt_llm_b200.peer import PeerBuffers
        try:
            peer = PeerBuffers(M, max(Ns for _, _, Ns, _, mode in mods if mode == "row"), device=dev)
            ok = 1
        except Exception as e:   # no peer mapping on this box (symmetric memory unavailable): every rank must agree
            peer_note, ok = repr(e)[:160], 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            peer = None
            peer_note = peer_note or "peer buffers unavailable on another rank"
            print(f"[bench] fused all-reduce unavailable ({peer_note}); row-parallel linears use mixq_enqueue + NCCL",
                  file=sys.stderr, flush=True)

    def step():
        for name, mod, Ns, Ks, mode in mods:
            out = out_buf[: M * Ns].view(M, Ns)
            W8 = mod.weight.view(torch.int8).view(Ns, Ks)
            if tp > 1 and mode == "row" and peer is not None:
                B.enqueue_allreduce(acts[Ks], W8, mod.weights_scaling_factor, mod.fp_weight, mod.fp_ind.view(torch.int32),
                                    ws, peer.peer_group(M, Ns))
            elif tp > 1 and mode == "row":
                # The one exchange step of the path.  The token dimension is cut into `chunks` row slabs:
                # the NCCL all-reduce of slab c (on NCCL's stream) overlaps the GEMM of slab c+1.
                works = []
                rows = M // chunks
                # the GEMM is persistent (one CTA per SM): while slabs of THIS linear are in flight keep a few
                # SMs free so NCCL's channels (32 CTAs) can run beside it (measured: 1.52 -> 1.43 ms for o_proj @ tp2)
                lim = nsm - args.comm_sms if (chunks > 1 and args.comm_sms > 0) else 0
                for c in range(chunks):
                    a, o = acts[Ks][c * rows:(c + 1) * rows], out[c * rows:(c + 1) * rows]
                    B.enqueue(a, W8, mod.weights_scaling_factor, mod.fp_weight, mod.fp_ind.view(torch.int32), o, ws, sm_limit=lim)
                    works.append(dist.all_reduce(o, async_op=True))
                for w in works:
                    w.wait()
            else:
                B.enqueue(acts[Ks], W8, mod.weights_scaling_factor, mod.fp_weight, mod.fp_ind.view(torch.int32), out, ws)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nl0 = lib.mixq_launch_count()
    step()
    launches_per_step = lib.mixq_launch_count() - nl0
    # Decode-sized steps are a handful of 10-30 us kernels: replay them from a CUDA graph, as a serving
    # runtime (TensorRT) would, so the step is not bounded by Python/launch latency.
    # With the all-reduce fused into the GEMM kernel a tensor-parallel step holds no NCCL call and is capturable too.
    use_graph = args.graph == "on" or (args.graph == "auto" and M <= 2048 and (world == 1 or peer is not None))
    run_step = step
    if use_graph:
        gs = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()      # the step above ran on another stream: fused all-reduce launches must not overlap
        with torch.cuda.stream(gs):
            step()
            gs.synchronize()
            with torch.cuda.graph(graph, stream=gs):
                step()
        torch.cuda.synchronize()
        run_step = graph.replay
    for _ in range(max(args.warmup, 3)):
        run_step()
    barrier()
    n0 = lib.mixq_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps if use_graph else lib.mixq_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = flops_step / (ms_step * 1e-3) / 1e12

    # ---- per-kernel pass (same steps): CUDA events around each of the two kernels of each linear
    kt = {}
    A8 = ws[: M * max(k for _, _, _, k, _ in mods)]
    evs = []
    for it in range(args.steps):
        for name, mod, Ns, Ks, mode in mods:
            a8 = A8[: M * Ks].view(torch.int8).view(M, Ks)
            sa = torch.empty(M, dtype=torch.float16, device=dev) if it == 0 else kt[name + "_sa"]
            fpA = torch.empty(M, 128, dtype=torch.float16, device=dev) if it == 0 else kt[name + "_fpA"]
            kt[name + "_sa"], kt[name + "_fpA"] = sa, fpA
            out = out_buf[: M * Ns].view(M, Ns)
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            B.quant_extract(acts[Ks], mod.fp_ind.view(torch.int32), a8, sa, fpA)
            b.record()
            B.gemm_dequant(a8, mod.weight.view(torch.int8).view(Ns, Ks), sa, mod.weights_scaling_factor, fpA,
                           mod.fp_weight, out)
            c.record()
            evs.append((name, Ns, Ks, a, b, c))
    torch.cuda.synchronize()
    quant_us, gemm_us = {}, {}
    for name, Ns, Ks, a, b, c in evs:
        quant_us.setdefault(name, []).append(a.elapsed_time(b) * 1e3)
        gemm_us.setdefault(name, []).append(b.elapsed_time(c) * 1e3)
    per_linear = {}
    tot_gemm = tot_quant = 0.0
    for name, mod, Ns, Ks, mode in mods:
        gu, qu = float(np.mean(gemm_us[name])), float(np.mean(quant_us[name]))
        tot_gemm += gu
        tot_quant += qu
        per_linear[name] = {"N": Ns, "K": Ks, "gemm_us": round(gu, 1), "gemm_tflops": round(2.0 * M * Ns * Ks / gu / 1e6, 1),
                            "quant_us": round(qu, 1), "quant_gbs": round((3.0 * M * Ks + 258.0 * M) / qu / 1e3, 1)}
    dom = max(per_linear.items(), key=lambda kv: kv[1]["gemm_us"])
    int8_peak = 2.0 * pk["bf16_sustained"]
    gemm_flops = sum(2.0 * M * Ns * Ks for _, _, Ns, Ks, _ in mods)
    achieved = gemm_flops / tot_gemm / 1e6
    roofline = {"bound": "tensor", "kernel": "mixq_gemm_dequant_streamk_kernel, whole-tile schedule (tcgen05 kind::i8 + kind::f16, cta_group::2, TMA-store epilogue)",
                "achieved": round(achieved, 1), "peak": round(int8_peak, 1), "unit": "TFLOP/s",
                "frac": round(achieved / int8_peak, 4),
                "peak_source": f"2 x bf16_tflops_sustained of MEASURED_PEAKS.json ({pk['src']}); INT8 dense = 2 x BF16 dense",
                "frac_of_spec_4500": round(achieved / 4500.0, 4),
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (ncu --set full capture)
                "traffic": (traffic_from_profile() or {}).get("bytes"), "traffic_detail": traffic_from_profile(),
                "share_of_step": round(tot_gemm / (tot_gemm + tot_quant), 4),
                "quant_kernel": {"bound": "hbm", "achieved": round(sum((3.0 * M * k + 258.0 * M) for _, _, _, k, _ in mods) / tot_quant / 1e3, 1),
                                 "peak": pk["hbm"], "unit": "GB/s"},
                "per_linear": per_linear, "dominant": dom[0]}

    # ---- measured INT8 library peak on this box, same run (cuBLASLt via torch._int_mm, 8192^3, best of 10)
    try:
        a = torch.randint(-128, 128, (8192, 8192), dtype=torch.int8, device=dev)
        b = torch.randint(-128, 128, (8192, 8192), dtype=torch.int8, device=dev).t()
        best = 1e9
        for i in range(13):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(); torch._int_mm(a, b); s1.record(); torch.cuda.synchronize()
            if i >= 3:
                best = min(best, s0.elapsed_time(s1))
        roofline["cublaslt_int8_8192_tflops"] = round(2.0 * 8192 ** 3 / best / 1e9, 1)
        roofline["frac_of_cublaslt_int8"] = round(achieved / roofline["cublaslt_int8_8192_tflops"], 4)
        del a, b
    except Exception as e:  # library may lack an int8 path; the number is informational
        roofline["cublaslt_int8_8192_tflops"] = f"unavailable: {e!r}"[:80]

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (pinned), H2D + enqueue + D2H per linear
    e2e = None
    if not args.no_e2e:
        try:
            maxK = max(k for _, _, _, k, _ in mods)
            hA = {k: torch.empty(M, k, dtype=torch.float16).pin_memory() for k in acts}
            for k in acts:
                hA[k].copy_(acts[k])
            hO = torch.empty(M * max_out, dtype=torch.float16).pin_memory()
            scratch = torch.empty(max(lib.mixq_host_scratch_size(M, n, k) for _, _, n, k, _ in mods), dtype=torch.uint8, device=dev)
            h2d = sum(M * k * 2 for _, _, _, k, _ in mods)
            d2h = sum(M * n * 2 for _, _, n, _, _ in mods)

            dA_buf = torch.empty(M * maxK, dtype=torch.float16, device=dev) if peer is not None else None
            s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
            ev_down = None

            def e2e_step():
                nonlocal ev_down
                for name, mod, Ns, Ks, mode in mods:
                    if peer is not None and mode == "row":
                        # row-parallel shard: H2D (copy stream), fused GEMM + all-reduce, D2H of the reduced result (copy
                        # stream; overlaps the next linear).  The peer Out buffer is reused: wait for its last D2H.
                        dA = dA_buf[: M * Ks].view(M, Ks)
                        s_up.wait_stream(stream)
                        with torch.cuda.stream(s_up):
                            dA.copy_(hA[Ks], non_blocking=True)
                        stream.wait_stream(s_up)
                        if ev_down is not None:
                            stream.wait_event(ev_down)
                        B.enqueue_allreduce(dA, mod.weight.view(torch.int8).view(Ns, Ks), mod.weights_scaling_factor, mod.fp_weight,
                                            mod.fp_ind.view(torch.int32), ws, peer.peer_group(M, Ns))
                        s_down.wait_stream(stream)
                        with torch.cuda.stream(s_down):
                            hO[: M * Ns].view(M, Ns).copy_(peer.out(M, Ns), non_blocking=True)
                            ev_down = torch.cuda.Event()
                            ev_down.record(s_down)
                        continue
                    t_ = B.make_tensors(None, mod.weight, mod.weights_scaling_factor, mod.fp_weight, mod.fp_ind, None)
                    B.check(lib.mixq_linear_host(ctypes.byref(t_), hA[Ks].data_ptr(), hO.data_ptr(), M, Ns, Ks,
                                                 scratch.data_ptr(), scratch.numel(), 0, stream.cuda_stream), "mixq_linear_host")
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.e2e_steps
            tt = torch.tensor([dt], device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e = {"value": flops_step / float(tt.item()) / 1e12, "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": float(tt.item()) * 1e3, "steps": args.e2e_steps,
                   "path": "mixq_linear_host (C ABI): pinned host A -> H2D -> mixq_enqueue -> D2H Out, per linear"}
            del hA, hO, scratch
        except Exception as e:
            e2e = {"value": None, "unit": UNIT, "error": repr(e)[:200]}

    # ---- the other batch size of the metric (decode, bs = 512 tokens per step) in the same run, N = 1: the same five
    # linears through mixq_enqueue, the step replayed from a CUDA graph; weights (202 MB) exceed L2, activations do not
    decode = None
    if world == 1 and M > 512 and not args.no_decode:
        try:
            Md = 512
            dacts = {k: acts[k][:Md].contiguous() for k in acts}
            dout = torch.empty(Md * max_out, dtype=torch.float16, device=dev)

            def dstep():
                for name, mod, Ns, Ks, mode in mods:
                    B.enqueue(dacts[Ks], mod.weight.view(torch.int8).view(Ns, Ks), mod.weights_scaling_factor, mod.fp_weight,
                              mod.fp_ind.view(torch.int32), dout[: Md * Ns].view(Md, Ns), ws)
            torch.cuda.synchronize()
            gs2 = torch.cuda.Stream()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.stream(gs2):
                dstep()
                gs2.synchronize()
                with torch.cuda.graph(g2, stream=gs2):
                    dstep()
            torch.cuda.synchronize()
            for _ in range(5):
                g2.replay()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            d0.record()
            nrep = 200
            for _ in range(nrep):
                g2.replay()
            d1.record()
            torch.cuda.synchronize()
            dms = d0.elapsed_time(d1) / nrep
            dfl = sum(2.0 * Md * N * K for _, N, K, _ in linears)
            wbytes = sum(N * K for _, N, K, _ in linears)
            decode = {"workload": args.workload.replace("bs32xseq2048", "decode-bs512"), "tokens_per_step": Md,
                      "ms_per_step": dms, "value": dfl / (dms * 1e-3) / 1e12, "unit": UNIT, "tokens_per_s": Md / (dms * 1e-3),
                      "launch": "cuda-graph replay of the step", "steps": nrep,
                      "frac_of_int8_peak": round(dfl / (dms * 1e-3) / 1e12 / (2.0 * pk["bf16_sustained"]), 4),
                      "weight_stream_gbs": round(wbytes / (dms * 1e-3) / 1e9, 1)}
            del dout
        except Exception as e:
            decode = {"error": repr(e)[:200]}

    # ---- same-box GPU baseline: the reference's own kernels recompiled for sm_100a (oracle/_ref), N=1 only
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_ref_gpu:
        try:
            sys.path.insert(0, str(ROOT / "tests"))
            import refgpu
            if refgpu.available():
                rws = torch.empty(max(refgpu.load().ref_workspace_size(M, n, k) for _, _, n, k, _ in mods), dtype=torch.uint8, device=dev)
                def ref_step():
                    for name, mod, Ns, Ks, mode in mods:
                        refgpu.enqueue(acts[Ks], mod.weight.view(torch.int8).view(Ns, Ks), mod.weights_scaling_factor,
                                       mod.fp_weight, mod.fp_ind.view(torch.int32), out_buf[: M * Ns].view(M, Ns), rws)
                ref_step(); torch.cuda.synchronize()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record()
                for _ in range(2):
                    ref_step()
                r1.record(); torch.cuda.synchronize()
                rms = r0.elapsed_time(r1) / 2
                ref_gpu = {"what": "reference kernel/i8gemm.cu + cuBLAS fp16, 4 launches per linear, recompiled for sm_100a",
                           "ms_per_step": rms, "value": flops_step / (rms * 1e-3) / 1e12, "unit": UNIT,
                           "speedup_ours": rms / ms_step}
                del rws
        except Exception as e:
            ref_gpu = {"error": repr(e)[:200]}

    # ---- CPU baseline (rank 0, N=1): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        O.build()
        O.set_threads()
        sample = min(M, max(args.cpu_sample_tokens, args.cpu_baseline_tokens))   # ~10 s of host work on 16 cores
        t_cpu = 0.0
        for name, N, K, _ in linears:
            lin = O.synth_linear(N, K, None, seed=1)
            A = O.synth_activations(sample, lin["act_scale"])
            t0 = time.perf_counter()
            O.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"])
            t_cpu += time.perf_counter() - t0
        cpu = {"value": sum(2.0 * sample * N * K for _, N, K, _ in linears) / t_cpu / 1e12, "unit": UNIT,
               "cores": O.num_threads(), "kind": "port",
               "sample": f"{sample} of {M} tokens, one pass through all {len(linears)} linears (oracle/mixq_oracle.c, OpenMP)",
               "seconds": round(t_cpu, 2)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
                "config": {"workload": args.workload, "tokens_per_step": M,
                           "linears": [[n, N, K, m] for n, N, K, m in linears],
                           "comm_sms": args.comm_sms if (tp > 1 and chunks > 1 and peer is None) else 0,
                           **({"fused_allreduce_unavailable": peer_note} if peer_note else {}),
                           "launch": "cuda-graph replay of the step" if use_graph else "direct launches",
                           "parallelism": ("single" if tp == 1 else
                                           f"tp{tp} (column: no collective; row: all-reduce fused into the GEMM kernel over NVLink peer memory)"
                                           if peer is not None else
                                           f"tp{tp} (column: no collective; row: one NCCL all-reduce in {chunks} overlapped row slabs)"),
                           "l2": "inputs larger than L2 (activations %.0f MB per linear), no flush" % (M * 4096 * 2 / 1e6)
                                 if M >= 16384 else "weights rotate through >126 MB per step; activations L2-resident"},
                "tokens_per_s": M / (ms_step * 1e-3), "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "decode_bs512": decode, "ref_gpu": ref_gpu,
                "int8_peak_note": "no INT8 figure in MEASURED_PEAKS.json; peak = 2 x measured bf16 (dense INT8 = 2 x dense BF16 on sm_100)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="llama2-7b-linears-bs32xseq2048", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-tokens", type=int, default=512, help="tokens per step of the --impl reference arm")
    ap.add_argument("--cpu-baseline-tokens", type=int, default=4096, help="token sample of the cpu_baseline leg of our arm")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--tp-chunks", type=int, default=4, help="row slabs per row-parallel linear (comm/compute overlap)")
    ap.add_argument("--comm-sms", type=int, default=40, help="SMs left free for NCCL when slabs overlap (tensor parallel only)")
    ap.add_argument("--tp-reduce", default="fused", choices=["fused", "nccl"],
                    help="row-parallel linears: all-reduce fused into the GEMM kernel (default) or NCCL after it")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (auto: decode-sized M on one GPU)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the bs=512 decode leg of the default run")
    args = ap.parse_args()
    M, linears = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, M, linears)
    else:
        run_ours(args, M, linears)


if __name__ == "__main__":
    main()
