#!/usr/bin/env python
"""bench.py -- W8A8O16 GEMM TFLOPS and tokens/s of the Llama-2-7B linears at batch 512 (BASELINE.json's metric).

One *step* = one decode step of the model's MixQ linears: for each of the 32 decoder layers the five linear shapes
(qkv 12288x4096, o 4096x4096, gate 11008x4096, up 11008x4096, down 4096x11008) are run over the same batch of 512
synthetic tokens through the reference-facing call `mixq_enqueue` (= MixQPlugin::enqueue: per-token INT8 quantise +
outlier gather kernel, then the tcgen05 INT8 GEMM with the fused fp16 outlier GEMM and dequant epilogue); the gate and up
projections, which read the same activations, go through ONE `mixq_enqueue_gated` call that also applies the SiLU and
the gate*up multiply of the reference's fused MLP (--mlp split: two plain plugin calls instead).  Every layer
has its own weights (6.5 GB in total), so every weight byte of a step comes from HBM.  The step is replayed from a
CUDA graph, as a serving runtime would.
    value        = sum over the 160 linears of 2*M*N*K / step time            (W8A8O16 GEMM TFLOP/s)
    tokens_per_s = M / step time                                              (8d (i): linears-only, all 32 layers;
                                                                               reference MixQ/src/benchflops.py:97-134,313)
Weights and activations follow SURVEY.md 8d: W ~ N(0, 0.02^2) fp16 packed exactly as the reference's
pack_linear_weights does (model_config_utils.py:429-466), activations randn * act_scale / 3 with the layer-0 activation
maxima of the reference's act_scales files, outlier columns = the 128 largest of them.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

N > 1 (torchrun, one rank per GPU): the same batch, linears sharded tensor-parallel (qkv/gate/up column-parallel with no
collective, o/down row-parallel with the all-reduce fused into the GEMM kernel over NVLink peer memory) -> strong scaling.
`--impl reference` times the reference's CPU path (oracle port of the same arithmetic, all host threads) on a bounded
sample of the same workload.  Other workloads (--workload): the prefill shapes of configs[1]/[3] (M = 65536, one layer per
step), Llama-2-70B bs 512 (configs[4]) and bs 32.
"""
from __future__ import annotations

import argparse
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

L7 = [("qkv", 12288, 4096, "column", "self_attn.q_proj"), ("o", 4096, 4096, "row", "self_attn.o_proj"),
      ("gate", 11008, 4096, "column", "mlp.gate_proj"), ("up", 11008, 4096, "column", "mlp.up_proj"),
      ("down", 4096, 11008, "row", "mlp.down_proj")]
L70 = [("qkv", 10240, 8192, "column", "self_attn.q_proj"), ("o", 8192, 8192, "row", "self_attn.o_proj"),
       ("gate", 28672, 8192, "column", "mlp.gate_proj"), ("up", 28672, 8192, "column", "mlp.up_proj"),
       ("down", 8192, 28672, "row", "mlp.down_proj")]
LQ = [("qkv", 4608, 3584, "column", "self_attn.q_proj"), ("o", 3584, 3584, "row", "self_attn.o_proj"),
      ("gate", 18944, 3584, "column", "mlp.gate_proj"), ("up", 18944, 3584, "column", "mlp.up_proj"),
      ("down", 3584, 18944, "row", "mlp.down_proj")]
WORKLOADS = {
    # name: tokens per step, decoder layers per step (each with its own weights), model layers, act_scales model, linears
    "llama2-7b-linears-decode-bs512": dict(M=512, layers=32, model_layers=32, scales="Llama-2-7b", linears=L7),
    "llama2-7b-linears-decode-bs32": dict(M=32, layers=32, model_layers=32, scales="Llama-2-7b", linears=L7),
    "llama2-7b-linears-bs32xseq2048": dict(M=65536, layers=1, model_layers=32, scales="Llama-2-7b", linears=L7),
    "llama2-70b-linears-decode-bs512": dict(M=512, layers=8, model_layers=80, scales="Llama-2-70b", linears=L70),
    "llama2-70b-linears-decode-bs32": dict(M=32, layers=8, model_layers=80, scales="Llama-2-70b", linears=L70),
    "qwen2-7b-linears-bs32xseq2048": dict(M=65536, layers=1, model_layers=28, scales="qwen2-7b-instruct", linears=LQ),
}
DEFAULT_WORKLOAD = "llama2-7b-linears-decode-bs512"
METRIC = "W8A8O16 GEMM TFLOPS (Llama-2-7B linears, bs=512)"
UNIT = "TFLOP/s"
SPEC_INT8_TOPS = 4500.0


def metric_name(workload):
    return METRIC if workload == DEFAULT_WORKLOAD else f"W8A8O16 GEMM TFLOPS ({workload})"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback of B200_PROFILING.md")


def act_scales(model, key, K):
    """Layer-0 per-channel activation maxima of the reference's act_scales/*.pt (tests/golden/act_scales_l0.npz);
    fallback: ~N(0,1) channels with 128 random ones x20."""
    p = ROOT / "tests" / "golden" / "act_scales_l0.npz"
    if p.exists():
        z = np.load(p)
        name = f"{model}/{key}"
        if name in z.files and z[name].shape[0] == K:
            return z[name].astype(np.float32), "reference act_scales layer 0"
    rng = np.random.default_rng(1234 + K)
    s = np.abs(rng.standard_normal(K)).astype(np.float32) * 0.5 + 0.05
    s[rng.choice(K, size=128, replace=False)] *= 20.0
    return s, "synthetic"


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region.  NVML is polled from a thread (about 1 kHz),
    so even a 50 ms region holds dozens of samples; nvidia-smi -lms 20 is the fallback."""
    R = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.samples, self.t0, self.t1, self.stop_flag, self.proc, self.how = index, [], None, None, False, None, None
        self.sm_max = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self.stop_flag:
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
                        try:
                            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.perf_counter(), float(sm), pw, int(rs)))
                    except Exception:
                        pass
                    time.sleep(0.0005)
            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            self.how = "nvml"
            return
        except Exception:
            self.how = None
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for ln in self.proc.stdout:
                    f = [x.strip() for x in ln.split(",")]
                    try:
                        rs = sum(bit for bit, v in zip((0x8, 0x40, 0x20, 0x4), f[3:7]) if v.lower().startswith("active"))
                        self.samples.append((time.perf_counter(), float(f[0]), float(f[2]), rs))
                        self.sm_max = float(f[1])
                    except (ValueError, IndexError):
                        pass
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
            self.how = "nvidia-smi"
        except OSError:
            self.how = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.t.join(timeout=2)
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= (self.t1 or 1e30) + 0.002]
        sm = [s[1] for s in inside]
        bits = 0
        for s in inside:
            bits |= s[3]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": self.sm_max,
                "power_w_max": max((s[2] for s in inside), default=None), "samples": len(sm), "source": self.how,
                "region_ms": round(((self.t1 or 0) - (self.t0 or 0)) * 1e3, 2),
                "reasons": sorted(n for n, b in self.R.items() if bits & b)}


def linear_bytes(M, N, K):
    """compulsory HBM traffic of one linear call, fused ideal"""
    return 2.0 * M * K + N * K + 256.0 * N + 2.0 * N + 512 + 2.0 * M * N


def bench_config(args, wl):
    """identical for both arms (the driver compares them); run details of our arm live under "run_details" """
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    return {"workload": args.workload, "tokens_per_step": wl["M"], "layers_per_step": wl["layers"],
            "linears": [[n, N, K, m] for n, N, K, m, _ in wl["linears"]],
            "parallelism": "single" if world == 1 else f"tp{world}"}


# ----------------------------------------------------------------------------------------------
def run_reference(args, wl):
    """The reference's CPU path: oracle port (oracle/mixq_oracle.c: the same arithmetic the GPU path is pinned to), all
    host threads.  One step = ONE decoder layer's five linears over the batch (a bounded sample of the workload's
    `layers` layers), or a token sample of it for the prefill workloads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:            # under torchrun rank 0 alone runs the CPU arm; the others leave without touching the build
        return
    from oracle import oracle as O
    O.build()
    O.set_threads()          # all host cores (torchrun exports OMP_NUM_THREADS=1)
    M, linears = wl["M"], wl["linears"]
    sample = min(M, args.cpu_sample_tokens)
    lins = []
    for name, N, K, _, key in linears:
        sc, _src = act_scales(wl["scales"], key, K)
        lin = O.synth_linear(N, K, sc, seed=1234)
        lins.append((lin, O.synth_activations(sample, lin["act_scale"], seed=4321)))
    flops = sum(2.0 * sample * N * K for _, N, K, _, _ in linears)

    def step():
        for lin, A in lins:
            O.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"])
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = flops / dt / 1e12
    line = {"metric": metric_name(args.workload), "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic", "config": bench_config(args, wl),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                             "sample": f"{sample} of {M} tokens through 1 of the step's {wl['layers']} decoder layers "
                                       f"({len(linears)} linears) per step (oracle/mixq_oracle.c, OpenMP, {O.num_threads()} threads)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tokens_per_s": sample / (dt * wl["model_layers"])}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def pack_gpu(torch, W, act_scale):
    """pack_linear_weights (model_config_utils.py:429-466) on the device: fp16 arithmetic is IEEE on both sides, so this
    equals the host packer bit for bit (tests/test_gpu_parity.py::test_bench_packer_matches_checkpoint_packer)."""
    sb = (W.abs().amax(dim=1, keepdim=True) / 127).to(torch.float16).reshape(-1)
    ind = torch.sort(act_scale.float(), stable=True)[1][-128:]
    fw = W[:, ind].contiguous()
    W[:, ind] = 0
    W8 = (W / sb[:, None]).round().clamp(-128, 127).nan_to_num(0).to(torch.int8)
    return W8, sb, fw, ind.to(torch.int32)


def shard_packed(torch, W8, sb, fw, ind, mode, tp, rank):
    """mixq_tensorrt_llm_b200/tp.py (shard_column / shard_row) on device tensors."""
    if tp == 1:
        return W8, sb, fw, ind, (0, W8.shape[1])
    N, K = W8.shape
    if mode == "column":
        lo, hi = rank * (N // tp), (rank + 1) * (N // tp)
        return W8[lo:hi].contiguous(), sb[lo:hi].contiguous(), fw[lo:hi].contiguous(), ind.clone(), (0, K)
    lo, hi = rank * (K // tp), (rank + 1) * (K // tp)
    mine = ((ind >= lo) & (ind < hi)).nonzero().reshape(-1)
    loc = torch.zeros(128, dtype=torch.int32, device=W8.device)
    f = torch.zeros(N, 128, dtype=torch.float16, device=W8.device)
    loc[: mine.numel()] = ind[mine] - lo
    f[:, : mine.numel()] = fw[:, mine]
    return W8[:, lo:hi].contiguous(), sb.clone(), f, loc, (lo, hi)


def mixed_close(torch, got, ref, A, fw, ind):
    """SURVEY.md 8c bound of the mixed output: 1 fp16 ulp of the outlier product + 1 ulp of the result +
    2^-20 * sum|a_j w_j| (accumulation order of the 128-term product); returns (ok, worst ratio, rel-Frobenius)."""
    fa = A[:, ind.long()].float()
    out0 = (fa @ fw.float().t()).half()
    mag = fa.abs() @ fw.float().abs().t()

    def ulp(x):
        x = x.abs().float().clamp_min(2.0 ** -14)
        return torch.exp2(torch.floor(torch.log2(x)) - 10)
    bound = ulp(out0) + ulp(ref) + mag * 2.0 ** -20
    d = (got.float() - ref.float()).abs()
    fin = torch.isfinite(ref.float())
    worst = float((d / bound)[fin].max())
    rel = float(torch.linalg.norm((got.float() - ref.float())[fin].double()) / torch.linalg.norm(ref.float()[fin].double()).clamp_min(1e-30))
    return bool(worst <= 1.0 and rel <= 1e-3 and torch.equal(torch.isfinite(got.float()), fin)), worst, rel


def measure_int8_peak(torch, dev, sustained_s):
    """cuBLASLt INT8 8192^3 (torch._int_mm) in this process: best of 10 (burst) and back to back for `sustained_s`."""
    a = torch.randint(-128, 128, (8192, 8192), dtype=torch.int8, device=dev)
    b = torch.randint(-128, 128, (8192, 8192), dtype=torch.int8, device=dev).t()
    fl = 2.0 * 8192 ** 3
    best = 1e9
    for i in range(13):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); torch._int_mm(a, b); s1.record(); torch.cuda.synchronize()
        if i >= 3:
            best = min(best, s0.elapsed_time(s1))
    out = {"burst": fl / best / 1e9}
    if sustained_s > 0:
        n = max(10, int(sustained_s / (best * 1e-3)))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n):
            torch._int_mm(a, b)
        s1.record(); torch.cuda.synchronize()
        out["sustained"] = fl * n / s0.elapsed_time(s1) / 1e9
        out["sustained_s"] = s0.elapsed_time(s1) / 1e3
    return out


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from mixq_tensorrt_llm_b200 import binding as B

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
    M, linears, n_layers = wl["M"], wl["linears"], wl["layers"]

    # ---- synthetic model: same seed on every rank, each rank keeps its tensor-parallel shard
    g = torch.Generator(device=dev).manual_seed(1234)
    layers, acts, act_src = [], {}, {}
    for name, N, K, mode, key in linears:
        sc, src = act_scales(wl["scales"], key, K)
        act_src[name] = src
        sct = torch.from_numpy(sc).to(dev)
        if name == "up" and "gate" in acts and acts["gate"][0].shape[1] == K:
            acts[name] = acts["gate"]          # gate and up read the same hidden state
        else:
            acts[name] = ((torch.randn(M, K, device=dev, generator=g) * (sct[None, :] / 3.0)).half(), sct)
    for li in range(n_layers):
        lay = []
        for name, N, K, mode, key in linears:
            W = (torch.randn(N, K, device=dev, generator=g) * 0.02).half()
            W8, sb, fw, ind = pack_gpu(torch, W, acts[name][1])
            full = (W8, sb, fw, ind) if (li == 0 and tp > 1 and rank == 0) else None     # kept for the TP parity check
            W8, sb, fw, ind, (klo, khi) = shard_packed(torch, W8, sb, fw, ind, mode, tp, rank)
            lay.append(dict(name=name, N=W8.shape[0], K=W8.shape[1], mode=mode, W8=W8, sb=sb, fw=fw, ind=ind, k=(klo, khi), full=full,
                            n_acc=W8.shape[0], full_n=N, full_k=K))
            del W
        layers.append(lay)
    # The gate and up projections of a layer read the same activations: by default they run as ONE mixq_enqueue_gated call
    # (one quantise launch, one GEMM launch, Out = fp16(silu(gate)) * fp16(up) -- the reference's fused MLP,
    # MixQ/src/mixquant/modules/fused/mlp.py:57-70); --mlp split keeps the two plugin calls.
    plain_layers = layers
    if args.mlp == "fused":
        fused = []
        for lay in layers:
            calls, i = [], 0
            while i < len(lay):
                if (lay[i]["name"] == "gate" and i + 1 < len(lay) and lay[i + 1]["name"] == "up"
                        and (lay[i]["N"], lay[i]["K"]) == (lay[i + 1]["N"], lay[i + 1]["K"])):
                    calls.append(dict(lay[i], name="gate_up", up=lay[i + 1], n_acc=2 * lay[i]["N"]))
                    i += 2
                else:
                    calls.append(lay[i])
                    i += 1
            fused.append(calls)
        layers = fused
    A_in = {}
    for lin in layers[0]:
        a = acts["gate" if lin["name"] == "gate_up" else lin["name"]][0]
        A_in[lin["name"]] = a[:, lin["k"][0]:lin["k"][1]].contiguous() if lin["mode"] == "row" and tp > 1 else a
    for lin in plain_layers[0]:
        A_in.setdefault(lin["name"], A_in.get("gate_up") if lin["name"] in ("gate", "up") else None)
    max_out = max(lin["N"] for lin in layers[0])
    out_buf = torch.empty(M * max_out, dtype=torch.float16, device=dev)
    ws = torch.empty(max(B.gated_workspace_size(M, lin["N"], lin["K"]) if "up" in lin else B.workspace_size(M, lin["N"], lin["K"])
                         for lin in layers[0]), dtype=torch.uint8, device=dev)
    flops_layer = sum(2.0 * M * N * K for _, N, K, _, _ in linears)      # whole job, all ranks together
    flops_step = flops_layer * n_layers
    stream = torch.cuda.current_stream()

    chunks = args.tp_chunks if (tp > 1 and M >= 4096 * args.tp_chunks) else 1
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    # Row-parallel linears: the all-reduce is fused into the GEMM kernel (partial tiles pushed to their owner over
    # NVLink peer memory, fp32 reduce, result written to every rank) -- mixq_enqueue_allreduce.  --tp-reduce nccl keeps
    # the unfused baseline (mixq_enqueue + NCCL all-reduce in overlapped row slabs) for comparison.
    peer, peer_note = None, None
    if tp > 1 and args.tp_reduce == "fused":
        from mixq_tensorrt_llm_b200.peer import PeerBuffers
        try:
            peer = PeerBuffers(M, max(lin["N"] for lin in layers[0] if lin["mode"] == "row"), device=dev)
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

    def run_linear(lin, out=None):
        a = A_in[lin["name"]]
        o = out if out is not None else out_buf[: M * lin["N"]].view(M, lin["N"])
        if "up" in lin:
            u = lin["up"]
            B.enqueue_gated(a, (lin["W8"], lin["sb"], lin["fw"]), (u["W8"], u["sb"], u["fw"]), lin["ind"], o, ws)
        elif tp > 1 and lin["mode"] == "row" and peer is not None:
            B.enqueue_allreduce(a, lin["W8"], lin["sb"], lin["fw"], lin["ind"], ws, peer.peer_group(M, lin["N"]))
        elif tp > 1 and lin["mode"] == "row":
            # The one exchange step of the path.  The token dimension is cut into `chunks` row slabs: the NCCL all-reduce of
            # slab c (on NCCL's stream) overlaps the GEMM of slab c+1; a few SMs are left free for NCCL's channels.
            works, rows = [], M // chunks
            lim = nsm - args.comm_sms if (chunks > 1 and args.comm_sms > 0) else 0
            for c in range(chunks):
                B.enqueue(a[c * rows:(c + 1) * rows], lin["W8"], lin["sb"], lin["fw"], lin["ind"], o[c * rows:(c + 1) * rows], ws, sm_limit=lim)
                works.append(dist.all_reduce(o[c * rows:(c + 1) * rows], async_op=True))
            for w in works:
                w.wait()
        else:
            B.enqueue(a, lin["W8"], lin["sb"], lin["fw"], lin["ind"], o, ws)
        return o

    def step():
        for lay in layers:
            for lin in lay:
                run_linear(lin)

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
    # Decode-sized steps are a few hundred 5-30 us kernels: replay them from a CUDA graph, as a serving runtime (TensorRT)
    # would, so the step is not bounded by Python/launch latency.  With the all-reduce fused into the GEMM kernel a
    # tensor-parallel step holds no NCCL call and is capturable too.
    use_graph = args.graph == "on" or (args.graph == "auto" and M <= 2048 and (world == 1 or peer is not None))
    run_step = step
    if use_graph:
        gs = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        barrier()                     # the step above ran on another stream: fused all-reduce launches must not overlap
        with torch.cuda.stream(gs):
            step()
            gs.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=gs):
                step()
        torch.cuda.synchronize()
        run_step = graph.replay
    warm = max(args.warmup, 3)
    for _ in range(warm):
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

    # ---- per-kernel pass: each kernel of each linear shape timed with CUDA events over a graph that walks the
    # `n_layers` layers' weights (so W streams from HBM exactly as in the step) -- back-to-back launches of ONE kernel
    maxK = max(lin["K"] for lin in layers[0])
    A8 = torch.empty(M * maxK, dtype=torch.int8, device=dev)
    sa = torch.empty(M, dtype=torch.float16, device=dev)
    fpA = torch.empty(M, 128, dtype=torch.float16, device=dev)
    gws = torch.zeros(lib.mixq_decode_workspace_size(min(M, 1024), max_out), dtype=torch.uint8, device=dev)

    def time_kernel(fn, reps):
        """fn(i) launches the kernel on layer i's weights; returns microseconds per launch"""
        for i in range(min(3, reps)):
            fn(i % n_layers)
        torch.cuda.synchronize()
        total = max(reps, n_layers)
        if M <= 2048:
            ks, kg = torch.cuda.Stream(), torch.cuda.CUDAGraph()
            with torch.cuda.stream(ks):
                with torch.cuda.graph(kg, stream=ks):
                    for i in range(total):
                        fn(i % n_layers)
            torch.cuda.synchronize()
            kg.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                kg.replay()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e3 / (3 * total)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(total):
            fn(i % n_layers)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / total

    per_linear, tot_gemm, tot_quant, floor_us = {}, 0.0, 0.0, 0.0
    reps = 32 if M <= 2048 else max(args.steps, 5)
    for idx, lin in enumerate(layers[0]):
        name, Ns, Ks = lin["name"], lin["N"], lin["K"]
        a8 = A8[: M * Ks].view(M, Ks)
        a = A_in[name]
        o = out_buf[: M * Ns].view(M, Ns)
        qu = time_kernel(lambda i: B.quant_extract(a, layers[i][idx]["ind"], a8, sa, fpA), reps)
        B.quant_extract(a, lin["ind"], a8, sa, fpA)
        if "up" in lin:
            gscratch = torch.empty(M * Ns * 2, dtype=torch.uint8, device=dev) if M > 1024 else None   # two GEMMs + multiply above 1024 tokens

            def gemm(i):
                L, U = layers[i][idx], layers[i][idx]["up"]
                B.gemm_dequant_gated(a8, sa, fpA, (L["W8"], L["sb"], L["fw"]), (U["W8"], U["sb"], U["fw"]), o, scratch=gscratch)
        else:
            def gemm(i):
                L = layers[i][idx]
                B.gemm_dequant(a8, L["W8"], sa, L["sb"], fpA, L["fw"], o, workspace=gws)
        gu = time_kernel(gemm, reps)
        tot_gemm += gu
        tot_quant += qu
        n_acc = lin["n_acc"]      # accumulator columns: 2 N for the gated call (gate and up)
        per_linear[name] = {"N": Ns, "K": Ks, "acc_columns": n_acc, "gemm_us": round(gu, 2),
                            "gemm_tflops": round(2.0 * M * n_acc * Ks / gu / 1e6, 1),
                            "quant_us": round(qu, 2), "quant_gbs": round((3.0 * M * Ks + 258.0 * M) / qu / 1e3, 1)}

    # ---- INT8 tensor peak measured in this process (cuBLASLt 8192^3): burst for short timed regions at full clocks,
    # sustained (power-capped clocks) for long ones -- the denominator follows the timed region's length
    region_s = ms * 1e-3
    try:
        ipk = measure_int8_peak(torch, dev, args.peak_sustained_s if rank == 0 else 0.0)
    except Exception as e:  # library may lack an int8 path
        ipk = {"error": repr(e)[:80]}
    if "burst" in ipk:
        use_sustained = region_s >= 1.0 and "sustained" in ipk
        int8_peak = ipk["sustained"] if use_sustained else ipk["burst"]
        how = ("sustained over %.1f s" % ipk.get("sustained_s", 0)) if use_sustained else "best of 10 (burst)"
        peak_src = f"cuBLASLt INT8 8192^3 measured in this process, {how}: the timed region lasted {region_s * 1e3:.0f} ms"
    else:
        int8_peak = 2.0 * pk["bf16_burst"]
        peak_src = f"2 x bf16_tflops of MEASURED_PEAKS.json ({pk['src']}); cuBLASLt INT8 unavailable"
    def call_bytes(v):
        """compulsory HBM bytes of a call: a gated call reads A once and both weight sets, and writes ONE [M, N] result"""
        if v["acc_columns"] == v["N"]:
            return linear_bytes(M, v["N"], v["K"])
        return 2.0 * M * v["K"] + 2.0 * v["N"] * v["K"] + 512.0 * v["N"] + 4.0 * v["N"] + 512 + 2.0 * M * v["N"]
    for name, v in per_linear.items():
        Ns, Ks = v["N"], v["K"]
        t_tensor = 2.0 * M * v["acc_columns"] * Ks / int8_peak / 1e6          # us
        t_hbm = call_bytes(v) / pk["hbm"] / 1e3                                # us
        v["floor_us"] = round(max(t_tensor, t_hbm), 2)
        v["floor_bound"] = "tensor" if t_tensor >= t_hbm else "hbm"
        v["frac_of_floor"] = round(max(t_tensor, t_hbm) / (v["gemm_us"] + v["quant_us"]), 4)
        floor_us += max(t_tensor, t_hbm)
    dom = max(per_linear.items(), key=lambda kv: kv[1]["gemm_us"])
    gemm_flops = sum(2.0 * M * v["acc_columns"] * v["K"] for v in per_linear.values())
    achieved = gemm_flops / tot_gemm / 1e6
    layer_ms = ms_step / n_layers
    hbm_bound = all(v["floor_bound"] == "hbm" for v in per_linear.values())
    if hbm_bound:
        ach_b = sum(call_bytes(v) for v in per_linear.values()) / (tot_gemm + tot_quant) / 1e3
        head = {"bound": "hbm", "achieved": round(ach_b, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(ach_b / pk["hbm"], 4),
                "peak_source": f"hbm_gbs of MEASURED_PEAKS.json ({pk['src']})"}
    else:
        head = {"bound": "tensor", "achieved": round(achieved, 1), "peak": round(int8_peak, 1), "unit": "TFLOP/s",
                "frac": round(achieved / int8_peak, 4), "peak_source": peak_src}
    roofline = {**head,
                "kernel": ("mixq_gemm_dequant_fat_kernel (tcgen05 kind::i8 + kind::f16, cta_group::2, one 256 x Nt<=336 tile per CTA pair and wave; "
                           "256x128 pair tiles where the cost model prefers them)" if 128 < M <= 1024 else
                           "mixq_gemm_dequant_kernel (tcgen05 kind::i8 + kind::f16, 128x128 tiles)" if M <= 128 else
                           "mixq_gemm_dequant_streamk_kernel, whole-tile schedule (tcgen05 kind::i8 + kind::f16, cta_group::2, TMA-store epilogue)"),
                "gemm_tflops": round(achieved, 1), "frac_of_spec_4500": round(achieved / SPEC_INT8_TOPS, 4),
                "int8_peak_measured": {k: (round(v, 1) if isinstance(v, float) else v) for k, v in ipk.items()},
                "frac_of_2x_bf16_burst_measured_peaks": round(achieved / (2.0 * pk["bf16_burst"]), 4),
                # whole step against the per-shape floors max(t_tensor, t_HBM)
                "step_floor_us_per_layer": round(floor_us, 2), "step_us_per_layer": round(layer_ms * 1e3, 2),
                "step_frac_of_floor": round(floor_us / (layer_ms * 1e3), 4),
                "traffic": None, "share_of_step": round(tot_gemm / (tot_gemm + tot_quant), 4),
                "quant_kernel": {"bound": "hbm", "achieved": round(sum((3.0 * M * v["K"] + 258.0 * M) for v in per_linear.values()) / tot_quant / 1e3, 1),
                                 "peak": pk["hbm"], "unit": "GB/s", "note": "latency-bound below ~8 MB per launch"},
                "per_linear": per_linear, "dominant": dom[0]}
    tr = ROOT / "profiles" / "r2_traffic.json"
    if tr.exists():
        try:
            d = json.loads(tr.read_text()).get(args.workload)
            if d:
                roofline["traffic"] = d["dram_read_bytes"] + d["dram_write_bytes"]
                roofline["traffic_detail"] = d
        except (ValueError, KeyError):
            pass

    # ---- parity of the benchmarked shapes (after the timed region): layer 0, every linear, whole batch --
    # ours vs the reference's own kernels on this GPU (oracle/_ref: kernel/i8gemm.cu + cuBLAS, enqueueImpl :518-532);
    # quantise/extract bit-exact, output within the stated fp16 bound.  Tensor parallel: vs the single-GPU call.
    parity = None
    if not args.no_parity:
        try:
            parity = {"checked": True, "ok": True, "linears": {}}
            sys.path.insert(0, str(ROOT / "tests"))
            import refgpu
            if world == 1 and refgpu.available():
                parity["against"] = "reference kernels (oracle/_ref: i8gemm.cu + cuBLAS fp16) on the same GPU, full batch"
                rws = torch.empty(max(refgpu.load().ref_workspace_size(M, lin["N"], lin["K"]) for lin in layers[0]), dtype=torch.uint8, device=dev)
                for lin in layers[0]:
                    a = A_in[lin["name"]]
                    got = run_linear(lin).clone()
                    rq, rsa = refgpu.int8quant(a)
                    a8 = A8[: M * lin["K"]].view(M, lin["K"])
                    B.quant_extract(a, lin["ind"], a8, sa, fpA)
                    bit = bool(torch.equal(rq, a8) and torch.equal(rsa.view(torch.int16), sa.view(torch.int16))
                               and torch.equal(refgpu.extract(a, lin["ind"]).view(torch.int16), fpA.view(torch.int16)))
                    rows = slice(0, M) if M <= 8192 else slice(M - 4096, M)     # the fp32 bound needs M x N floats
                    if "up" in lin:
                        # reference sequence of the fused MLP half (MixQ/src/mixquant/modules/fused/mlp.py:57-70): up through the
                        # plugin kernels, gate through the reference's GemmDequantSilu (oracle/_ref/libref_mixsrc.so) over the
                        # reference's own INT8 codes and cuBLAS outlier product, then an fp16 multiply
                        U = lin["up"]
                        ref_u = refgpu.enqueue(a, U["W8"], U["sb"], U["fw"], lin["ind"], None, rws)
                        out0_g = refgpu.enqueue(a, torch.zeros_like(lin["W8"]), lin["sb"], lin["fw"], lin["ind"], None, rws)
                        if refgpu.mixsrc_available():
                            ref_g = refgpu.int8_fused_dequant_silu(rq, lin["W8"], rsa, lin["sb"], out0_g)
                            how = "GemmDequantSilu"
                        else:
                            x = refgpu.enqueue(a, lin["W8"], lin["sb"], lin["fw"], lin["ind"], None, rws).float()
                            ref_g = (x * torch.sigmoid(x)).half()
                            how = "torch SiLU of the plugin output (libref_mixsrc.so missing: one extra rounding)"
                        ref = ref_g * ref_u
                        fa = a[rows][:, lin["ind"].long()].float()

                        def ulp(x):
                            return torch.exp2(torch.floor(torch.log2(x.abs().float().clamp_min(2.0 ** -14))) - 10)
                        out0_u = (fa @ U["fw"].float().t()).half()
                        bg = 2 * ulp(ref_g[rows]) + 1.2 * ulp(out0_g[rows]) + (fa.abs() @ lin["fw"].float().abs().t()) * 2.0 ** -20
                        bu = ulp(ref_u[rows]) + ulp(out0_u) + (fa.abs() @ U["fw"].float().abs().t()) * 2.0 ** -20
                        bound = ref_u[rows].float().abs() * bg + ref_g[rows].float().abs() * bu + bg * bu + ulp(ref[rows])
                        d = (got[rows].float() - ref[rows].float()).abs()
                        worst = float((d / bound).max())
                        rel = float(torch.linalg.norm(d.double()) / torch.linalg.norm(ref[rows].float().double()).clamp_min(1e-30))
                        ok = bool(worst <= 1.0 and rel <= 2e-3)
                        parity.setdefault("gated_against", how)
                        del ref_u, ref_g, out0_g, out0_u, bg, bu, bound, d, fa
                    else:
                        ref = refgpu.enqueue(a, lin["W8"], lin["sb"], lin["fw"], lin["ind"], None, rws)
                        ok, worst, rel = mixed_close(torch, got[rows], ref[rows], a[rows], lin["fw"], lin["ind"])
                    del rq
                    same = float((got.view(torch.int16) == ref.view(torch.int16)).float().mean())
                    relf = float(torch.linalg.norm((got.float() - ref.float()).flatten()[:: max(1, got.numel() // (1 << 26))].double()) /
                                 torch.linalg.norm(ref.float().flatten()[:: max(1, got.numel() // (1 << 26))].double()))
                    parity["linears"][lin["name"]] = {"quant_bit_exact": bit, "within_bound": ok, "worst_over_bound": round(worst, 3),
                                                      "rel_frobenius": rel, "rel_frobenius_all_rows": relf, "bit_identical_frac": same}
                    parity["ok"] = parity["ok"] and ok and bit and relf <= (2e-3 if "up" in lin else 1e-3)
                    del got, ref
                del rws
            elif world > 1:
                parity["against"] = ("rank 0: column-parallel = bit-identical slices of the single-GPU call on the unsharded linear; row-parallel = "
                                     "bit-identical to the rank-order fp32 sum of the single-GPU calls on the shards")
                for lin in layers[0]:
                    o = run_linear(lin)
                    if lin["mode"] == "row":
                        res = (peer.out(M, lin["N"]) if peer is not None else o).clone()
                    else:
                        parts = [torch.empty_like(o) for _ in range(world)]
                        dist.all_gather(parts, o.contiguous())
                        res = torch.cat(parts, dim=1)
                    if rank == 0:
                        W8f, sbf, fwf, indf = lin["full"]
                        ref = torch.empty(M, W8f.shape[0], dtype=torch.float16, device=dev)
                        if "up" in lin:
                            wsf = torch.empty(B.gated_workspace_size(M, W8f.shape[0], W8f.shape[1]), dtype=torch.uint8, device=dev)
                            B.enqueue_gated(acts["gate"][0], (W8f, sbf, fwf), lin["up"]["full"][:3], indf, ref, wsf)
                        else:
                            wsf = torch.empty(B.workspace_size(M, W8f.shape[0], W8f.shape[1]), dtype=torch.uint8, device=dev)
                            B.enqueue(acts[lin["name"]][0], W8f, sbf, fwf, indf, ref, wsf)
                        rel = float(torch.linalg.norm((res.float() - ref.float()).double()) / torch.linalg.norm(ref.float().double()))
                        same = bool(torch.equal(res.view(torch.int16), ref.view(torch.int16)))
                        if lin["mode"] == "row":
                            # Row-parallel: every rank quantises ITS K slice with its own per-token scale, so the
                            # result is a different -- finer -- quantisation of the same product than the unsharded call and
                            # differs from it at the level of the INT8 quantisation error (reported, not bounded).  The
                            # checkable statement: the all-reduced result equals fp16(sum_r fp32(partial_r)) in rank order,
                            # where partial_r is the single-GPU mixq_enqueue of shard r -- bit for bit.
                            acc = torch.zeros(M, W8f.shape[0], dtype=torch.float32, device=dev)
                            part = torch.empty(M, W8f.shape[0], dtype=torch.float16, device=dev)
                            for r in range(world):
                                w8r, sbr, fwr, indr, (lo, hi) = shard_packed(torch, W8f, sbf, fwf, indf, "row", world, r)
                                B.enqueue(acts[lin["name"]][0][:, lo:hi].contiguous(), w8r, sbr, fwr, indr, part, wsf)
                                acc += part.float()
                            want = acc.half()
                            same_sum = bool(torch.equal(res.view(torch.int16), want.view(torch.int16)))
                            nbad = int((res.view(torch.int16) != want.view(torch.int16)).sum())
                            parity["linears"][lin["name"]] = {"bit_identical_to_rank_order_sum_of_shard_results": same_sum, "mismatches": nbad,
                                                              "rel_frobenius_vs_unsharded_call": rel, "ok": same_sum}
                            okl = same_sum
                            del acc, part, want
                        else:
                            okl = same and rel == 0.0          # column-parallel shards are bit-identical slices of the unsharded result
                            parity["linears"][lin["name"]] = {"rel_frobenius": rel, "bit_identical": same, "ok": okl}
                        parity["ok"] = parity["ok"] and okl
                        del ref, wsf
                    barrier()
            else:
                parity = {"checked": False, "reason": "oracle/_ref not built on this box"}
        except Exception as e:
            parity = {"checked": False, "error": repr(e)[:300]}

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (pinned): H2D + enqueue + D2H; linears that consume
    # the same activations (gate and up) share one upload (mixq_linears_host)
    e2e = None
    if not args.no_e2e:
        try:
            groups = []   # lists of call indices sharing their input
            for i, lin in enumerate(layers[0]):
                if lin["name"] == "up" and groups and layers[0][groups[-1][0]]["name"] == "gate":
                    groups[-1].append(i)
                else:
                    groups.append([i])
            hA, hO = {}, {}
            for gi, grp in enumerate(groups):
                lin0 = layers[0][grp[0]]
                src = acts["gate" if lin0["name"] == "gate_up" else lin0["name"]][0]
                if tp > 1 and lin0["mode"] == "column":
                    rows = M // tp                                   # each rank uploads 1/tp of the rows; NVLink all-gather
                    hA[gi] = torch.empty(rows, src.shape[1], dtype=torch.float16).pin_memory()
                    hA[gi].copy_(src[rank * rows:(rank + 1) * rows])
                else:
                    hA[gi] = torch.empty(A_in[lin0["name"]].shape, dtype=torch.float16).pin_memory()
                    hA[gi].copy_(A_in[lin0["name"]])
                for i in grp:
                    lin = layers[0][i]
                    rows_out = M // tp if (tp > 1 and lin["mode"] == "row") else M   # a rank downloads its share of a replicated result
                    hO[i] = torch.empty(rows_out, lin["N"], dtype=torch.float16).pin_memory()
            h2d = sum(h.numel() * 2 for h in hA.values()) * n_layers * tp
            d2h = sum(h.numel() * 2 for h in hO.values()) * n_layers * tp
            if tp == 1:
                # four times the largest call's scratch (+ alignment): consecutive asynchronous calls take the four parts in turn
                scratch = torch.empty(4 * max(B.gated_host_scratch_size(M, layers[0][grp[0]]["N"], layers[0][grp[0]]["K"]) if "up" in layers[0][grp[0]] else
                                              B.linears_host_scratch_size(M, [layers[0][i]["N"] for i in grp], layers[0][grp[0]]["K"]) for grp in groups) + 1024,
                                      dtype=torch.uint8, device=dev)
                # decode-sized calls are queued (MIXQ_FLAG_HOST_ASYNC) and drained once per step: uploads, kernels and downloads
                # of consecutive calls overlap; bulk calls pipeline their own row slabs and stay synchronous
                hflags = B.FLAG_HOST_ASYNC if M <= 4096 else 0

                def table(L):
                    return B.make_tensors(None, L["W8"], L["sb"], L["fw"], L["ind"], None)

                def e2e_step():
                    for lay in layers:
                        for gi, grp in enumerate(groups):
                            if "up" in lay[grp[0]]:
                                B.gated_host(table(lay[grp[0]]), table(lay[grp[0]]["up"]), hA[gi], hO[grp[0]], scratch, flags=hflags, stream=stream)
                            else:
                                B.linears_host([table(lay[i]) for i in grp], hA[gi], [hO[i] for i in grp], scratch, flags=hflags, stream=stream)
                    if hflags:
                        B.host_drain(stream=stream)
                path = ("mixq_linears_host / mixq_gated_host (C ABI): pinned host A -> H2D once per distinct activation -> mixq_enqueue per "
                        "linear (mixq_enqueue_gated for gate+up) -> D2H Out" +
                        ("; calls queued with MIXQ_FLAG_HOST_ASYNC and drained once per step (upload of call i+1 | kernels | download of "
                         "call i overlap)" if hflags else ""))
            else:
                # three streams: uploads | all-gather + kernels (the caller's stream) | downloads.  Device activations and
                # outputs are double-buffered by layer parity, so the upload of the next call and the download of the previous
                # one use both directions of this rank's PCIe link while the kernels run; a buffer is rewritten only after the
                # event that says its last reader has finished.
                n_slot = 2 if M <= 4096 else 1
                dA = {}
                for gi, grp in enumerate(groups):
                    lin0 = layers[0][grp[0]]
                    dA[gi] = [torch.empty(M, lin0["full_k"] if lin0["mode"] == "column" else lin0["K"], dtype=torch.float16, device=dev)
                              for _ in range(n_slot)]
                dO = {i: [torch.empty(M, layers[0][i]["N"], dtype=torch.float16, device=dev) for _ in range(n_slot)]
                      for grp in groups for i in grp}
                s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
                ev_read, ev_drained = {}, {}     # (group, slot) -> kernels have read dA;  result address -> download finished
                ev_pool = {}                     # upload-done / kernels-done events, re-recorded every layer (a wait binds to the
                                                 # record that preceded it)

                def e2e_body():
                    cur = torch.cuda.current_stream()
                    ev_read.clear()              # a step ends drained: no hazard crosses steps (and a captured step must not
                    ev_drained.clear()           # depend on events recorded outside the capture)
                    s_up.wait_stream(cur)        # fork (also what makes the side streams part of a capture)
                    s_dn.wait_stream(cur)
                    for li, lay in enumerate(layers):
                        slot = li % n_slot
                        for gi, grp in enumerate(groups):
                            lin0 = lay[grp[0]]
                            a = dA[gi][slot]
                            rows = M // tp
                            with torch.cuda.stream(s_up):
                                if (gi, slot) in ev_read:
                                    s_up.wait_event(ev_read[(gi, slot)])
                                if lin0["mode"] == "column":
                                    a[rank * rows:(rank + 1) * rows].copy_(hA[gi], non_blocking=True)
                                else:
                                    a.copy_(hA[gi], non_blocking=True)
                                ev_up = ev_pool.setdefault(("up", gi, slot), torch.cuda.Event())
                                ev_up.record(s_up)
                            cur.wait_event(ev_up)
                            if lin0["mode"] == "column":
                                dist.all_gather_into_tensor(a, a[rank * rows:(rank + 1) * rows])
                            for i in grp:
                                lin = lay[i]
                                fused = lin["mode"] == "row" and peer is not None
                                res = peer.out(M, lin["N"]) if fused else dO[i][slot]
                                key = res.data_ptr()
                                if key in ev_drained:
                                    cur.wait_event(ev_drained[key])
                                if fused:
                                    B.enqueue_allreduce(a, lin["W8"], lin["sb"], lin["fw"], lin["ind"], ws, peer.peer_group(M, lin["N"]))
                                elif "up" in lin:
                                    B.enqueue_gated(a, (lin["W8"], lin["sb"], lin["fw"]), (lin["up"]["W8"], lin["up"]["sb"], lin["up"]["fw"]),
                                                    lin["ind"], res, ws)
                                else:
                                    B.enqueue(a, lin["W8"], lin["sb"], lin["fw"], lin["ind"], res, ws)
                                    if lin["mode"] == "row":
                                        dist.all_reduce(res)
                                ev_done = ev_pool.setdefault(("done", i, slot), torch.cuda.Event())
                                ev_done.record(cur)
                                with torch.cuda.stream(s_dn):
                                    s_dn.wait_event(ev_done)
                                    if lin["mode"] == "row":
                                        hO[i].copy_(res[rank * rows:(rank + 1) * rows], non_blocking=True)
                                    else:
                                        hO[i].copy_(res, non_blocking=True)
                                    ev_drained.setdefault(key, torch.cuda.Event()).record(s_dn)
                            ev_read.setdefault((gi, slot), torch.cuda.Event()).record(cur)
                    cur.wait_stream(s_up)        # join
                    cur.wait_stream(s_dn)

                # (Capturing this step -- copies, NCCL all-gathers, kernels on three streams -- in a CUDA graph was tried: the replay
                # was slower than direct launches, 18.5 vs 12.6 ms at 2 ranks, and the process hung at exit; not used.)
                def e2e_step():
                    e2e_body()
                    torch.cuda.synchronize()
                path = ("pinned host A -> each rank uploads 1/tp of a replicated activation, NVLink all-gather -> mixq_enqueue / "
                        "mixq_enqueue_allreduce -> each rank downloads its shard (column) or 1/tp of the rows (row); uploads, kernels "
                        "and downloads of consecutive calls overlap on three streams (direct launches)")
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
                   "tokens_per_s": M / (float(tt.item()) * wl["model_layers"] / n_layers), "path": path}
            # the host buffers hold the LAST layer's results: recompute that layer call by call, nothing overlapped, and compare
            # bit for bit (a buffer rewritten before its download finished, or read before its upload, would show here)
            lay, bad = layers[-1], 0
            for gi, grp in enumerate(groups):
                lin0 = lay[grp[0]]
                rows = M // tp
                a = torch.empty(M, lin0["full_k"] if (tp > 1 and lin0["mode"] == "column") else lin0["K"], dtype=torch.float16, device=dev)
                if tp > 1 and lin0["mode"] == "column":
                    a[rank * rows:(rank + 1) * rows].copy_(hA[gi])
                    dist.all_gather_into_tensor(a, a[rank * rows:(rank + 1) * rows].clone())
                else:
                    a.copy_(hA[gi])
                for i in grp:
                    lin = lay[i]
                    chk = torch.empty(M, lin["N"], dtype=torch.float16, device=dev)
                    if tp > 1 and lin["mode"] == "row" and peer is not None:
                        B.enqueue_allreduce(a, lin["W8"], lin["sb"], lin["fw"], lin["ind"], ws, peer.peer_group(M, lin["N"]))
                        chk = peer.out(M, lin["N"])
                    elif "up" in lin:
                        B.enqueue_gated(a, (lin["W8"], lin["sb"], lin["fw"]), (lin["up"]["W8"], lin["up"]["sb"], lin["up"]["fw"]), lin["ind"], chk, ws)
                    else:
                        B.enqueue(a, lin["W8"], lin["sb"], lin["fw"], lin["ind"], chk, ws)
                        if tp > 1 and lin["mode"] == "row":
                            dist.all_reduce(chk)
                    torch.cuda.synchronize()
                    if tp > 1 and lin["mode"] == "row":
                        chk = chk[rank * rows:(rank + 1) * rows]
                    bad += int((chk.cpu().view(torch.int16) != hO[i].view(torch.int16)).sum().item())
            tb = torch.tensor([bad], device=dev)
            if world > 1:
                dist.all_reduce(tb)
            e2e["results_checked"] = "last layer recomputed call by call: %d differing outputs" % int(tb.item())
            if int(tb.item()) != 0:
                e2e["value"] = None
                e2e["error"] = "pipelined host-buffer results differ from the call-by-call results"
        except Exception as e:
            e2e = {"value": None, "unit": UNIT, "error": repr(e)[:300]}

    # ---- same-box GPU baseline: the reference's own kernels recompiled for sm_100a (oracle/_ref), N=1 only
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_ref_gpu:
        try:
            sys.path.insert(0, str(ROOT / "tests"))
            import refgpu
            if refgpu.available():
                rws = torch.empty(max(refgpu.load().ref_workspace_size(M, lin["N"], lin["K"]) for lin in plain_layers[0]), dtype=torch.uint8, device=dev)

                def ref_step():
                    for lay in plain_layers:
                        for lin in lay:
                            refgpu.enqueue(A_in[lin["name"]], lin["W8"], lin["sb"], lin["fw"], lin["ind"], out_buf[: M * lin["N"]].view(M, lin["N"]), rws)
                ref_step(); torch.cuda.synchronize()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record()
                for _ in range(2):
                    ref_step()
                r1.record(); torch.cuda.synchronize()
                rms = r0.elapsed_time(r1) / 2
                ref_gpu = {"what": "reference kernel/i8gemm.cu + cuBLAS fp16, 4 launches per linear, five linears per layer (gate and up as two "
                                   "plugin calls, no SiLU / multiply), recompiled for sm_100a, direct launches",
                           "ms_per_step": rms, "value": flops_step / (rms * 1e-3) / 1e12, "unit": UNIT, "speedup_ours": rms / ms_step}
                del rws
        except Exception as e:
            ref_gpu = {"error": repr(e)[:200]}

    # ---- CPU baseline (rank 0, N=1): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        O.build()
        O.set_threads()
        sample = min(M, args.cpu_baseline_tokens)   # ~10 s of host work on 16 cores
        t_cpu = 0.0
        for name, N, K, _, key in linears:
            lin = O.synth_linear(N, K, act_scales(wl["scales"], key, K)[0], seed=1)
            A = O.synth_activations(sample, lin["act_scale"])
            t0 = time.perf_counter()
            O.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"])
            t_cpu += time.perf_counter() - t0
        cpu = {"value": sum(2.0 * sample * N * K for _, N, K, _, _ in linears) / t_cpu / 1e12, "unit": UNIT,
               "cores": O.num_threads(), "kind": "port",
               "sample": f"{sample} of {M} tokens, one pass through one layer's {len(linears)} linears (oracle/mixq_oracle.c, OpenMP)",
               "seconds": round(t_cpu, 2)}

    if rank == 0:
        line = {"metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
                "config": bench_config(args, wl),
                "run_details": dict(
                    comm_sms=args.comm_sms if (tp > 1 and chunks > 1 and peer is None) else 0,
                    **({"fused_allreduce_unavailable": peer_note} if peer_note else {}),
                    launch="cuda-graph replay of the step" if use_graph else "direct launches",
                    mlp=("gate and up projections in one mixq_enqueue_gated call: Out = fp16(silu(gate)) * fp16(up), the reference's fused MLP "
                         "(MixQ/src/mixquant/modules/fused/mlp.py:57-70); FLOPs counted for both projections" if args.mlp == "fused"
                         else "gate and up as two mixq_enqueue calls"),
                    calls_per_layer=len(layers[0]),
                    parallelism=("single" if tp == 1 else
                                 f"tp{tp} (column: no collective; row: all-reduce fused into the GEMM kernel over NVLink peer memory)"
                                 if peer is not None else
                                 f"tp{tp} (column: no collective; row: one NCCL all-reduce in {chunks} overlapped row slabs)"),
                    inputs=f"SURVEY 8d recipe: W ~ N(0, 0.02^2) packed as pack_linear_weights, A = randn * act_scale / 3 ({sorted(set(act_src.values()))})",
                    l2=("inputs larger than L2 (activations %.0f MB per linear), no flush" % (M * 4096 * 2 / 1e6) if M >= 16384 else
                        "%d layers with their own weights (%.1f GB per step) stream from HBM; activations (%.1f MB) are L2-resident"
                        % (n_layers, sum(N * K for _, N, K, _, _ in linears) * n_layers / tp / 1e9, M * 4096 * 2 / 1e6))),
                "tokens_per_s": M / (ms_step * 1e-3 * wl["model_layers"] / n_layers),
                "ms_per_layer": layer_ms, "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "parity_checked": bool(parity and parity.get("checked") and parity.get("ok")), "parity": parity, "ref_gpu": ref_gpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--layers", type=int, default=0, help="decoder layers per step (0 = the workload's default)")
    ap.add_argument("--cpu-sample-tokens", type=int, default=512, help="tokens per step of the --impl reference arm")
    ap.add_argument("--cpu-baseline-tokens", type=int, default=4096, help="token sample of the cpu_baseline leg of our arm")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--peak-sustained-s", type=float, default=2.0, help="length of the sustained INT8 peak measurement")
    ap.add_argument("--tp-chunks", type=int, default=4, help="row slabs per row-parallel linear (comm/compute overlap)")
    ap.add_argument("--comm-sms", type=int, default=40, help="SMs left free for NCCL when slabs overlap (tensor parallel only)")
    ap.add_argument("--tp-reduce", default="fused", choices=["fused", "nccl"],
                    help="row-parallel linears: all-reduce fused into the GEMM kernel (default) or NCCL after it")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (auto: decode-sized M)")
    ap.add_argument("--mlp", default="auto", choices=["auto", "fused", "split"],
                    help="gate and up projections: one mixq_enqueue_gated call (auto: decode batches, M <= 1024, where it is one GEMM "
                         "launch) or two mixq_enqueue calls (auto: larger M, where the gated call is two GEMMs plus an elementwise pass)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.mlp == "auto":
        args.mlp = "fused" if wl["M"] <= 1024 else "split"
    if args.layers > 0:
        wl["layers"] = args.layers
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
