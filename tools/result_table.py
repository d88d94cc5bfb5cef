#!/usr/bin/env python
"""BASELINE.md 3.3 as one artefact: per linear shape of SURVEY.md 8 (Llama-2-7B, Qwen2-7B, Llama-2-70B / TP 8 shards) and
M in {32, 512, 65536}, on ONE B200:
    ours (mixq_enqueue: quantise + GEMM launches) time, W8A8O16 TFLOPS = 2MNK / t, fraction of the INT8 peak measured in this
    process (cuBLASLt 8192^3, burst) and of the 4.5 POPS spec, HBM GB/s of the compulsory bytes (the bound at M = 32),
    the reference's own kernels on the same GPU (oracle/_ref: gather, cuBLAS fp16, int8quant, CUTLASS GemmDequant),
    the CPU oracle port on a bounded token sample, and the error of sampled rows against the oracle.
Multi-GPU rows come from `bench.py --gpus N` lines (per-linear entries of `roofline.per_linear`): pass them with --merge.

    python tools/result_table.py [--out profiles/r2_result_table] [--m 32,512,65536] [--merge bench_tp2.json ...]
Writes <out>.json and <out>.md.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SHAPES = [  # model, linear, N, K, act-scale key
    ("Llama-2-7b", "qkv", 12288, 4096, "self_attn.q_proj"), ("Llama-2-7b", "o", 4096, 4096, "self_attn.o_proj"),
    ("Llama-2-7b", "gate/up", 11008, 4096, "mlp.gate_proj"), ("Llama-2-7b", "down", 4096, 11008, "mlp.down_proj"),
    ("qwen2-7b-instruct", "qkv", 4608, 3584, "self_attn.q_proj"), ("qwen2-7b-instruct", "o", 3584, 3584, "self_attn.o_proj"),
    ("qwen2-7b-instruct", "gate/up", 18944, 3584, "mlp.gate_proj"), ("qwen2-7b-instruct", "down", 3584, 18944, "mlp.down_proj"),
    ("Llama-2-70b", "qkv / TP8 (column)", 1280, 8192, "self_attn.q_proj"), ("Llama-2-70b", "o / TP8 (row)", 8192, 1024, "self_attn.o_proj"),
    ("Llama-2-70b", "gate/up / TP8 (column)", 3584, 8192, "mlp.gate_proj"), ("Llama-2-70b", "down / TP8 (row)", 8192, 3584, "mlp.down_proj"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r2_result_table"))
    ap.add_argument("--m", default="32,512,65536")
    ap.add_argument("--cpu-tokens", type=int, default=128)
    ap.add_argument("--merge", nargs="*", default=[])
    args = ap.parse_args()
    import torch
    import bench
    import refgpu
    from mixq_tensorrt_llm_b200 import binding as B
    from oracle import oracle as O
    B.require_device()
    O.build()
    O.set_threads()
    dev = torch.device("cuda", 0)
    pk = bench.peaks()
    int8_peak = bench.measure_int8_peak(torch, dev, 0.0)["burst"]
    have_ref = refgpu.available()
    rows = []
    for model, lname, N, K, key in SHAPES:
        sc, _ = bench.act_scales(model, key, K)
        if sc.shape[0] != K:
            sc = sc[:K]
        sct = torch.from_numpy(sc).to(dev)
        g = torch.Generator(device=dev).manual_seed(1234)
        sets = []
        n_sets = max(1, min(4, int(200e6 // (N * K)) + 1))          # rotate weight sets so that decode-sized calls stream W from HBM
        for _ in range(n_sets):
            W = (torch.randn(N, K, device=dev, generator=g) * 0.02).half()
            sets.append(bench.pack_gpu(torch, W, sct))
            del W
        for M in [int(x) for x in args.m.split(",")]:
            A = (torch.randn(M, K, device=dev, generator=g) * (sct[None, :] / 3.0)).half()
            out = torch.empty(M, N, dtype=torch.float16, device=dev)
            ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=dev)

            def ours(i):
                W8, sb, fw, ind = sets[i % n_sets]
                B.enqueue(A, W8, sb, fw, ind, out, ws)

            def ref(i):
                W8, sb, fw, ind = sets[i % n_sets]
                refgpu.enqueue(A, W8, sb, fw, ind, out, rws)

            def timed(fn, graph):
                for i in range(2):
                    fn(i)
                torch.cuda.synchronize()
                reps = 8 if M <= 2048 else 3
                if graph:
                    s, gr = torch.cuda.Stream(), torch.cuda.CUDAGraph()
                    with torch.cuda.stream(s):
                        with torch.cuda.graph(gr, stream=s):
                            for i in range(reps):
                                fn(i)
                    torch.cuda.synchronize()
                    gr.replay()
                    run, n = gr.replay, reps
                else:
                    def run():
                        for i in range(reps):
                            fn(i)
                    n = reps
                best = 1e30
                for _ in range(3):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); run(); b.record(); torch.cuda.synchronize()
                    best = min(best, a.elapsed_time(b) * 1e3 / n)
                return best
            t_us = timed(ours, M <= 2048)
            r = dict(model=model, linear=lname, N=N, K=K, M=M, gpus=1, ours_us=round(t_us, 2),
                     tflops=round(2.0 * M * N * K / t_us / 1e6, 1), launch="graph replay over %d weight sets" % n_sets if M <= 2048 else "direct")
            r["frac_of_measured_int8_peak"] = round(r["tflops"] / int8_peak, 4)
            r["frac_of_spec_4500"] = round(r["tflops"] / 4500.0, 4)
            r["hbm_gbs"] = round(bench.linear_bytes(M, N, K) / t_us / 1e3, 1)
            r["frac_of_measured_hbm"] = round(r["hbm_gbs"] / pk["hbm"], 4)
            if have_ref:
                rws = torch.empty(refgpu.load().ref_workspace_size(M, N, K), dtype=torch.uint8, device=dev)
                r["ref_kernels_us"] = round(timed(ref, False), 2)
                r["speedup_vs_ref_kernels"] = round(r["ref_kernels_us"] / t_us, 2)
                del rws
            # error of sampled rows against the CPU oracle + CPU time on a bounded sample
            W8, sb, fw, ind = sets[0]
            B.enqueue(A, W8, sb, fw, ind, out, ws)
            torch.cuda.synchronize()
            n_s = min(M, args.cpu_tokens)
            idx = torch.linspace(0, M - 1, n_s, device=dev).long()
            A_s = A[idx].cpu().numpy()
            w8c, sbc, fwc, indc = W8.cpu().numpy(), sb.cpu().numpy(), fw.cpu().numpy(), ind.cpu().numpy()
            t0 = time.perf_counter()
            want = O.forward(A_s, w8c, sbc, fwc, indc)
            cpu_s = time.perf_counter() - t0
            got = out[idx].cpu().numpy()
            d = got.astype(np.float64) - want.astype(np.float64)
            r["cpu_port_ms_per_token"] = round(cpu_s * 1e3 / n_s, 3)
            r["cpu_cores"] = O.num_threads()
            r["cpu_sample_tokens"] = n_s
            r["max_abs_err_vs_oracle"] = float(np.abs(d).max())
            r["rel_frobenius_vs_oracle"] = float(np.linalg.norm(d) / max(np.linalg.norm(want.astype(np.float64)), 1e-30))
            r["bit_identical_frac_vs_oracle"] = float((got.view(np.uint16) == want.view(np.uint16)).mean())
            rows.append(r)
            print(json.dumps(r), flush=True)
            del A, out, ws
        del sets
        torch.cuda.empty_cache()
    merged = []
    for f in args.merge:
        for ln in Path(f).read_text().splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                for name, v in d.get("roofline", {}).get("per_linear", {}).items():
                    merged.append(dict(workload=d["config"]["workload"], gpus=d["n_gpus"], linear=name, N_per_rank=v["N"], K_per_rank=v["K"],
                                       M=d["config"]["tokens_per_step"], gemm_us=v["gemm_us"], quant_us=v["quant_us"], tflops_per_gpu=v["gemm_tflops"],
                                       step_ms=d["ms_per_step"], whole_step_tflops=d["value"], tokens_per_s=d.get("tokens_per_s")))
    result = dict(int8_peak_measured_tflops=round(int8_peak, 1), hbm_peak_gbs=pk["hbm"], rows=rows, multi_gpu=merged)
    Path(args.out + ".json").write_text(json.dumps(result, indent=1))
    md = ["# BASELINE.md 3.3 result table (one B200; INT8 peak measured in this run: %.0f TOPS burst, cuBLASLt 8192^3)" % int8_peak, "",
          "| model | linear (N x K) | M | ours us | TFLOPS | of measured INT8 peak | of 4.5 POPS | HBM GB/s (of measured) | reference kernels us | speed-up | "
          "CPU port ms/token (cores) | max-abs / rel-Frobenius vs oracle |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        md.append("| %s | %s (%d x %d) | %d | %.1f | %.0f | %.3f | %.3f | %.0f (%.2f) | %s | %s | %.2f (%d) | %.2e / %.1e |" % (
            r["model"], r["linear"], r["N"], r["K"], r["M"], r["ours_us"], r["tflops"], r["frac_of_measured_int8_peak"], r["frac_of_spec_4500"],
            r["hbm_gbs"], r["frac_of_measured_hbm"], r.get("ref_kernels_us", "-"), r.get("speedup_vs_ref_kernels", "-"),
            r["cpu_port_ms_per_token"], r["cpu_cores"], r["max_abs_err_vs_oracle"], r["rel_frobenius_vs_oracle"]))
    if merged:
        md += ["", "## tensor-parallel runs (bench.py --gpus N; per-rank shard shapes)", "",
               "| workload | GPUs | linear | N x K per rank | M | GEMM us | quantise us | TFLOPS per GPU | step ms | whole-step TFLOPS | tokens/s |",
               "|---|---|---|---|---|---|---|---|---|---|---|"]
        for m in merged:
            md.append("| %s | %d | %s | %d x %d | %d | %.1f | %.1f | %.0f | %.3f | %.0f | %s |" % (
                m["workload"], m["gpus"], m["linear"], m["N_per_rank"], m["K_per_rank"], m["M"], m["gemm_us"], m["quant_us"], m["tflops_per_gpu"],
                m["step_ms"], m["whole_step_tflops"], ("%.0f" % m["tokens_per_s"]) if m.get("tokens_per_s") else "-"))
    Path(args.out + ".md").write_text("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
