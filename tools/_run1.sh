mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2/t15.log; cat gpurun_out/r2/t15.log
timeout 600 python bench.py > gpurun_out/r2/bench_final.json 2> gpurun_out/r2/bench_final.err; tail -2 gpurun_out/r2/bench_final.err
timeout 600 python bench.py --impl reference > gpurun_out/r2/bench_final_ref.json 2> gpurun_out/r2/bench_final_ref.err; tail -2 gpurun_out/r2/bench_final_ref.err; cat gpurun_out/r2/bench_final_ref.json | cut -c1-600
timeout 600 python bench.py --workload llama2-7b-linears-bs32xseq2048 > gpurun_out/r2/bench_prefill.json 2> gpurun_out/r2/bench_prefill.err; tail -2 gpurun_out/r2/bench_prefill.err
timeout 600 python bench.py --workload qwen2-7b-linears-bs32xseq2048 --no-e2e > gpurun_out/r2/bench_qwen.json 2> gpurun_out/r2/bench_qwen.err; tail -2 gpurun_out/r2/bench_qwen.err
timeout 600 python bench.py --workload llama2-70b-linears-decode-bs512 --no-e2e > gpurun_out/r2/bench_70b_bs512.json 2> gpurun_out/r2/bench_70b_bs512.err; tail -2 gpurun_out/r2/bench_70b_bs512.err
python - <<'P'
import json
for f in ("bench_final","bench_prefill","bench_qwen","bench_70b_bs512"):
    d=json.load(open(f"gpurun_out/r2/{f}.json"))
    r=d["roofline"]
    print(f, round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "tok/s", round(d["tokens_per_s"]), "frac", r["frac"], "step_frac", r["step_frac_of_floor"], "parity", d["parity_checked"], "e2e", (d.get("e2e") or {}).get("value"), "ref_gpu x", (d.get("ref_gpu") or {}).get("speedup_ours"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
P
