# the non-default workloads, one JSON line each; usage: bash tools/run_workloads.sh TAG
T=${1:-s3}
mkdir -p gpurun_out/$T
for W in llama2-7b-linears-bs32xseq2048 qwen2-7b-linears-bs32xseq2048 llama2-70b-linears-decode-bs512 llama2-7b-linears-decode-bs32; do
  timeout 400 python bench.py --workload $W --no-cpu > gpurun_out/$T/bench_$W.json 2> gpurun_out/$T/bench_$W.err || tail -3 gpurun_out/$T/bench_$W.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/$T/bench_$W.json"))
    r = d["roofline"]
    print("$W", round(d["value"], 1), "TFLOPS", round(d["ms_per_step"], 3), "ms/step frac", r["frac"], "step_frac", r.get("step_frac_of_floor"), "parity", d.get("parity_checked"),
          "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("results_checked"), "ref_gpu x", (d.get("ref_gpu") or {}).get("speedup_ours"), "traffic", r.get("traffic"))
except Exception as e:
    print("$W FAILED", e)
P
done
