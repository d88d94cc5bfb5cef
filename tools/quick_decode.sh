# quick device-only decode numbers (bs 512 and bs 32) + a parity subset; usage: bash tools/quick_decode.sh TAG
T=${1:-x}
mkdir -p gpurun_out/s3
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "enqueue or gated or full_size or benchmarked" 2>&1 | tail -2
for W in llama2-7b-linears-decode-bs512 llama2-7b-linears-decode-bs32; do
  timeout 300 python bench.py --workload $W --no-e2e --no-cpu --no-ref-gpu > gpurun_out/s3/q_${T}_$W.json 2> gpurun_out/s3/q_${T}_$W.err || tail -3 gpurun_out/s3/q_${T}_$W.err
  python - <<P
import json
d = json.load(open("gpurun_out/s3/q_${T}_$W.json"))
r = d["roofline"]
print("$W", round(d["value"], 1), "TFLOPS", round(d["ms_per_step"], 4), "ms/step", round(d.get("ms_per_layer", 0) * 1e3, 2), "us/layer", "parity", d.get("parity_checked"),
      {k: (v["gemm_us"], v["quant_us"]) for k, v in r.get("per_linear", {}).items()})
P
done
