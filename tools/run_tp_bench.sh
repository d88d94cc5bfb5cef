# usage: bash tools/run_tp_bench.sh N   (inside gpurun --gpus N)
N=$1
mkdir -p gpurun_out/r2
run() { # workload tag extra
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $1 $3 2> gpurun_out/r2/tp${N}_$2.err | grep "^{" > gpurun_out/r2/tp${N}_$2.json
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2/tp${N}_$2.json"))
    print("$2", "N=$N", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "tok/s", round(d["tokens_per_s"]), "parity", d["parity_checked"], "e2e ms", (d.get("e2e") or {}).get("ms_per_step"), d["run_details"].get("fused_allreduce_unavailable"))
except Exception as e:
    print("$2 FAILED", e); print(open("gpurun_out/r2/tp${N}_$2.err").read()[-1500:])
P
}
run llama2-7b-linears-decode-bs512 bs512 ""
run llama2-7b-linears-bs32xseq2048 prefill ""
run llama2-70b-linears-decode-bs512 70b_bs512 "--no-e2e"
