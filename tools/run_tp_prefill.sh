# usage: bash tools/run_tp_prefill.sh N   (inside gpurun --gpus N): configs[1] (M = 65536) at N ranks, e2e leg included
N=$1
mkdir -p gpurun_out/s3
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload llama2-7b-linears-bs32xseq2048 2> gpurun_out/s3/tp${N}_prefill.err | grep "^{" > gpurun_out/s3/tp${N}_prefill.json
python - <<P
import json
try:
    d=json.load(open("gpurun_out/s3/tp${N}_prefill.json"))
    print("N=$N", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "parity", d["parity_checked"], "e2e", {k: v for k, v in (d.get("e2e") or {}).items() if k != "path"})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/s3/tp${N}_prefill.err").read()[-2500:])
P
