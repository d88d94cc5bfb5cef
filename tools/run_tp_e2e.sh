# usage: bash tools/run_tp_e2e.sh N   (inside gpurun --gpus N): the metric's configuration at N ranks, e2e leg included
N=$1
mkdir -p gpurun_out/s3
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N 2> gpurun_out/s3/tp${N}_bs512.err | grep "^{" > gpurun_out/s3/tp${N}_bs512.json
python - <<P
import json
try:
    d=json.load(open("gpurun_out/s3/tp${N}_bs512.json"))
    print("N=$N", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "tok/s", round(d["tokens_per_s"]), "parity", d["parity_checked"], "e2e", d.get("e2e"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/s3/tp${N}_bs512.err").read()[-2500:])
P
