mkdir -p gpurun_out/r2
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/gpu_tp_fused.py 512x4096x4096 512x4096x11008 512x8192x8192 32x4096x4096 > gpurun_out/r2/tp${N}_tool.log 2>&1
grep "^{\|PASS\|FAIL\|Error\|error" gpurun_out/r2/tp${N}_tool.log | python -c "
import sys, json
for ln in sys.stdin:
    ln=ln.strip()
    if ln.startswith('{'):
        d=json.loads(ln)
        if 'shape' in d: print(d['shape'], 'mism', d['mismatches'], d['default_path_mismatches'], 'gemm', round(d['gemm_only_us'],1), 'bulk', round(d['fused_us'],1), 'default', round(d['default_path_us'],1), 'gemm+nccl', round(d['unfused_us'],1))
        else: print(d)
    else: print(ln[:300])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-e2e 2> gpurun_out/r2/tp${N}_bs512b.err | grep "^{" > gpurun_out/r2/tp${N}_bs512b.json
python -c "
import json
d=json.load(open('gpurun_out/r2/tp${N}_bs512b.json'))
print('bs512 N=$N', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'parity', d['parity_checked'], d['parity']['linears'])
"
