mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2/t9.log
cat gpurun_out/r2/t9.log
python bench.py > gpurun_out/r2/bench_pdl.json 2> gpurun_out/r2/bench_pdl.err
tail -3 gpurun_out/r2/bench_pdl.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2/bench_pdl.json'))
print(d['value'], d['ms_per_layer'], d['roofline']['step_frac_of_floor'], d['parity_checked'])
for k,v in d['roofline']['per_linear'].items(): print(k, v['gemm_us'], v['quant_us'])
P
