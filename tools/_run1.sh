mkdir -p gpurun_out/r2
timeout 300 python bench.py --workload llama2-7b-linears-decode-bs32 --no-e2e --no-cpu > gpurun_out/r2/bench_bs32e.json 2> gpurun_out/r2/bench_bs32e.err; tail -2 gpurun_out/r2/bench_bs32e.err
python -c "
import json
d=json.load(open('gpurun_out/r2/bench_bs32e.json'))
print(d['value'], d['ms_per_layer'], d['tokens_per_s'], d['roofline']['frac'], d['roofline']['bound'], d['parity_checked'], d['ref_gpu']['speedup_ours'])
for k,v in d['roofline']['per_linear'].items(): print(k, v['gemm_us'], v['quant_us'], v['floor_us'])
"
timeout 300 python bench.py --workload llama2-70b-linears-decode-bs32 --no-e2e --no-cpu > gpurun_out/r2/bench_70b_bs32.json 2> gpurun_out/r2/bench_70b_bs32.err; tail -2 gpurun_out/r2/bench_70b_bs32.err
python -c "
import json
d=json.load(open('gpurun_out/r2/bench_70b_bs32.json'))
print(d['value'], d['ms_per_layer'], d['tokens_per_s'], d['roofline']['frac'], d['roofline']['bound'], d['parity_checked'], d['ref_gpu']['speedup_ours'])
for k,v in d['roofline']['per_linear'].items(): print(k, v['gemm_us'], v['quant_us'], v['floor_us'])
"
