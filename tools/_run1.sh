mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm_dequant or enqueue_matches or benchmarked" 2>&1 | tail -4 > gpurun_out/r2/t13.log; cat gpurun_out/r2/t13.log
for shp in "32 4096 4096" "32 4096 11008" "32 12288 4096" "64 4096 4096" "128 4096 11008" "32 8192 1024"; do python tests/gpu_ab.py "1,3,15,0" $shp 2>&1 | tail -5; done
python bench.py --workload llama2-7b-linears-decode-bs32 --no-e2e --no-cpu > gpurun_out/r2/bench_bs32c.json 2> gpurun_out/r2/bench_bs32c.err; tail -2 gpurun_out/r2/bench_bs32c.err
python -c "
import json
d=json.load(open('gpurun_out/r2/bench_bs32c.json'))
print(d['value'], d['ms_per_layer'], d['tokens_per_s'], d['roofline']['frac'], d['roofline']['bound'], d['parity_checked'], d['ref_gpu']['speedup_ours'])
for k,v in d['roofline']['per_linear'].items(): print(k, v['gemm_us'], v['quant_us'], v['floor_us'])
"
