mkdir -p gpurun_out/s3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s3/t_final.log; cat gpurun_out/s3/t_final.log
timeout 600 python bench.py > gpurun_out/s3/bench_final2.json 2> gpurun_out/s3/bench_final2.err; tail -2 gpurun_out/s3/bench_final2.err
python -c "
import json
d=json.load(open('gpurun_out/s3/bench_final2.json'))
r=d['roofline']
print(round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'tok/s', round(d['tokens_per_s']), 'frac', r['frac'], 'step_frac', r['step_frac_of_floor'], 'parity', d['parity_checked'], 'e2e', d['e2e']['value'], 'ref_gpu x', d['ref_gpu']['speedup_ours'], d['clocks'])
"
python -c "import __graft_entry__ as g; g.smoke()"
