mkdir -p gpurun_out/prof2
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:fat -s 2 -c 1 -f -o gpurun_out/prof2/r2_fat_512x12288x4096 python tests/gpu_profile_target.py 0 512 12288 4096 4 > gpurun_out/prof2/p1.log 2>&1
$NCU -k regex:fat -s 2 -c 1 -f -o gpurun_out/prof2/r2_gated_512x11008x4096 python tests/gpu_profile_gated.py 512 11008 4096 4 > gpurun_out/prof2/p2.log 2>&1
$NCU -k regex:gemm_dequant -s 2 -c 1 -f -o gpurun_out/prof2/r2_cfg5_512x4096x11008 python tests/gpu_profile_target.py 0 512 4096 11008 4 > gpurun_out/prof2/p3.log 2>&1
$NCU -k regex:quant_extract -s 2 -c 1 -f -o gpurun_out/prof2/r2_quant_512x4096 python tests/gpu_profile_target.py 0 512 12288 4096 4 > gpurun_out/prof2/p4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/prof2/r2_bench_launches.csv python bench.py --layers 4 --steps 3 --warmup 1 --no-e2e --no-cpu --no-ref-gpu --no-parity --graph off > gpurun_out/prof2/bench_under_ncu.log 2>&1
tail -2 gpurun_out/prof2/p*.log; ls -la gpurun_out/prof2
