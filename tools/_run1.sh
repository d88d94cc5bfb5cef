mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "virtual_ranks" 2>&1 | tail -12 > gpurun_out/r2/t17.log; cat gpurun_out/r2/t17.log
