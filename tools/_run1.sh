mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2/t10.log
cat gpurun_out/r2/t10.log
