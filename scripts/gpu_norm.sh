# GroupNorm / LayerNorm micro-benchmark + the kernel tests that cover them
mkdir -p gpurun_out
python scripts/bench_norm.py 20 > gpurun_out/norm.log 2>&1; echo "rc $?"; cat gpurun_out/norm.log
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "groupnorm or gn_ or stats" 2>&1 | tail -5
