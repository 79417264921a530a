mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "split or gemm or conv" 2>&1 | tail -8
timeout 900 python scripts/bench_splitk.py 2>&1 | tee gpurun_out/splitk.log | tail -30
