mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv or geglu or pair" --timeout 120 -p no:cacheprovider 2>&1 | tail -15 | cut -c1-200
python scripts/bench_gemm.py all 2>&1 | tee gpurun_out/gemm_bench.log
