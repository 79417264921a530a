mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" --timeout 300 -p no:cacheprovider 2>&1 | tail -3
python scripts/bench_attn.py 0 10 ${ATTN_VARIANTS:-2,8,12} 2>&1 | tee gpurun_out/attn_bench.log
