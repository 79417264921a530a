mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" --timeout 180 -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "exit $?" >> gpurun_out/kernels.log
tail -5 gpurun_out/kernels.log
timeout 300 python scripts/bench_attn.py 0 10 2>&1 | tee gpurun_out/attn_bench.log
