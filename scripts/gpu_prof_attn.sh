mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" --timeout 180 -p no:cacheprovider 2>&1 | tail -3
python scripts/bench_attn.py > gpurun_out/attn_bench.log 2>&1 ; cat gpurun_out/attn_bench.log
