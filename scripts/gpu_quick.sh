mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "groupnorm or layernorm" --timeout 180 -p no:cacheprovider 2>&1 | tail -3
bash scripts/gpu_bench.sh 14
grep groupnorm gpurun_out/bench.err | head -12
