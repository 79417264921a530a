# GroupNorm / LayerNorm micro-benchmark + the kernel tests that cover them
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
python scripts/bench_norm.py 20 > gpurun_out/norm.log 2>&1; echo "rc $?"; cat gpurun_out/norm.log
