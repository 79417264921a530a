# quick iteration: kernel + unet tests, gemm microbench, bench breakdown
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "kernels rc $?"; tail -15 gpurun_out/kernels.log | cut -c1-200
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -s --timeout 600 -p no:cacheprovider > gpurun_out/unet.log 2>&1
echo "unet rc $?"; grep -n "rel-L2\|passed\|failed\|Error" gpurun_out/unet.log | cut -c1-250 | head -30
python scripts/bench_gemm.py all > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
bash scripts/gpu_bench.sh ${1:-60}
