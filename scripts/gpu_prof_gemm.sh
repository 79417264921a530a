mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gemm or conv3x3" --timeout 180 -p no:cacheprovider 2>&1 | tail -3
python scripts/bench_gemm.py all > gpurun_out/gemm_bench.log 2>&1 ; cat gpurun_out/gemm_bench.log
python scripts/bench_gemm.py geglu 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/gemm_prof -f python scripts/bench_gemm.py geglu 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu.log
