mkdir -p gpurun_out
python scripts/bench_gemm.py all 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 33 -o gpurun_out/gemm_all -f python scripts/bench_gemm.py all 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
