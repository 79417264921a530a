mkdir -p gpurun_out
python scripts/bench_gemm.py ${1:-res} 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/gemm_prof -f python scripts/bench_gemm.py ${1:-res} 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
