mkdir -p gpurun_out
for c in "$@"; do
python scripts/bench_gemm.py $c 1 > gpurun_out/plain_$c.log 2>&1 &&
ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/gemm_$c -f python scripts/bench_gemm.py $c 1 > gpurun_out/ncu_$c.log 2>&1
echo "ncu $c rc $?"; cat gpurun_out/plain_$c.log
done
