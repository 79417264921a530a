# one `ncu --set full` capture per hot kernel (after the plain run exited 0), reports under gpurun_out/
mkdir -p gpurun_out
python scripts/bench_gemm.py conv64 1 > gpurun_out/plain_conv64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/full_conv64 -f python scripts/bench_gemm.py conv64 1 > gpurun_out/ncu_conv64.log 2>&1
echo "conv64 rc $?"
python scripts/bench_gemm.py res 1 > gpurun_out/plain_res.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/full_gemm_res -f python scripts/bench_gemm.py res 1 > gpurun_out/ncu_res.log 2>&1
echo "res rc $?"
python scripts/attn_one.py > gpurun_out/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:attention_rowsplit|attention_pair" -s 2 -c 1 -o gpurun_out/full_attn -f python scripts/attn_one.py > gpurun_out/ncu_attn.log 2>&1
echo "attn rc $?"
python scripts/bench_norm.py 1 > gpurun_out/plain_norm.log 2>&1 &&
ncu --set full --clock-control none -k regex:gn_apply_kernel -s 3 -c 1 -o gpurun_out/full_gn_apply -f python scripts/bench_norm.py 1 > gpurun_out/ncu_norm.log 2>&1
echo "gn rc $?"
ls -la gpurun_out/full_*.ncu-rep
