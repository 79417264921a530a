mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" --timeout 180 -p no:cacheprovider 2>&1 | tail -3
python scripts/bench_attn.py > gpurun_out/attn_bench.log 2>&1 && cat gpurun_out/attn_bench.log
python scripts/bench_attn.py 40 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o gpurun_out/attn_prof -f python scripts/bench_attn.py 40 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu.log
