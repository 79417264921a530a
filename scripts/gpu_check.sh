# full GPU check without ncu: attention micro-check, all GPU tests, bench with breakdown
mkdir -p gpurun_out
bash scripts/gpu_attn_ab.sh
( time timeout 1700 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1 ) 2>&1 | grep real
echo "tests rc $?" | tee -a gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
bash scripts/gpu_bench.sh 60
