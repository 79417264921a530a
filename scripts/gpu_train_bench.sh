mkdir -p gpurun_out
timeout 900 python scripts/bench_train.py --steps 3 --warmup 2 > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err
echo "rc $?"; cat gpurun_out/train_bench.json; tail -5 gpurun_out/train_bench.err
