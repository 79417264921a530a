mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q --timeout 600 -p no:cacheprovider -k "${1:-attention_backward}" > gpurun_out/train_tests.log 2>&1; echo "rc $?"
tail -25 gpurun_out/train_tests.log
timeout 900 python scripts/bench_train.py --steps 3 --warmup 2 --breakdown > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; echo "rc $?"
cat gpurun_out/train_bench.json; head -${2:-30} gpurun_out/train_bench.err | cut -c1-150
