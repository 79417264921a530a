mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q --timeout 600 -p no:cacheprovider -k "trainer or distill or conditioning or multi_step" 2>&1 | tail -15
timeout 600 python scripts/bench_train.py --steps 5 --warmup 3 2> gpurun_out/train_step.err | tee gpurun_out/train_step.json; tail -3 gpurun_out/train_step.err
timeout 600 python scripts/bench_train.py --steps 5 --warmup 3 --unet-graph 2>/dev/null
