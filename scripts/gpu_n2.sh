mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc $?"; cat gpurun_out/bench_n2.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/bench_train.py --steps 3 --warmup 2 > gpurun_out/train_bench_n2.json 2> gpurun_out/train_bench_n2.err
echo "train n2 rc $?"; cat gpurun_out/train_bench_n2.json; tail -3 gpurun_out/train_bench_n2.err
