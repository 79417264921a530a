mkdir -p gpurun_out
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 2 --warmup 3 --no-vae > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N rc $?"; grep "^{" gpurun_out/bench_n$N.json | tail -1 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 scripts/bench_train.py --steps 3 --warmup 2 > gpurun_out/train_bench_n$N.json 2> gpurun_out/train_bench_n$N.err
echo "train n$N rc $?"; grep "^{" gpurun_out/train_bench_n$N.json | tail -1 | cut -c1-300
