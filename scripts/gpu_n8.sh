mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 3 --no-vae > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N rc $?"; grep "^{" gpurun_out/bench_n$N.json | tail -1 | cut -c1-300
