# bench.py on N GPUs of one box (torchrun, NCCL): sampling (no collective) + the Stage-1 training leg (gradient all-reduce)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-vae > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc $?"
python - <<PY
import json
d = json.load(open('gpurun_out/bench_n$N.json'))
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'unet_step_ms')}, 'e2e', d['e2e']['value'])
print('train', {k: d['train_step'][k] for k in ('value', 'ms_per_optimizer_step', 'allreduce_ms', 'allreduce_bytes_per_step')})
print('parity', d['parity']['ok'], d['parity']['eps_rel_l2'])
PY
tail -3 gpurun_out/bench_n$N.err
