# full GPU test suite + the sampling / training bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'unet_step_ms', 'gpu_launches')}, 'e2e', d['e2e']['value'])
print('parity', d['parity']['ok'], d['parity']['eps_rel_l2'], d['parity']['final_latent_rel_l2_image0'])
print('train', {k: d['train_step'][k] for k in ('value', 'ms_per_optimizer_step')})
r = d['roofline']
print('roofline', r['kernel'], round(r['achieved'], 1), round(r['frac'], 3), '| gn', round(r['groupnorm_silu']['frac'], 3), '| attn', round(r['attention']['achieved'], 1), round(r['attention']['frac'], 3))
PY
