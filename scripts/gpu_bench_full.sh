mkdir -p gpurun_out
( time timeout 1700 python bench.py --steps 3 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real
echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'unet_step_ms', 'gpu_launches')})
print('e2e', d['e2e']['value'])
print('parity', d.get('parity'))
print('eager', d.get('gpu_eager_baseline'))
print('train', d.get('train_step'))
print('roofline gn', d['roofline'].get('groupnorm_silu'))
print('roofline attn', d['roofline'].get('attention'))
print('cpu', d.get('cpu_baseline'))
PY
head -${1:-20} gpurun_out/bench.err | cut -c1-130
