mkdir -p gpurun_out
timeout 1500 python bench.py --steps 2 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','gpu_launches')}); print('e2e',d['e2e']['value'])"
head -${1:-45} gpurun_out/bench.err | cut -c1-130
