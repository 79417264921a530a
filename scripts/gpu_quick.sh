# sampler / UNet / kernel tests + a short sampling bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py tests/test_kernels_gpu.py tests/test_vae_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-vae --no-train --no-parity > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_quick.json'))
print('images/s', round(d['value'], 3), 'unet_step_ms', round(d['unet_step_ms'], 3), 'e2e', round(d['e2e']['value'], 3), 'launches', d['gpu_launches'])
PY
