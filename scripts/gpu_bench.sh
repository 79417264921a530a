mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -s --timeout 600 -p no:cacheprovider -k "ddim or guidance" 2>&1 | grep -n "rel-L2\|passed\|failed\|FAILED\|Error" | head
timeout 1500 python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','gpu_launches','clocks')}); print('e2e',d['e2e']); r=d['roofline']; print({k:r[k] for k in r if k not in ('how','peak_source')})"
tail -14 gpurun_out/bench.err
