# full GPU check: tests, bench, ncu launch list of one UNet step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1
echo "tests rc $?" | tee -a gpurun_out/tests_gpu.log
tail -3 gpurun_out/tests_gpu.log
timeout 1500 python bench.py --steps 3 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','unet_step_ms','gpu_launches')}); print('e2e',d['e2e']['value'], d['cpu_baseline'])"
head -70 gpurun_out/bench.err | cut -c1-130
python scripts/unet_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:gemm_tc|attention|gn_|layernorm|conv_in|conv_out|linear_small|timestep_emb|cast_bf16|upsample2x|nhwc_to|cfg_ddim|advance_step|xattn" --csv --log-file gpurun_out/launches.csv python scripts/unet_step.py 2 > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
