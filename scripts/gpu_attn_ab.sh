mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_ab.txt 2>&1
timeout 120 scripts/ubench/mufu2 > gpurun_out/mufu2.log 2>&1; echo "mufu2 rc=$?"
cat gpurun_out/mufu2.log
timeout 900 python scripts/attn_tile_check.py ${1:-8,16,17,18,19,20,21,22,23,48,49} 10 > gpurun_out/attn_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/attn_ab.log | tail -40
