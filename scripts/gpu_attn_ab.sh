mkdir -p gpurun_out
timeout 900 python scripts/attn_tile_check.py 10 > gpurun_out/attn_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/attn_ab.log | tail -30
