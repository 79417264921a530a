mkdir -p gpurun_out
for v in ${1:-24 6168}; do
timeout 300 python scripts/attn_tile_trace.py 40 $v > gpurun_out/attn_tile_trace40_$v.log 2>&1; echo rc=$?
cat gpurun_out/attn_tile_trace40_$v.log
done
