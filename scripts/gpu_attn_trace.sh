mkdir -p gpurun_out
for d in 40 80; do
timeout 300 python scripts/attn_tile_trace.py $d > gpurun_out/attn_tile_trace$d.log 2>&1; echo rc=$?
cat gpurun_out/attn_tile_trace$d.log
done
