mkdir -p gpurun_out
timeout 900 python scripts/attn_tile_check.py $1 10 > gpurun_out/attn_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/attn_ab.log | tail -30
for v in $2; do
timeout 300 python scripts/attn_tile_trace.py 40 $v > gpurun_out/attn_tile_trace40_$v.log 2>&1; echo rc=$?
cat gpurun_out/attn_tile_trace40_$v.log
done
