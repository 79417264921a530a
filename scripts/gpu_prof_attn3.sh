mkdir -p gpurun_out
V=${1:-64}
python scripts/attn_one.py $V > gpurun_out/plain_attn_$V.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_ -s 2 -c 1 -o gpurun_out/attn_r02_$V -f python scripts/attn_one.py $V > gpurun_out/ncu_attn_$V.log 2>&1
echo "ncu rc $?"; cat gpurun_out/plain_attn_$V.log; tail -2 gpurun_out/ncu_attn_$V.log
