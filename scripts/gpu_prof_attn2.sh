mkdir -p gpurun_out
python scripts/attn_one.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_pair -s 2 -c 1 -o gpurun_out/attn_prof -f python scripts/attn_one.py > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
