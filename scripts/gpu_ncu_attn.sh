# ncu --set full of the attention tile kernel (d = 40, N = 4096, B = 16), only after the plain command exited 0
mkdir -p gpurun_out
python scripts/attn_one.py > gpurun_out/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tile -s 2 -c 1 -o gpurun_out/attn_r02 -f python scripts/attn_one.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc $?"; cat gpurun_out/plain_attn.log; tail -2 gpurun_out/ncu_attn.log
