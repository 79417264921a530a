# round-2 evidence: (1) launch list + DRAM bytes of one UNet step, (2) ncu --set full of the attention tile kernel (d = 40)
# and the flash backward kernels.  Each ncu command only after the same plain command exited 0.
mkdir -p gpurun_out
python scripts/unet_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:gemm_tc|attention|gn_|layernorm|conv_in|conv_out|linear_small|timestep_emb|cast_bf16|upsample2x|nhwc_to|cfg_ddim|advance_step|xattn" --csv --log-file gpurun_out/launches.csv python scripts/unet_step.py 2 > gpurun_out/ncu.log 2>&1
echo "ncu step rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
python scripts/attn_one.py > gpurun_out/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tile -s 2 -c 1 -o gpurun_out/attn_r02 -f python scripts/attn_one.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc $?"; cat gpurun_out/plain_attn.log; tail -2 gpurun_out/ncu_attn.log
