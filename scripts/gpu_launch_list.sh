# launch list + DRAM bytes of one UNet step (ncu pass only after the plain command exited 0)
mkdir -p gpurun_out
python scripts/unet_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:gemm_tc|attention|gn_|layernorm|conv_in|conv_out|linear_small|timestep_emb|cast_bf16|upsample2x|nhwc_to|cfg_ddim|advance_step|xattn" --csv --log-file gpurun_out/launches.csv python scripts/unet_step.py 2 > gpurun_out/ncu.log 2>&1
echo "ncu step rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
