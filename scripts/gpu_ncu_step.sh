mkdir -p gpurun_out
python scripts/unet_step.py 1 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gemm_tc|attention|gn_|layernorm|conv_in|conv_out|linear_small|timestep_emb|cast_bf16|upsample2x|nhwc_to|cfg_ddim|advance_step" --csv --log-file gpurun_out/launches.csv python scripts/unet_step.py 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
