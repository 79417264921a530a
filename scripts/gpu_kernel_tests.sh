mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
for k in "gemm" "conv3x3" "attention" "groupnorm or layernorm" "conv_in or time_embedding or casts or cfg"; do
  echo "=== $k" >> gpurun_out/kernels.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$k" --timeout 180 -p no:cacheprovider >> gpurun_out/kernels.log 2>&1
  echo "exit $?" >> gpurun_out/kernels.log
done
tail -5 gpurun_out/kernels.log
