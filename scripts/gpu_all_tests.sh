mkdir -p gpurun_out
rm -f gpurun_out/kernels.log
for k in "attention"; do
  echo "=== $k" >> gpurun_out/kernels.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$k" --timeout 180 -p no:cacheprovider >> gpurun_out/kernels.log 2>&1
  echo "exit $?" >> gpurun_out/kernels.log
done
tail -5 gpurun_out/kernels.log
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/unet.log 2>&1
echo "exit $?" >> gpurun_out/unet.log
grep -n "rel-L2\|passed\|failed\|FAILED\|res_out5\|Error" gpurun_out/unet.log | head -40
