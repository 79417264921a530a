mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/unet.log 2>&1
echo "exit $?" >> gpurun_out/unet.log
tail -40 gpurun_out/unet.log
