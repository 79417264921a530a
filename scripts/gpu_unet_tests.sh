mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -q -s --timeout 600 -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/unet.log 2>&1
echo "exit $?" >> gpurun_out/unet.log
grep -n "rel-L2\|passed\|failed\|FAILED\|Error\|exit" gpurun_out/unet.log | head -40
