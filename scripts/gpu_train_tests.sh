mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -x -s --timeout 600 -p no:cacheprovider > gpurun_out/train_tests.log 2>&1
echo "exit $?" >> gpurun_out/train_tests.log
grep -n "passed\|failed\|FAILED\|Error\|error\|rel-L2\|assert" gpurun_out/train_tests.log | head -40
tail -5 gpurun_out/train_tests.log
