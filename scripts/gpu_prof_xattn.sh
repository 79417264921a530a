mkdir -p gpurun_out
python scripts/xattn_one.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:xattn -s 2 -c 1 -o gpurun_out/xattn_prof -f python scripts/xattn_one.py > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
