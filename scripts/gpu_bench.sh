mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc $?"; cat gpurun_out/bench.json; tail -20 gpurun_out/bench.err
