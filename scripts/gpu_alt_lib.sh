# A/B of alternative builds of the library (adaprompt_b200/_alt/*.so, git-ignored): each is swapped in for one run of the
# attention micro-check
mkdir -p gpurun_out
echo "--- base"; timeout 600 python scripts/attn_tile_check.py 10 2>&1 | tail -7
cp adaprompt_b200/libadaface_b200.so /tmp/base.so
for alt in adaprompt_b200/_alt/*.so; do
  echo "--- $alt"
  cp $alt adaprompt_b200/libadaface_b200.so
  timeout 600 python scripts/attn_tile_check.py 10 2>&1 | tail -7 | head -5
done
cp /tmp/base.so adaprompt_b200/libadaface_b200.so
