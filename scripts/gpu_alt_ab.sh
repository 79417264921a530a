# A/B of alternative builds (adaprompt_b200/_alt/*.so, scripts/build_alt.sh) on the sampling bench: graph UNet step time
mkdir -p gpurun_out
run() {
  timeout 500 python bench.py --steps 3 --warmup 3 --no-vae --no-train --no-parity > gpurun_out/pdl_$1.json 2> gpurun_out/pdl_$1.err
  echo "$1 rc $?"
  python - <<PY
import json
d = json.load(open('gpurun_out/pdl_$1.json'))
print('$1', 'images/s', round(d['value'], 3), 'unet_step_ms', round(d['unet_step_ms'], 3), 'e2e', round(d['e2e']['value'], 3), 'eager sum', round(d['roofline']['unet_step_ms_eager_sum'], 3))
PY
}
cp adaprompt_b200/libadaface_b200.so /tmp/base.so
run base
for alt in adaprompt_b200/_alt/*.so; do
  cp $alt adaprompt_b200/libadaface_b200.so
  run $(basename $alt .so)
done
cp /tmp/base.so adaprompt_b200/libadaface_b200.so
