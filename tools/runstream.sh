# trim / mask device times: default engines, everything on the warp engine, trim with the gather pass
for cfg in "SK_WARP_STREAM=0" "" "SK_WARP_STREAM=1" "SK_TRIM_GATHER=1"; do
  env $cfg python bench.py --skip-e2e --skip-cpu --steps 5 > gpurun_out/v.json 2>gpurun_out/v.err
  python -c "
import json;d=json.load(open('gpurun_out/v.json'));r=d['roofline'];o=r['other_ops'];print('[$cfg]', r['ms_per_launch'], r['other_kernel']['ms_per_launch'], 'trim', o['trim_by_quality']['ms_per_launch'], o['trim_by_quality']['engine_bits'], 'mask', o['mask_by_quality']['ms_per_launch'], o['mask_by_quality']['engine_bits'])"
done
