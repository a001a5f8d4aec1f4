# trim / mask device times with the warp engine switched on for them (SK_WARP_STREAM=1), per library variant
for ws in 0 1; do for v in "" $(ls seqkit_b200/variants/*.so 2>/dev/null); do
  SK_WARP_STREAM=$ws SK_LIB=$v python bench.py --skip-e2e --skip-cpu --steps 5 > gpurun_out/v.json 2>gpurun_out/v.err
  python -c "
import json;d=json.load(open('gpurun_out/v.json'));r=d['roofline'];o=r['other_ops'];print('ws=$ws $v', r['ms_per_launch'], r['other_kernel']['ms_per_launch'], 'trim', o['trim_by_quality']['ms_per_launch'], o['trim_by_quality']['engine_bits'], 'mask', o['mask_by_quality']['ms_per_launch'], o['mask_by_quality']['engine_bits'])"
done
done
