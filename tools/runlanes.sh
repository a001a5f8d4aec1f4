# bench-only sweep of the warp engine's tile size (SK_TILE_LANES)
for v in 27 28 29 30; do
  SK_TILE_LANES=$v python bench.py --skip-e2e --skip-cpu --steps 10 > gpurun_out/v.json 2>gpurun_out/v.err
  python -c "
import json;d=json.load(open('gpurun_out/v.json'));r=d['roofline'];print('lanes $v',d['value'],r['frac'],r['ms_per_launch'],r['other_kernel']['ms_per_launch'])"
done
