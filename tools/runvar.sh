# A/B of library variants built with different -D switches (SK_LIB selects the .so)
for v in "" $(ls seqkit_b200/variants/*.so 2>/dev/null); do
  for rep in 1 2; do
  SK_LIB=$v python bench.py --skip-e2e --skip-cpu --steps 10 > gpurun_out/v.json 2>gpurun_out/v.err
  python -c "
import json;d=json.load(open('gpurun_out/v.json'));r=d['roofline'];print('$v',d['value'],r['frac'],r['ms_per_launch'],r['other_kernel']['ms_per_launch'])"
  done
done
