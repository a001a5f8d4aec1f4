for v in "" seqkit_b200/variants/lib_t8.so seqkit_b200/variants/lib_su.so seqkit_b200/variants/lib_su_t8.so; do
  SK_LIB=$v python bench.py --skip-e2e --skip-cpu --steps 10 > gpurun_out/v.json 2>gpurun_out/v.err
  python -c "
import json;d=json.load(open('gpurun_out/v.json'));r=d['roofline'];print('$v',d['value'],r['frac'],r['ms_per_launch'],r['other_kernel']['ms_per_launch'])"
done
