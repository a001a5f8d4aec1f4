set -x
timeout 120 python __graft_entry__.py smoke > gpurun_out/rm_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/rm_bench.json 2> gpurun_out/rm_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/rm_bench_ref.json 2> gpurun_out/rm_bench_ref.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sk_ -c 400 --csv --log-file gpurun_out/rm_launches.csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --cli-pairs 0 --skip-configs > gpurun_out/rm_ncu_bench.log 2>&1
timeout 300 bash tools/ncu_run.sh
tail -1 gpurun_out/rm_smoke.log
cut -c1-200 gpurun_out/rm_bench.json
