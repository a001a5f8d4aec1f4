# One GPU call: parity tests, smoke, the default bench line, the ncu launch list of the same command and a full
# capture of the two hot kernels (1 M pairs per launch).  Outputs under gpurun_out/.
set -x
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/rm_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/rm_smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/rm_bench.json 2> gpurun_out/rm_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rm_launches.csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --cli-pairs 0 > gpurun_out/rm_ncu_bench.log 2>&1
timeout 240 bash tools/ncu_run.sh
cat gpurun_out/rm_tests.log
tail -2 gpurun_out/rm_smoke.log
tail -c 600 gpurun_out/rm_bench.err
cut -c1-1500 gpurun_out/rm_bench.json
