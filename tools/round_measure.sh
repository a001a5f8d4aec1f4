# One GPU call at the end of a round: parity tests, smoke, the default bench line, the ncu launch list of the same command
# and full captures of the hot kernels (1 M pairs per launch).  Outputs under gpurun_out/; tools/refresh_profiles.py r2_warp
# then copies the summaries into profiles/.
set -x
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -30 > gpurun_out/rm_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/rm_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/rm_bench.json 2> gpurun_out/rm_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/rm_bench_ref.json 2> gpurun_out/rm_bench_ref.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sk_ -c 400 --csv --log-file gpurun_out/rm_launches.csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --cli-pairs 0 --skip-configs > gpurun_out/rm_ncu_bench.log 2>&1
timeout 400 bash tools/ncu_run.sh
cat gpurun_out/rm_tests.log
tail -2 gpurun_out/rm_smoke.log
tail -c 600 gpurun_out/rm_bench.err
cut -c1-1500 gpurun_out/rm_bench.json
cut -c1-600 gpurun_out/rm_bench_ref.json
