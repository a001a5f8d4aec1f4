# Last GPU call of a round: the whole GPU suite, smoke, the default bench line.
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/rm_tests.log
cat gpurun_out/rm_tests.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py > gpurun_out/rm_bench.json 2> gpurun_out/rm_bench.err
tail -c 300 gpurun_out/rm_bench.err
cut -c1-200 gpurun_out/rm_bench.json
