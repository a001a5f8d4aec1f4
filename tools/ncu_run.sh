set -x
# full capture of the two hot kernels of one bench step (skip the synth launches: -k regex)
ncu --set full --clock-control none --import-source on -k regex:'sk_(fast|chunk|warp)_kernel' -s 6 -c 2 -f -o gpurun_out/prof_head python bench.py --steps 1 --warmup 3 --pairs 1000000 --skip-e2e --skip-cpu --cli-pairs 0 > gpurun_out/ncu_head.log 2>&1
tail -3 gpurun_out/ncu_head.log | cut -c1-300
