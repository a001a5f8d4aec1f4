set -x
# full captures (1 M pairs / reads per launch): the two hot kernels of one bench step, then one launch each of the trim and
# mask kernels of roofline.other_ops.  Launch order of sk_warp_kernel with --steps 1 --warmup 3: 6 warm-up, 2 timed, 2 timed
# with the compaction, 2 under CUDA-event profiling (= 12 demultiplex launches), then 6 x trim, 6 x mask.
B="python bench.py --steps 1 --warmup 3 --pairs 1000000 --skip-e2e --skip-cpu --cli-pairs 0 --skip-configs"
ncu --set full --clock-control none --import-source on -k regex:'sk_(chunk|warp)_kernel' -s 6 -c 2 -f -o gpurun_out/prof_head $B > gpurun_out/ncu_head.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sk_warp_kernel' -s 14 -c 1 -f -o gpurun_out/prof_trim $B > gpurun_out/ncu_trim.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sk_warp_kernel' -s 20 -c 1 -f -o gpurun_out/prof_mask $B > gpurun_out/ncu_mask.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sk_compact_move' -s 2 -c 1 -f -o gpurun_out/prof_move $B > gpurun_out/ncu_move.log 2>&1
# add barcode on the warp engine: tools/stream_time.py launches sk_warp_kernel 7 x for trim, 7 x for mask, then add barcode (1 M reads)
ncu --set full --clock-control none --import-source on -k regex:'sk_warp_kernel' -s 16 -c 1 -f -o gpurun_out/prof_addbc python tools/stream_time.py 1000000 > gpurun_out/ncu_addbc.log 2>&1
tail -3 gpurun_out/ncu_head.log | cut -c1-300
