# One GPU call while iterating on the warp engine: the GPU parity suite on the in-tree library, then the
# A/B bench of every variant under seqkit_b200/variants/ (tools/runvar.sh).  Outputs under gpurun_out/.
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/ab_tests.log
cat gpurun_out/ab_tests.log
bash tools/runvar.sh 2>&1 | tee gpurun_out/ab_variants.log
