# One call on an 8-GPU box (gpurun --gpus 8): the box's pinned-copy ceiling at N = 1, 2, 4, 8, the bench line at N = 8
# (kernels, e2e through the C ABI, CRC check of per-sample streams across GPUs), the `fasta` binary on all GPUs.
set -x
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/topo.txt
nproc >> gpurun_out/topo.txt
python tools/pcie_ceiling.py > gpurun_out/pcie_ceiling_1.json 2> gpurun_out/pcie_1.err
for n in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/pcie_ceiling.py > gpurun_out/pcie_ceiling_$n.json 2> gpurun_out/pcie_$n.err
done
cat gpurun_out/pcie_ceiling_*.json | cut -c1-600
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 30 --skip-configs --skip-cpu --cli-pairs 0 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
tail -2 gpurun_out/bench_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print('N=$n value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['pcie_gbs'], d['verify'])"
done
timeout 200 python tools/multi_gpu_cli_check.py 2>&1 | grep -v "^    "
timeout 300 python tools/cli_profile.py 4000000 2>&1 | head -3
