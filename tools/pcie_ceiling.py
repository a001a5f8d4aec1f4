#!/usr/bin/env python
"""Host<->device copy ceiling of the box, to put the `e2e` number against ("nobody has measured the
box's plain cudaMemcpyAsync ceiling at N=8").

One process per GPU (run it under torchrun like bench.py, or alone for N=1).  Every rank copies 1 GiB blocks
from / to pinned host memory with plain cudaMemcpyAsync on two streams -- H2D alone, D2H alone, and both at
once (what the e2e loop does) -- for about two seconds per mode, all ranks at the same time.  Two placements of
the host buffers: wherever the process happens to run ("default"), and with the thread and its pinned pages on
the NUMA node of the rank's GPU (sk_bind_thread_to_device).  Rank 0 prints one JSON object:

  python tools/pcie_ceiling.py                                        # N=1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    # libraries print banners on stdout (NCCL: its version line); the JSON object goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from seqkit_b200 import _lib as L
    lib = L.lib()
    nbytes = 1 << 30
    dev_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    affinity0 = os.sched_getaffinity(0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(mode, h_in, h_out, seconds=2.0):
        barrier()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s_in):
                    dev_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s_out):
                    h_out.copy_(dev_out, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
            n += 1
        dt = time.perf_counter() - t0
        gbs = n * nbytes / dt / 1e9
        t = torch.tensor([gbs], device="cuda")
        if world > 1:
            dist.all_reduce(t)  # aggregate over the ranks (every rank ran for the same wall time)
        return float(t.item())

    out = {"n_gpus": world, "block_bytes": nbytes, "placements": {}}
    for placement in ("default", "numa_local"):
        node = None
        if placement == "numa_local":
            node = lib.sk_bind_thread_to_device(local)
        h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h_in.fill_(7)  # first touch
        h_out.fill_(0)
        res = {"numa_node": node, "cpus": len(os.sched_getaffinity(0))}
        for mode in ("h2d", "d2h", "both"):
            res[mode + "_gbs_aggregate"] = run(mode, h_in, h_out)
        res["both_gbs_aggregate"] = {"per_direction": res.pop("both_gbs_aggregate")}
        out["placements"][placement] = res
        del h_in, h_out
        os.sched_setaffinity(0, affinity0)
    if rank == 0:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
