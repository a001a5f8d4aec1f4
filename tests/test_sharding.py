"""The N>1 path on CPU: two gloo ranks each demultiplex their contiguous range of read pairs (with the CPU
oracle standing in for the per-GPU kernels), merge the counters with one all-reduce and their per-sample
files in rank order; the result must be the single-stream oracle's."""
import os
import random
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fuzzgen as G  # noqa: E402
from seqkit_b200 import shard  # noqa: E402


def test_pair_ranges_cover_and_are_contiguous():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 9, 1000, 1001):
            rs = [shard.pair_range(r, world, n) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(hi - lo for lo, hi in rs) - min(hi - lo for lo, hi in rs) <= 1


def test_split_records_cuts_at_four_line_boundaries():
    data = G.clean_fastq(3, 101)
    for world in (1, 2, 3, 8):
        parts = shard.split_records(data, world)
        assert b"".join(parts) == data
        assert all(p.count(b"\n") % 4 == 0 for p in parts)
    ragged = data[:-7]  # last record without its final bytes: still the last rank's
    assert b"".join(shard.split_records(ragged, 4)) == ragged


def _reads(seed, n, bcs):
    rng = random.Random(seed)
    r1, r2 = [], []
    for i in range(n):
        bc = G.observed_barcode(rng, bcs)
        L = rng.randrange(30, 120)
        for mate, dst in ((1, r1), (2, r2)):
            dst.append(G.header(rng, i, mate, bc) + b"\n" + G.rand_seq(rng, L) + b"\n+\n" + G.rand_qual(rng, L) + b"\n")
    return b"".join(r1), b"".join(r2)


def _worker(rank, world, port, sheet, r1, r2, q):
    import torch.distributed as dist
    from oracle import pyoracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p1, p2 = shard.split_records(r1, world)[rank], shard.split_records(r2, world)[rank]
        res = O.demultiplex(sheet, p1, p2)
        counts, total, ident = shard.allreduce_counts(res["counts"], res["total"], res["identified"])
        gathered = [None] * world
        dist.all_gather_object(gathered, res["files"])
        if rank == 0:
            q.put({"counts": counts, "total": total, "identified": ident, "files": shard.merge_files(gathered),
                   "exit_codes": res["exit_code"]})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_reproduce_the_single_stream_output():
    import torch.multiprocessing as mp
    from oracle import pyoracle as O

    sheet, bcs = G.make_sheet(31, 12, 8, umi=4)
    r1, r2 = _reads(7, 1500, bcs)
    want = O.demultiplex(sheet, r1, r2)
    assert want["exit_code"] == 0 and want["identified"] > 0
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sheet, r1, r2, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got["counts"] == want["counts"] and got["total"] == want["total"] and got["identified"] == want["identified"]
    assert got["files"] == want["files"]  # per sample: rank 0's records, then rank 1's = the input order
