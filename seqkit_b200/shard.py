"""Multi-GPU sharding of the demultiplex path: records are independent, so every
rank takes a contiguous range of read pairs, writes its own per-sample slices, and the only exchange is
one all-reduce (sum, u64, S + 2 values) of the counters (fasta_demultiplex.rs:108-109,169,177-178).  The
per-sample files of the job are the ranks' files concatenated in rank order, which is the order the
single-stream reference writes them in (fasta_demultiplex.rs:196-238).

This module is host logic only (no kernels): it is what bench.py's weak-scaling run and a multi-GPU
driver of the `fasta` binary do around the per-GPU sk_demultiplex calls."""
from __future__ import annotations


def pair_range(rank: int, world: int, n_pairs: int) -> tuple[int, int]:
    """Pairs [lo, hi) of rank `rank`: contiguous, covering, sizes differ by at most one."""
    if not (0 <= rank < world) or n_pairs < 0:
        raise ValueError("bad rank / world / n_pairs")
    base, extra = divmod(n_pairs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def record_offsets(data: bytes) -> list[int]:
    """Byte offset of every 4-line record of a FASTQ buffer plus the end offset (common.rs:106-112: records
    are four lines; a trailing partial record counts as one, like the reference's last read_line calls)."""
    offs, pos, n = [0], 0, len(data)
    while pos < n:
        for _ in range(4):
            nl = data.find(b"\n", pos)
            pos = n if nl < 0 else nl + 1
            if pos >= n:
                break
        offs.append(pos)
    return offs


def split_records(data: bytes, world: int) -> list[bytes]:
    """The buffer cut into `world` contiguous record ranges (rank order)."""
    offs = record_offsets(data)
    n = len(offs) - 1
    return [data[offs[lo]:offs[hi]] for lo, hi in (pair_range(r, world, n) for r in range(world))]


def merge_files(per_rank: list[dict[str, bytes]]) -> dict[str, bytes]:
    """Per-sample output of the job: every file is the ranks' files appended in rank order."""
    out: dict[str, bytes] = {}
    for files in per_rank:
        for name, blob in files.items():
            out[name] = out.get(name, b"") + blob
    return out


def allreduce_counts(counts: list[int], total: int, identified: int):
    """[per-sample..., total, identified] summed over the ranks of the default torch.distributed group (gloo on
    CPU; on the GPUs the same payload goes through sk_allreduce_counts = one ncclAllReduce)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(counts) + [total, identified], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    v = [int(x) for x in t.tolist()]
    return v[:-2], v[-2], v[-1]
