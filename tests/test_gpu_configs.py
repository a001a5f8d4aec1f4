"""Parity at the shapes BASELINE.json names (-m gpu): the bench configuration itself (configs[4]: fused Q20 trim +
demultiplex on the bench's 384-sample sheet), configs[3] as the real chain `fasta add barcode` -> `fasta
demultiplex` and its --index1/--index2 variant (README.md:44-47, fasta_add_barcode.rs:19-44,
fasta_demultiplex.rs:126-136), configs[0] / [1] at a million reads, and the device-side per-sample compaction
against the per-record slice tables.  Everything goes through the C ABI and is compared byte for byte with
the CPU oracle (oracle/fasta_oracle.c)."""
import ctypes as C

import pytest

import fuzzgen as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    from oracle import pyoracle
    return pyoracle


def _cmp_demux(a, b, ctx=None):
    assert a["exit_code"] == b["exit_code"], ctx
    assert a["stderr"] == b["stderr"], ctx
    assert set(a["files"]) == set(b["files"]), ctx
    for k in b["files"]:
        assert a["files"][k] == b["files"][k], (ctx, k)
    assert a["counts"] == b["counts"] and a["total"] == b["total"] and a["identified"] == b["identified"], ctx


def _synth_pairs(eng, n, seed, with_bc=True):
    n1 = eng.synth(0, n, seed=seed, mate=1, with_bc=with_bc)
    r1 = eng.download_in(0, n1)
    n2 = eng.synth(1, n, seed=seed, mate=2, with_bc=with_bc)
    r2 = eng.download_in(1, n2)
    return r1, r2


def test_bench_configuration_bytes_vs_oracle(O):
    """configs[4] on the bench's own sheet and generator (bench.make_sheet, seed 5 as in bench.synth_pair),
    200 k pairs: fused trim+demultiplex == demultiplex(trim(R1), trim(R2)) of the oracle, every output file."""
    import bench
    from seqkit_b200 import Engine
    n = 200_000
    bcs = bench.make_sheet()
    sheet = bench.sheet_text(bcs)
    with Engine(max_stream_bytes=n * 410 + (1 << 20), max_records=n, max_samples=bench.N_SAMPLES, aux_streams=False) as eng:
        eng.set_sheet(bcs)
        r1, r2 = _synth_pairs(eng, n, seed=5)
        t1, t2 = O.trim_by_quality(r1, bench.MIN_BASEQ), O.trim_by_quality(r2, bench.MIN_BASEQ)
        assert t1[0] == 0 and t2[0] == 0
        want = O.demultiplex(sheet, t1[1], t2[1])
        got = eng.demultiplex(sheet, r1, r2, fused_trim=bench.MIN_BASEQ)
        assert eng.last_result.reserved & 1 and not eng.last_result.reserved & 2  # warp engine, no re-run
        _cmp_demux(got, want, "configs[4]")
        assert 0.9 * n < got["identified"] < n


def _strip_bc(fastq):
    """(reads without their ' BC:' field, the fields themselves in record order)"""
    out, bcs = [], []
    lines = fastq.split(b"\n")
    for i, ln in enumerate(lines):
        if i % 4 == 0 and ln.startswith(b"@"):
            k = ln.rfind(b" BC:")
            bcs.append(ln[k + 4:])
            ln = ln[:k]
        out.append(ln)
    return b"\n".join(out), bcs


def test_config3_chain_add_barcode_then_demultiplex(O):
    """configs[3]: 384 samples, dual index 10+10 + 8 bp UMI.  The reads carry no barcode; a barcode FASTQ holds
    `i7+i5UMI` on its sequence line.  Chain as in README.md:44-47: add barcode to both mates, then demultiplex.
    Variant: the same barcodes as two index FASTQ files (--index1 = i7, --index2 = i5 + UMI)."""
    import bench
    from seqkit_b200 import Engine
    n = 60_000
    bcs = bench.make_sheet()
    sheet = bench.sheet_text(bcs)
    with Engine(max_stream_bytes=n * 480 + (1 << 20), max_records=n, max_samples=bench.N_SAMPLES) as eng:
        eng.set_sheet(bcs)
        r1, r2 = _synth_pairs(eng, n, seed=21)
        p1, obs = _strip_bc(r1)
        p2, obs2 = _strip_bc(r2)
        assert obs == obs2 and len(obs) == n and all(len(b) == 29 for b in obs[:100])
        bcfile = b"".join(b"@bc%07d\n%s\n+\n%s\n" % (i, b, b"I" * len(b)) for i, b in enumerate(obs))
        steps = []
        for reads in (p1, p2):
            got, want = eng.add_barcode(reads, bcfile), O.add_barcode(reads, bcfile)
            assert got == want
            steps.append(got[1])
        assert steps[0] == r1 and steps[1] == r2  # the chain rebuilds the generator's headers
        _cmp_demux(eng.demultiplex(sheet, steps[0], steps[1]), O.demultiplex(sheet, steps[0], steps[1]), "chain")
        i1 = b"".join(b"@i%07d\n%s\n+\n%s\n" % (i, b[:10], b"I" * 10) for i, b in enumerate(obs))
        i2 = b"".join(b"@i%07d\n%s\n+\n%s\n" % (i, b[11:], b"I" * 18) for i, b in enumerate(obs))
        got = eng.demultiplex(sheet, p1, p2, index1=i1, index2=i2)
        want = O.demultiplex(sheet, p1, p2, index1=i1, index2=i2)
        _cmp_demux(got, want, "index route")
        # the two routes assign the same reads (headers differ: the index route leaves them untouched)
        assert got["counts"] == eng.demultiplex(sheet, steps[0], steps[1])["counts"]


def test_config0_and_config1_at_a_million_reads(O):
    """configs[0] (trim by quality, Q20) and configs[1] (mask by quality, 3'-decaying qualities) on 1 M
    single-end 150 bp reads of the bench generator: stdout bytes against the oracle."""
    from seqkit_b200 import Engine
    n = 1_000_000
    with Engine(max_stream_bytes=n * 360 + (1 << 20), max_records=n, max_samples=0, aux_streams=False) as eng:
        n1 = eng.synth(0, n, seed=1, mate=1, with_bc=False)
        data = eng.download_in(0, n1)
        assert data.count(b"\n") == 4 * n
        got, want = eng.trim_by_quality(data, 20), O.trim_by_quality(data, 20)
        assert got[0] == want[0] == 0 and got[1] == want[1]
        assert eng.last_result.reserved & 1 and not eng.last_result.reserved & 2
        n1 = eng.synth(0, n, seed=2, mate=1, with_bc=False)
        data = eng.download_in(0, n1)
        got, want = eng.mask_by_quality(data, 20), O.mask_by_quality(data, 20)
        assert got[0] == want[0] == 0 and got[1] == want[1]
        assert got[1].count(b"N") > data.count(b"N")


def test_compaction_equals_the_slice_tables():
    """sk_demux_compact (one contiguous run per sample and mate, sk_compact.cu) against the per-record slice
    tables read with the host helper sk_demux_gather: same bytes for every sample, on the warp engine (paired,
    single-end, fused trim, several rounds per tile) and on the general engine (index route); slices start on
    128-byte lines and account for every payload byte."""
    import random
    from seqkit_b200 import Engine, _lib as L
    rng = random.Random(77)
    with Engine(max_stream_bytes=48 << 20, max_records=1 << 18, max_samples=512) as eng:
        lib = eng.lib
        cases = []
        for it in range(10):
            S = rng.choice((1, 3, 24, 96, 384))
            sheet, bcs = G.make_sheet(rng.randrange(1 << 30), S, rng.choice((6, 8, 20)), umi=rng.choice((0, 4, 8)),
                                      dual=rng.random() < 0.4)
            n = rng.choice((0, 1, 33, 500, 6000))
            r1, r2 = G.clean_pairs(rng.randrange(1 << 30), n, bcs, p_sub=0.05, p_random=0.1,
                                   read_len=rng.choice(((20, 60), (100, 160))))
            cases.append((sheet, r1, r2 if it % 3 else None, {"fused_trim": 20} if it % 2 else {}))
        sheet, bcs = G.make_sheet(9, 12, 16, umi=4, dual=True)
        r1, r2 = G.clean_pairs(10, 300, bcs, bc_in_r2=False)
        lit = [b.rstrip(b"U") for b in bcs]
        i1 = G.index_reads(11, 300, [b.split(b"+")[0] for b in lit])
        i2 = G.index_reads(12, 300, [b.split(b"+")[1] + b"ACGT" for b in lit])
        cases.append((sheet, r1, r2, {"index1": i1, "index2": i2}))
        for sheet, r1, r2, kw in cases:
            eng.compact = True
            a = eng.demultiplex(sheet, r1, r2, **kw)
            res = eng.last_result
            if a["exit_code"] == 0 and r1:
                assert res.reserved & 8, "the compacted path must have run"
                S = eng.S
                sl = (L.Slice * (S + 1))()
                for m in range(2 if r2 is not None else 1):
                    assert lib.sk_download_slices(eng.ctx, 0, m, sl) == 0
                    eng.wait()
                    assert all(sl[s].offset % 128 == 0 for s in range(S))
                    assert all(sl[s].offset + sl[s].len <= sl[s + 1].offset for s in range(S))
                    assert sum(sl[s].len for s in range(S)) == sl[S].len == res.out_bytes[m]
            eng.compact = False
            b = eng.demultiplex(sheet, r1, r2, **kw)
            assert not eng.last_result.reserved & 8
            assert a["files"] == b["files"] and a["counts"] == b["counts"] and a["stderr"] == b["stderr"]
        eng.compact = True


def test_host_generator_is_the_device_generator():
    """oracle/synth_host.c (input of `bench.py --impl reference`, which must not load the product's library)
    yields the bytes of the device generator for the same (seed, pair range, mate, sheet)."""
    import bench
    from oracle import pyoracle as O
    from seqkit_b200 import Engine
    bcs = bench.make_sheet()
    with Engine(max_stream_bytes=32 << 20, max_records=1 << 16, max_samples=bench.N_SAMPLES, aux_streams=False) as eng:
        eng.set_sheet(bcs)
        for seed, first, n, mate, with_bc, prof in ((5, 0, 3000, 1, True, 0), (5, 123456789, 2000, 2, True, 0),
                                                    (1, 7, 2500, 1, False, 0), (9, 1 << 40, 1000, 2, True, 1)):
            nb = eng.synth(0, n, seed=seed, first_pair=first, mate=mate, with_bc=with_bc, qual_profile=prof)
            dev = eng.download_in(0, nb)
            host = O.synth_fastq(n, seed=seed, first_pair=first, mate=mate, barcodes=bcs if with_bc else None, qual_profile=prof)
            assert dev == host, (seed, first, mate, with_bc)
    r1, r2 = bench.host_pairs(bcs, 5000, first_pair=11, seed=5, threads=3)
    assert r1 == O.synth_fastq(5000, seed=5, first_pair=11, mate=1, barcodes=bcs)
    assert r2 == O.synth_fastq(5000, seed=5, first_pair=11, mate=2, barcodes=bcs)


def test_run_totals_accumulate_on_the_device():
    """sk_counts_accumulate / sk_download_totals: the counters of finished batches add up on the device (what the
    `fasta` binary merges over its GPUs with one grouped NCCL all-reduce; one context: nothing to merge)."""
    import numpy as np
    import bench
    from seqkit_b200 import Engine, _lib as L
    S = 96
    sheet, bcs = G.make_sheet(3, S, 8)
    with Engine(max_stream_bytes=32 << 20, max_records=1 << 16, max_samples=S, aux_streams=False) as eng:
        lib = eng.lib
        eng.set_sheet(bcs)
        want = np.zeros(S + 2, dtype=np.uint64)
        opts = L.DemuxOpts(-1, 0, 0, 0, 0)
        for k in range(3):
            eng.synth(0, 20000, seed=4, first_pair=20000 * k, mate=1, with_bc=True)
            eng.synth(1, 20000, seed=4, first_pair=20000 * k, mate=2, with_bc=True)
            assert lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0
            eng.wait()
            c = np.zeros(S + 2, dtype=np.uint64)
            assert lib.sk_download_counts(eng.ctx, 0, c.ctypes.data) == 0
            want += c
            assert lib.sk_counts_accumulate(eng.ctx, 0) == 0
        ctxs = (C.c_void_p * 1)(eng.ctx)
        assert lib.sk_allreduce_totals(ctxs, 1) == 0
        got = np.zeros(S + 2, dtype=np.uint64)
        assert lib.sk_download_totals(eng.ctx, got.ctypes.data) == 0
        assert np.array_equal(got, want) and int(got[S]) == 60000
        assert lib.sk_totals_reset(eng.ctx) == 0
        assert lib.sk_download_totals(eng.ctx, got.ctypes.data) == 0 and int(got.sum()) == 0
