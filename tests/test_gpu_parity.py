"""Parity tests proper (-m gpu): every operator goes through the C ABI (libseqkit_b200.so) and is
compared byte for byte with the CPU oracle on the same seeded inputs, with the committed golden
fixtures, and -- at sizes the oracle cannot reach -- through size-independent properties."""
import os
import random

import pytest

import fuzzgen as G
import golden_util as GU

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from seqkit_b200 import Engine
    e = Engine(max_stream_bytes=48 << 20, max_records=1 << 18, max_samples=512)
    yield e
    e.close()


@pytest.fixture(scope="module")
def O():
    from oracle import pyoracle
    return pyoracle


def fq(qual, seq=None, hdr=b"@r"):
    seq = seq if seq is not None else b"A" * len(qual)
    return hdr + b"\n" + seq + b"\n+\n" + qual + b"\n"


def check3(a, b, ctx=None):
    assert a[0] == b[0], ctx
    assert a[1] == b[1], ctx
    if a[0] != 101:
        assert a[2] == b[2], ctx


def test_native_library_is_the_path(eng):
    import ctypes
    from seqkit_b200 import LIB_PATH
    assert LIB_PATH.endswith("libseqkit_b200.so")
    assert isinstance(eng.lib, ctypes.CDLL)


@pytest.mark.parametrize("qual,kept", [(b"IIIIIIIIII", 10), (b"IIIIIII###", 7), (b"##########", 0), (b"IIII#IIII#", 9),
                                       (b"III#I#I#I#", 9), (b"IIIIII5555", 10), (b"IIIIIIII4", 8), (b"I", 1), (b"#", 0),
                                       (b"", 0)])
def test_trim_kats(eng, qual, kept):
    seq = bytes(b"ACGT"[i % 4] for i in range(len(qual)))
    code, out, err = eng.trim_by_quality(fq(qual, seq), 20)
    assert code == 0 and err == b""
    assert out == (b"@r\nN\n+\n!\n" if kept == 0 else b"@r\n" + seq[:kept] + b"\n+\n" + qual[:kept] + b"\n")


def test_mask_kat_and_error_paths(eng, O):
    assert eng.mask_by_quality(b"@h x\nACGTA\n+junk\nI#5 4\n", 20) == (0, b"@h x\nANGTN\n+\nI#5 4\n", b"")
    for data in (fq(b"IIII") + b"\n", fq(b"IIII") + b"@x\nAC\n+\nIII\n", b"@p\nAC\n+\nIIIII\n", b"@t", b"@t\nACG", b"",
                 b"\n", b"@a\n", b"@a\nAC\n+", fq(b"II") * 3 + b"X\n"):
        for q in (20, 0, 255):
            check3(eng.trim_by_quality(data, q), O.trim_by_quality(data, q), (data, q))
            check3(eng.mask_by_quality(data, q), O.mask_by_quality(data, q), (data, q))


def test_golden_stream_ops(eng):
    n = 0
    for c in GU.cases():
        if c["op"] == "demux":
            continue
        if c["op"] == "trim":
            got = eng.trim_by_quality(GU.blob(c["input"]), c["min_baseq"])
        elif c["op"] == "mask":
            got = eng.mask_by_quality(GU.blob(c["input"]), c["min_baseq"])
        else:
            got = eng.add_barcode(GU.blob(c["input"]), GU.blob(c["barcodes"]))
        assert got[0] == c["exit_code"], c["tag"]
        assert got[1] == GU.blob(c["stdout"]), c["tag"]
        if got[0] != 101:
            assert got[2] == GU.blob(c["stderr"]), c["tag"]
        n += 1
    assert n > 50


def test_golden_demux(eng):
    for c in GU.cases("demux"):
        res = eng.demultiplex(GU.blob(c["sheet"]), GU.blob(c["r1"]), GU.blob(c["r2"]))
        assert res["exit_code"] == c["exit_code"], c["tag"]
        assert res["stderr"] == GU.blob(c["stderr"]), c["tag"]
        assert res["counts"] == c["counts"] and res["total"] == c["total"] and res["identified"] == c["identified"]
        assert set(res["files"]) == set(c["files"])
        for k, v in c["files"].items():
            assert res["files"][k] == GU.blob(v), (c["tag"], k)


def test_fuzz_stream_ops_vs_oracle(eng, O):
    rng = random.Random(2024)
    for it in range(120):
        n = rng.choice((0, 1, 2, 7, 40, 200))
        data = G.nasty_fastq(rng.randrange(1 << 30), n) if it % 3 else G.clean_fastq(rng.randrange(1 << 30), n, qual_style="mix")
        q = rng.choice((0, 2, 10, 20, 30, 41, 93, 200, 255))
        check3(eng.trim_by_quality(data, q), O.trim_by_quality(data, q), ("trim", it, q))
        check3(eng.mask_by_quality(data, q), O.mask_by_quality(data, q), ("mask", it, q))
        bc = G.index_reads(rng.randrange(1 << 30), rng.choice((0, 1, n, n + 3, max(n - 2, 0))), [b"ACGT", b"GG+TT", b"ACGTACGTAC"])
        check3(eng.add_barcode(data, bc), O.add_barcode(data, bc), ("addbc", it))


def test_dense_long_and_utf8_records_take_the_line_engine(eng, O):
    """Inputs outside the chunk engines' geometry.  trim / mask by quality fall through to the
    line engine (sk_result.reserved bit 4), which frames records by a global line table: records of any length,
    any density, and UTF-8 in header and '+' lines give the oracle's bytes.  Non-ASCII bases or qualities and
    invalid UTF-8 are still refused, and so are demultiplex inputs of that kind -- explicitly."""
    from seqkit_b200.engine import Unsupported
    tiny = b"@a\nAC\n+\nII\n" * 4000  # 11-byte records: > 512 per 32 KiB chunk
    big = (b"@long\n" + b"A" * 20000 + b"\n+\n" + b"I" * 12000 + b"#" * 8000 + b"\n") * 3
    mixed = G.clean_fastq(3, 500) + big + tiny + G.clean_fastq(4, 500)
    utf8 = "@réad 日本 1\nACGTACGT\n+réad\nIIII##II\n".encode() * 50 + G.clean_fastq(5, 200)
    nbsp = "@x \nACGT\n+\nIIII\n".encode()  # header ends in U+00A0: printed verbatim by trim and mask
    for label, blob in (("tiny", tiny), ("big", big), ("mixed", mixed), ("utf8", utf8), ("nbsp", nbsp)):
        for q in (20, 0, 41):
            check3(eng.trim_by_quality(blob, q), O.trim_by_quality(blob, q), ("trim", label, q))
            assert eng.last_result.reserved & 16, ("trim", label, eng.last_result.reserved)
            check3(eng.mask_by_quality(blob, q), O.mask_by_quality(blob, q), ("mask", label, q))
            assert eng.last_result.reserved & 16, ("mask", label, eng.last_result.reserved)
    # add barcode (fasta_add_barcode.rs:19-44): header.trim_end() is Unicode-aware, every line is copied as it is
    bcf = G.index_reads(9, 5000, [b"ACGTACGT", b"GG+TT"])
    for label, blob in (("tiny", tiny), ("big", big), ("mixed", mixed), ("utf8", utf8), ("nbsp", nbsp)):
        check3(eng.add_barcode(blob, bcf), O.add_barcode(blob, bcf), ("addbc", label))
        assert eng.last_result.reserved & 16, ("addbc", label, eng.last_result.reserved)
    fa_long = (b">s desc \n" + b"A" * 30000 + b"\n") * 5
    check3(eng.add_barcode(fa_long, b">b\nAA\n>b\nCC\n"), O.add_barcode(fa_long, b">b\nAA\n>b\nCC\n"), "addbc long FASTA")
    check3(eng.add_barcode(big + b"oops\nAC\n+\nII\n" + big, bcf), O.add_barcode(big + b"oops\nAC\n+\nII\n" + big, bcf), "addbc bad line")
    # failing records behind long ones: the reference's messages after the output of the records before them
    for bad in (big + b"oops\nAC\n+\nII\n" + big, big + b"@s\nACGT\n+\nII\n", tiny + b"\n"):
        check3(eng.trim_by_quality(bad, 20), O.trim_by_quality(bad, 20), "trim after long")
        check3(eng.mask_by_quality(bad, 20), O.mask_by_quality(bad, 20), "mask after long")
    for blob in ("@r\nACéT\n+\nIIIII\n".encode(), "@r\nACGT\n+\nIIé\n".encode(), b"@r\xff\nACGT\n+\nIIII\n", b"@r\nACGT\n+\xc3\nIIII\n"):
        with pytest.raises(Unsupported):
            eng.trim_by_quality(blob, 20)
        with pytest.raises(Unsupported):
            eng.mask_by_quality(blob, 20)
    sheet, bcs = G.make_sheet(1, 4, 8)
    with pytest.raises(Unsupported):  # a record whose output does not fit a 16-bit group length
        eng.demultiplex(sheet, (b"@long BC:" + bcs[0] + b"\n" + b"A" * 40000 + b"\n+\n" + b"I" * 40000 + b"\n") * 3)
    utf8_seq = "@r BC:".encode() + bcs[0] + "\nACé\n+\nIIII\n".encode()
    _cmp_demux(eng.demultiplex(sheet, utf8_seq), O.demultiplex(sheet, utf8_seq), "plain demultiplex copies lines: UTF-8 anywhere")
    with pytest.raises(Unsupported):  # the fused quality trim wants ASCII bases and qualities
        eng.demultiplex(sheet, utf8_seq, fused_trim=20)
    with pytest.raises(Unsupported):  # invalid UTF-8 (the reference: I/O error while reading from file)
        eng.demultiplex(sheet, b"@r BC:" + bcs[0] + b"\nAC\xff\n+\nIII\n")


def test_demultiplex_takes_the_line_engine_for_long_dense_and_utf8_records(eng, O):
    """Header-route demultiplex of batches outside the chunk engines' geometry (sk_result.reserved bit 4): long and
    dense records, UTF-8 and Unicode white space in headers -- single and paired, plain and with the fused quality trim,
    with the device-side compaction -- against the oracle (fasta_demultiplex.rs:117-249)."""
    rng = random.Random(77)
    sheet, bcs = G.make_sheet(5, 24, 8, umi=4)
    nb = lambda: G.observed_barcode(rng, bcs, p_sub=0.04, p_random=0.05)

    def rec(name, bc, seq, qual, tail=b""):
        return b"@" + name + b" BC:" + bc + tail + b"\n" + seq + b"\n+\n" + qual + b"\n"

    def quals(n):
        return bytes(rng.choice(b"#+5?IIII") for _ in range(n))

    def long_pair(i, n):
        bc = nb()
        return (rec(b"L%d 1:N" % i, bc, G.rand_seq(rng, n), quals(n)), rec(b"L%d 2:N" % i, bc, G.rand_seq(rng, n), quals(n)))

    def check(r1, r2, ctx, want_line=True):
        for fused in (None, 20):
            if fused is None:
                want = O.demultiplex(sheet, r1, r2)
            else:
                t1 = O.trim_by_quality(r1, fused)
                t2 = O.trim_by_quality(r2, fused) if r2 is not None else None
                if t1[0] or (t2 is not None and t2[0]):
                    continue  # the trim of the reference's pipeline fails first: not this test's business
                want = O.demultiplex(sheet, t1[1], t2[1] if t2 is not None else None)
            got = eng.demultiplex(sheet, r1, r2, fused_trim=fused) if fused is not None else eng.demultiplex(sheet, r1, r2)
            _cmp_demux(got, want, (ctx, fused))
            if want_line:
                assert eng.last_result.reserved & 16, (ctx, fused, eng.last_result.reserved)

    # long records among ordinary ones
    p1, p2 = G.clean_pairs(31, 3000, bcs)
    longs = [long_pair(i, n) for i, n in enumerate((9000, 15000, 25000, 12000))]
    r1 = p1 + b"".join(a for a, _ in longs) + p1
    r2 = p2 + b"".join(b for _, b in longs) + p2
    check(r1, r2, "long paired")
    check(r1, None, "long single")
    # denser than any chunk engine takes
    tiny1 = b"".join(rec(b"t", nb(), b"", b"") for _ in range(6000))
    check(tiny1, None, "dense single")
    # UTF-8 in headers and '+' lines, Unicode white space at the end of a header and around the barcode
    u1, u2 = [], []
    for i in range(400):
        bc = nb()
        name = ("ré%d 日本" % i).encode()
        tail = rng.choice((b"", " \u00a0".encode(), "\u2003\u3000".encode(), b" x:1", " caf\u00e9 \u2028".encode()))
        sq, q = G.rand_seq(rng, 40), quals(40)
        u1.append(b"@" + name + b" 1 BC:" + bc + tail + b"\n" + sq + b"\n+" + "ü".encode() + b"\n" + q + b"\n")
        u2.append(b"@" + name + b" 2 BC:" + bc + tail + b"\n" + sq + b"\n+\n" + q + b"\n")
    check(b"".join(u1), b"".join(u2), "utf8 paired")
    check(p1 + b"".join(u1), p2 + b"".join(u2), "utf8 behind ordinary pairs")
    # failures behind a long record: message and the files written so far
    bad = p1[: len(p1) // 2] + longs[0][0] + b"@nobc\nAC\n+\nII\n" + p1
    check(bad, None, "no barcode behind a long record", want_line=False)
    # --index1 / --index2 route (:126-136): the headers stay whole, the barcode comes from the index reads
    sheet2, bcs2 = G.make_sheet(6, 12, 16, umi=4, dual=True)
    lit = [b.rstrip(b"U") for b in bcs2]
    n_i = 600
    q1, q2 = G.clean_pairs(33, n_i, bcs2, bc_in_r2=False)
    recs1, recs2 = q1.split(b"\n@"), q2.split(b"\n@")
    for k in (5, 300):  # two long reads, one with a UTF-8 header
        name = ("@né%d" % k).encode() if k == 300 else b"@long%d" % k
        big_rec = lambda m: name + b" %d\n" % m + G.rand_seq(rng, 12000) + b"\n+\n" + quals(12000)
        recs1[k] = big_rec(1)[1:] if k else big_rec(1)
        recs2[k] = big_rec(2)[1:] if k else big_rec(2)
    q1, q2 = b"\n@".join(recs1), b"\n@".join(recs2)
    i1 = G.index_reads(34, n_i, [b.split(b"+")[0] for b in lit], p_sub=0.03)
    i2 = b"".join(b"@i\n" + r.split(b"\n")[1] + b"ACGT\n+\nIIII\n" for r in G.index_reads(35, n_i, [b.split(b"+")[1] for b in lit], p_sub=0.03).split(b"@")[1:])
    for a2 in (q2, None):
        _cmp_demux(eng.demultiplex(sheet2, q1, a2, index1=i1, index2=i2), O.demultiplex(sheet2, q1, a2, index1=i1, index2=i2), ("index route", a2 is None))
        assert eng.last_result.reserved & 16
    _cmp_demux(eng.demultiplex(sheet2, q1, q2, index1=i1[: len(i1) // 2], index2=i2), O.demultiplex(sheet2, q1, q2, index1=i1[: len(i1) // 2], index2=i2), "index file runs out")
    amb_sheet = b"P\tACGTACGT\nQ\tACGTACGA\n"
    amb = rec(b"a", b"ACGTACGC", b"A" * 9000, b"I" * 9000) * 3
    _cmp_demux(eng.demultiplex(amb_sheet, amb), O.demultiplex(amb_sheet, amb), "ambiguous long records")


def test_add_barcode_fasta_and_reuse(eng, O):
    fa = b">s1 desc\nACGT\n>s2\nGGCC\n>s3\nTT\n"
    for bc in (b">b\nAA\n>b\nCC\n", b">b\nAA\n", b"", b"@x\nACGT+TTAA\n+\nIIIIIIIII\n@y\nGGGG\n+\nIIII\n", b"junk\nlines\n"):
        check3(eng.add_barcode(fa, bc), O.add_barcode(fa, bc), bc)
    check3(eng.add_barcode(b">s\nACGT\nbad\n", b">b\nAA\n>b\nCC\n"), O.add_barcode(b">s\nACGT\nbad\n", b">b\nAA\n>b\nCC\n"))


def _cmp_demux(a, b, ctx=None):
    assert a["exit_code"] == b["exit_code"], ctx
    if a["exit_code"] != 101:
        assert a["stderr"] == b["stderr"], ctx
    assert a["files"] == b["files"], ctx
    if a["exit_code"] == 0:
        assert a["counts"] == b["counts"] and a["total"] == b["total"] and a["identified"] == b["identified"], ctx


def test_demux_kats(eng, O):
    sheet = b"A\tACGTACGT\nB\tACGTACGA\n"
    rec = lambda name, bc: b"@%s BC:%s\nAC\n+\nII\n" % (name, bc)
    cases = [
        (sheet, rec(b"r1", b"ACGTACGC") + rec(b"r2", b"ACGTACGT") + rec(b"r3", b"NCGTACGT"), None, {}),
        (b"X\tACGTUUUU\n", rec(b"u1 1:N", b"ACGTTTGA"), b"@u1 2:N BC:ACGTTTGA\nGG\n+\nII\n", {}),
        (b"P\tACGT\nQ\tACGT\n", rec(b"d", b"ACGT"), None, {}),
        (sheet, b"@x\nA\n+\nI\n", None, {}),
        (sheet, rec(b"ok", b"ACGTACGT") + rec(b"x", b"ACG"), None, {}),
        (b"A\tAC\nA\tGG\n", b"", None, {}),
        (b"S\tAC+GT\n", b"@r BC:zz\nAC\n+\nII\n", None, {"index1": b"@i\nAC\n+\nII\n", "index2": b"@i\nGT\n+\nII\n"}),
        (sheet, rec(b"ok", b"ACGTACGT") + b"bad\nAC\n+\nII\n", None, {}),
    ]
    for sh, r1, r2, kw in cases:
        _cmp_demux(eng.demultiplex(sh, r1, r2, **kw), O.demultiplex(sh, r1, r2, **kw), (sh, r1))


def test_fuzz_demux_vs_oracle(eng, O):
    rng = random.Random(99)
    for it in range(60):
        S = rng.choice((1, 2, 3, 8, 20, 96))
        Lb = rng.choice((4, 6, 8, 12, 20))
        umi = rng.choice((0, 0, 3, 8))
        sheet, bcs = G.make_sheet(rng.randrange(1 << 30), S, Lb, umi=umi, dual=rng.random() < 0.3,
                                  min_dist=rng.choice((0, 1, 2, 3)), wild_n=rng.choice((0, 0, 0.1)))
        n = rng.choice((0, 1, 5, 60, 400))
        if it % 2:
            r1, r2 = G.nasty_headers_pairs(rng.randrange(1 << 30), n, bcs)
        else:
            r1, r2 = G.clean_pairs(rng.randrange(1 << 30), n, bcs, p_sub=0.1, p_n=0.05, p_random=0.1, p_lower=0.02,
                                   bc_in_r2=rng.random() < 0.5)
        for paired in (True, False):
            _cmp_demux(eng.demultiplex(sheet, r1, r2 if paired else None), O.demultiplex(sheet, r1, r2 if paired else None),
                       ("it", it, paired))


def test_demux_index_route(eng, O):
    rng = random.Random(5)
    for it in range(8):
        sheet, bcs = G.make_sheet(100 + it, 12, 16, umi=rng.choice((0, 4)), dual=True)
        n = rng.choice((1, 30, 200))
        r1, r2 = G.clean_pairs(200 + it, n, bcs, bc_in_r2=False)
        lit = [b.rstrip(b"U") for b in bcs]
        half1 = [b.split(b"+")[0] for b in lit]
        half2 = [b.split(b"+")[1] for b in lit]
        i1 = G.index_reads(300 + it, n, half1, p_sub=0.03)
        i2 = G.index_reads(400 + it, n, half2, p_sub=0.03)
        if rng.random() < 0.5 and bcs[0].endswith(b"U"):
            # UMI bases ride at the end of the second index read
            i2 = b"".join(b"@i\n" + rec.split(b"\n")[1] + b"ACGT"[:len(bcs[0]) - len(lit[0])] + b"\n+\nIIII\n"
                          for rec in i2.split(b"@")[1:])
        _cmp_demux(eng.demultiplex(sheet, r1, r2, index1=i1, index2=i2), O.demultiplex(sheet, r1, r2, index1=i1, index2=i2), it)


def test_fused_trim_demux_equals_composition(eng, O):
    """North-star config 5: fused trim+demux == demultiplex(trim(R1), trim(R2)) of the reference."""
    for seed in range(4):
        sheet, bcs = G.make_sheet(seed, 24, 20, umi=8, dual=True)
        r1, r2 = G.clean_pairs(50 + seed, 300, bcs, qual_style="decay")
        t1, t2 = O.trim_by_quality(r1, 20), O.trim_by_quality(r2, 20)
        assert t1[0] == 0 and t2[0] == 0
        want = O.demultiplex(sheet, t1[1], t2[1])
        got = eng.demultiplex(sheet, r1, r2, fused_trim=20)
        _cmp_demux(got, want, seed)


def test_large_synthetic_multi_chunk(eng, O):
    """~60 MB per stream: thousands of chunks, so the look-back chains and the demux slice tables
    are exercised for real; the oracle still finishes in seconds."""
    sheet, bcs = G.make_sheet(7, 96, 8)
    eng.set_sheet(bcs)
    n = 120_000
    n1 = eng.synth(0, n, seed=11, mate=1, with_bc=True)
    r1 = eng.download_in(0, n1)
    n2 = eng.synth(1, n, seed=11, mate=2, with_bc=True)
    r2 = eng.download_in(1, n2)
    assert r1.count(b"\n") == 4 * n and r2.count(b"\n") == 4 * n
    check3(eng.trim_by_quality(r1, 20), O.trim_by_quality(r1, 20), "trim-large")
    check3(eng.mask_by_quality(r1, 20), O.mask_by_quality(r1, 20), "mask-large")
    _cmp_demux(eng.demultiplex(sheet, r1, r2), O.demultiplex(sheet, r1, r2), "demux-large")
    t1, t2 = O.trim_by_quality(r1, 20), O.trim_by_quality(r2, 20)
    _cmp_demux(eng.demultiplex(sheet, r1, r2, fused_trim=20), O.demultiplex(sheet, t1[1], t2[1]), "fused-large")


def test_dual_index_umi_384_samples(eng, O):
    """Config-4 shape: 384 samples, i7(10)+i5(10)+UMI(8) = 29-character sheet barcodes."""
    sheet, bcs = G.make_sheet(4, 384, 20, umi=8, dual=True)
    eng.set_sheet(bcs)
    n = 40_000
    n1 = eng.synth(0, n, seed=4, mate=1, with_bc=True)
    r1 = eng.download_in(0, n1)
    n2 = eng.synth(1, n, seed=4, mate=2, with_bc=True)
    r2 = eng.download_in(1, n2)
    _cmp_demux(eng.demultiplex(sheet, r1, r2), O.demultiplex(sheet, r1, r2), "cfg4")


def test_properties_at_scale(eng):
    """Size-independent checks on a batch too big for a byte-for-byte oracle run in the test budget:
    mask keeps every byte count, trim output re-trims to itself (idempotence), demux conserves reads."""
    n = 130_000
    n1 = eng.synth(0, n, seed=3, mate=1)
    data = eng.download_in(0, n1)
    code, masked, _ = eng.mask_by_quality(data, 20)
    assert code == 0 and len(masked) == len(data) and masked.count(b"\n") == data.count(b"\n")
    code, trimmed, _ = eng.trim_by_quality(data, 20)
    assert code == 0 and trimmed.count(b"\n") == 4 * n
    code2, again, _ = eng.trim_by_quality(trimmed, 20)
    assert code2 == 0 and again.count(b"\n") == 4 * n and len(again) <= len(trimmed)


def _bc_headers(data, bcs, seed):
    """Appends ' BC:<barcode>' (a sheet barcode with its U positions filled) to every header line."""
    rng = random.Random(seed)
    out, lines = [], data.split(b"\n")
    for i, ln in enumerate(lines):
        if i % 4 == 0 and ln.startswith(b"@"):
            bc = bytes(rng.choice(b"ACGT") if ch in b"UN" else ch for ch in rng.choice(bcs))
            ln = ln + b" BC:" + bc
        out.append(ln)
    return b"\n".join(out)


def test_warp_engine_is_the_default_and_reruns_beyond_its_limits(O, monkeypatch):
    """Trim, mask and header-route demultiplex run on the warp engine (sk_warp.cu).  Records of ~5.5 KB are
    longer than its overhang (1200 B with the tile pinned at 29 lanes) but inside the general engine's
    (6128 B): as soon as one of them starts near the end of a tile the operator is re-run on the general
    engine (sk_result.reserved bit 1), and the bytes stay those of the oracle either way."""
    from seqkit_b200 import Engine
    monkeypatch.setenv("SK_TILE_LANES", "29")
    with Engine(max_stream_bytes=48 << 20, max_records=1 << 18, max_samples=512) as eng:
        _rerun_body(eng, O)


def _rerun_body(eng, O):
    sheet, bcs = G.make_sheet(3, 24, 8, umi=4)
    data = G.clean_fastq(5, 3000, qual_style="decay")
    rng = random.Random(17)
    long_list = []
    for i in range(60):
        n = rng.randrange(2100, 2900)  # record of 4.2 .. 5.8 KB
        long_list.append(b"@long read %d\n" % i + bytes(rng.choice(b"ACGT") for _ in range(n)) + b"\n+\n" +
                         b"I" * (n // 2) + b"#" * (n - n // 2) + b"\n")
    long_recs = b"".join(long_list)
    # the construction must really leave the warp engine's window (tile 11600 B + 1200 B overhang) somewhere
    pos, overruns = len(data), 0
    for rec in long_list:
        start_in_tile = pos - (pos // 11600) * 11600
        overruns += start_in_tile + len(rec) > 12800
        pos += len(rec)
    assert overruns > 0
    for label, blob, rerun in (("normal", data, False), ("long", data + long_recs + data, True)):
        check3(eng.trim_by_quality(blob, 20), O.trim_by_quality(blob, 20), ("trim", label))
        assert eng.last_result.reserved == (2 if rerun else 1), ("trim", label, eng.last_result.reserved)
        check3(eng.mask_by_quality(blob, 20), O.mask_by_quality(blob, 20), ("mask", label))
        assert eng.last_result.reserved == (2 if rerun else 1), ("mask", label, eng.last_result.reserved)
        r1 = _bc_headers(blob, bcs, 7)
        for fused in (None, 20):
            want_in = r1 if fused is None else O.trim_by_quality(r1, fused)[1]
            _cmp_demux(eng.demultiplex(sheet, r1, None, fused_trim=fused), O.demultiplex(sheet, want_in, None), (label, fused))
            assert eng.last_result.reserved & 3 == (2 if rerun else 1), ("demux", label, fused, eng.last_result.reserved)


def test_framing_guess_is_verified_against_the_line_count(eng, O):
    """The warp engine guesses a tile's record framing from the text ('@' line whose second successor
    starts with '+') and checks the guess against the global line count.  Here every sequence starts
    with '+' and every quality string with '@', so the guess is wrong for chunks that begin inside a
    record; the output must still be that of plain line counting (common.rs:106-112)."""
    rng = random.Random(99)
    sheet, bcs = G.make_sheet(5, 12, 8, umi=0)
    recs = []
    for i in range(6000):
        n = rng.randrange(20, 260)
        seq = b"+" + bytes(rng.choice(b"ACGT") for _ in range(n))
        qual = b"@" + bytes(rng.choice(b"#+5?I@") for _ in range(n))
        recs.append(b"@read%d %d:N:0\n%s\n+\n%s\n" % (i, i % 7, seq, qual))
    data = b"".join(recs)
    assert len(data) > 40 * 16320
    check3(eng.trim_by_quality(data, 20), O.trim_by_quality(data, 20), "trim")
    assert eng.last_result.reserved == 1
    check3(eng.mask_by_quality(data, 20), O.mask_by_quality(data, 20), "mask")
    r1 = _bc_headers(data, bcs, 11)
    _cmp_demux(eng.demultiplex(sheet, r1, r1, fused_trim=20),
               O.demultiplex(sheet, O.trim_by_quality(r1, 20)[1], O.trim_by_quality(r1, 20)[1]), "fused")
    assert eng.last_result.reserved & 3 == 1  # (bit 3: compacted on the device)


def test_properties_at_bench_batch_size():
    """BASELINE configs[4] shape at a size no CPU oracle run fits into the test budget (2 M pairs, 384 samples,
    device-resident inputs, straight through the C ABI).  Size-independent properties: read conservation
    (fasta_demultiplex.rs:169,177-178), the fused quality trim changes bytes but never an assignment, every
    emitted record has one slice-table group, the trimmed output is never longer than the untrimmed one, and
    shard linearity -- the counters of two half ranges add up to those of the whole range, which is what the
    multi-GPU all-reduce relies on."""
    import ctypes as C
    import numpy as np
    import bench
    from seqkit_b200 import Engine, _lib as L

    n, S = 2_000_000, 384
    bcs = bench.make_sheet(S)
    with Engine(max_stream_bytes=n * 410 + (1 << 20), max_records=n, max_samples=S, aux_streams=False) as eng:
        lib = eng.lib
        eng.set_sheet(bcs)

        def run(first, cnt, fused):
            bench.synth_pair(eng, cnt, first)
            opts = L.DemuxOpts(20 if fused else -1, 0, 0, 0, 0)
            assert lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0, lib.sk_last_error(eng.ctx)
            res = eng.wait()
            assert res.status == 0 and res.n_records == cnt and res.reserved == 1, (res.status, res.n_records, res.reserved)
            counts = np.zeros(S + 2, dtype=np.uint64)
            assert lib.sk_download_counts(eng.ctx, 0, counts.ctypes.data) == 0
            assign = np.zeros(cnt, dtype=np.int16)
            assert lib.sk_download_assign(eng.ctx, 0, assign.ctypes.data, cnt) == 0
            eng.wait()
            ngroups = []
            for m in range(2):
                nc = res.n_chunks[m]
                rows = np.zeros(max(nc, 1) * 2, dtype=np.uint64)
                groups = np.zeros(cnt, dtype=np.uint32)
                assert lib.sk_download_demux_tables(eng.ctx, 0, m, rows.ctypes.data, groups.ctypes.data, cnt) == 0
                eng.wait()
                ng = (rows[1::2][:nc] >> np.uint64(32)).astype(np.int64)  # {u64 base, u32 first_group, u32 n_groups}
                ngroups.append(int(ng.sum()))
            return counts, assign, [int(res.out_extent[0]), int(res.out_extent[1])], ngroups

        c_plain, a_plain, ext_plain, g_plain = run(0, n, False)
        c_fused, a_fused, ext_fused, g_fused = run(0, n, True)
        for c, a, g in ((c_plain, a_plain, g_plain), (c_fused, a_fused, g_fused)):
            assert int(c[S]) == n and int(c[:S].sum()) == int(c[S + 1]) <= n
            assert int((a >= 0).sum()) == int(c[S + 1]) and a.min() >= -2 and a.max() < S
            assert np.array_equal(np.bincount(a[a >= 0], minlength=S).astype(np.uint64), c[:S])
            assert g == [int(c[S + 1])] * 2  # one group per emitted record and mate
        assert np.array_equal(a_plain, a_fused) and np.array_equal(c_plain, c_fused)
        assert 0.9 * n < int(c_plain[S + 1]) < n  # the generator leaves 2 % random barcodes and some double errors
        assert all(0 < f <= p for f, p in zip(ext_fused, ext_plain))
        half = n // 2
        c_lo, a_lo, _, _ = run(0, half, True)
        c_hi, a_hi, _, _ = run(half, n - half, True)
        assert np.array_equal(c_lo + c_hi, c_fused)
        assert np.array_equal(np.concatenate([a_lo, a_hi]), a_fused)


def test_more_records_than_the_context_was_sized_for(O):
    """sk_limits.max_records sizes the per-record tables (assign, umi, groups, record tables, the line engine's arrays).
    A batch with more records is refused with SK_DATA_TOO_MANY_RECORDS before anything is written past them -- on every
    engine -- and the context stays usable; operators without per-record tables (trim / mask by quality on the chunk
    engines) are not affected."""
    from seqkit_b200 import Engine
    from seqkit_b200.engine import Unsupported
    sheet, bcs = G.make_sheet(3, 8, 8, umi=4)
    r1, r2 = G.clean_pairs(5, 6000, bcs)
    small1, small2 = G.clean_pairs(6, 900, bcs)
    idx = G.index_reads(7, 6000, [b"ACGTACGT"])
    with Engine(max_stream_bytes=8 << 20, max_records=1000, max_samples=16, line_ops=True) as eng:
        for env_general in (False, True):
            if env_general:
                os.environ["SK_NO_WARP"] = "1"
            try:
                with Engine(max_stream_bytes=8 << 20, max_records=1000, max_samples=16, line_ops=True) as e2:
                    for call in (lambda: e2.demultiplex(sheet, r1, r2), lambda: e2.demultiplex(sheet, r1, r2, fused_trim=20),
                                 lambda: e2.add_barcode(r1, idx), lambda: e2.statistics(r1)):
                        with pytest.raises(Unsupported, match="max_records"):
                            call()
                    _cmp_demux(e2.demultiplex(sheet, small1, small2), O.demultiplex(sheet, small1, small2), "after a refusal")
                    check3(e2.trim_by_quality(r1, 20), O.trim_by_quality(r1, 20), "trim has no per-record table")
                    idx1k = G.index_reads(8, 1000, [b"ACGTACGT"])
                    check3(e2.add_barcode(small1, idx1k), O.add_barcode(small1, idx1k), "barcode file longer than the reads")
                    with pytest.raises(Unsupported, match="max_records"):  # the limit holds for every stream of a batch
                        e2.add_barcode(small1, idx)
            finally:
                os.environ.pop("SK_NO_WARP", None)
        # what INTEGRATION.md tells a byte-filling batcher to do: the same batch again with rec_limit = err_record
        import ctypes as C
        from seqkit_b200 import _lib as L
        eng.set_sheet(bcs)
        eng.upload(L.IN_R1, r1)
        eng.upload(L.IN_R2, r2)
        opts = L.DemuxOpts(-1, 0, 0, 0, 0)
        assert eng.lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0
        res = eng.wait()
        assert res.status == L.DATA_TOO_MANY_RECORDS and res.err_record == 1000
        opts.rec_limit = res.err_record
        assert eng.lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0
        res = eng.wait()
        lines1 = r1.split(b"\n")
        assert res.status == 0 and res.n_records == 1000 and res.consumed[0] == len(b"\n".join(lines1[:4000])) + 1
        # the line engine's own arrays: dense records take it there
        tiny = b"".join(b"@t BC:" + bcs[0][:8] + b"ACGT\n\n+\n\n" for _ in range(3000))
        with pytest.raises(Unsupported, match="max_records"):
            eng.demultiplex(sheet, tiny)
        with pytest.raises(Unsupported, match="max_records"):
            eng.add_barcode(tiny, idx)
