"""CPU tests of the parity oracle (no GPU): hand-derived KATs from the reference source
(SURVEY.md section 8c), the committed golden fixtures, and a differential fuzz of the two independent
restatements (oracle/fasta_oracle.c vs oracle/restatement.py)."""
import random

import pytest

import fuzzgen as G
import golden_util as GU
from oracle import pyoracle as O
from oracle import restatement as R


def fq(qual, seq=None, hdr=b"@r"):
    seq = seq if seq is not None else b"A" * len(qual)
    return hdr + b"\n" + seq + b"\n+\n" + qual + b"\n"


# kept length at min_baseq=20; I=Q40 5=Q20 4=Q19 #=Q2  (fasta_trim_by_quality.rs:28-48)
TRIM_KATS = [(b"IIIIIIIIII", 10), (b"IIIIIII###", 7), (b"##########", 0), (b"IIII#IIII#", 9), (b"III#I#I#I#", 9),
             (b"IIIIII5555", 10), (b"IIIIIIII4", 8), (b"I", 1), (b"#", 0), (b"", 0)]


@pytest.mark.parametrize("impl", [O, R], ids=["c", "py"])
@pytest.mark.parametrize("qual,kept", TRIM_KATS)
def test_trim_kats(impl, qual, kept):
    seq = bytes(b"ACGT"[i % 4] for i in range(len(qual)))
    code, out, err = impl.trim_by_quality(fq(qual, seq), 20)
    assert code == 0 and err == b""
    if kept == 0:
        assert out == b"@r\nN\n+\n!\n"
    else:
        assert out == b"@r\n" + seq[:kept] + b"\n+\n" + qual[:kept] + b"\n"


@pytest.mark.parametrize("impl", [O, R], ids=["c", "py"])
def test_mask_kat(impl):
    # 'I' keep; '#'->N; '5'=Q20 is not < 20; ' ' (byte 32) wraps to 255, keep; '4'=Q19 -> N
    code, out, err = impl.mask_by_quality(b"@h x\nACGTA\n+junk\nI#5 4\n", 20)
    assert (code, out, err) == (0, b"@h x\nANGTN\n+\nI#5 4\n", b"")


@pytest.mark.parametrize("impl", [O, R], ids=["c", "py"])
def test_error_paths(impl):
    # non-'@' header: everything before it is still emitted, status 255
    code, out, err = impl.trim_by_quality(fq(b"IIII") + b"\n", 20)
    assert code == 255 and out == b"@r\nAAAA\n+\nIIII\n" and err == b"ERROR: Invalid FASTQ format encountered.\n"
    code, out, err = impl.mask_by_quality(fq(b"IIII") + b"@x\nAC\n+\nIII\n", 20)
    assert code == 255 and out == b"@r\nAAAA\n+\nIIII\n"
    assert err == b"ERROR: Read sequence and base qualities are of different length.\n"
    # seq shorter than the kept quality prefix -> slice panic (status 101) after the header was printed
    code, out, err = impl.trim_by_quality(b"@p\nAC\n+\nIIIII\n", 20)
    assert code == 101 and out == b"@p\n" and err
    # truncated record: missing lines read as empty strings
    assert impl.trim_by_quality(b"@t", 20)[:2] == (0, b"@tN\n+\n!\n")
    assert impl.mask_by_quality(b"@t\nACG", 20)[0] == 255
    assert impl.mask_by_quality(b"@t", 20)[:2] == (0, b"@t\n+\n\n")
    # invalid UTF-8 -> I/O error
    code, out, err = impl.trim_by_quality(b"@r\nAC\n+\n\xff\xfe\n", 20)
    assert code == 255 and err == b"ERROR: I/O error while reading from file.\n"
    # the offending line is quoted whole, whatever its length (error! formats the String, common.rs:11-16)
    long_line = b"x" * 5000
    res = impl.demultiplex(b"A\tACGT\n", long_line + b"\nAC\n+\nII\n")
    assert res["exit_code"] == 255 and res["stderr"].endswith(b"ERROR: Invalid FASTQ header line:\n" + long_line + b"\n\n")
    code, out, err = impl.add_barcode(long_line + b"\nAC\n", b"")
    assert code == 255 and err == b"ERROR: Invalid FASTQ line:\n" + long_line + b"\n\n"


@pytest.mark.parametrize("impl", [O, R], ids=["c", "py"])
def test_demux_kats(impl):
    sheet = b"A\tACGTACGT\nB\tACGTACGA\n"
    rec = lambda name, bc: b"@%s BC:%s\nAC\n+\nII\n" % (name, bc)
    res = impl.demultiplex(sheet, rec(b"r1", b"ACGTACGC") + rec(b"r2", b"ACGTACGT") + rec(b"r3", b"NCGTACGT"))
    assert res["exit_code"] == 0 and res["counts"] == [2, 0] and res["total"] == 3 and res["identified"] == 2
    assert res["files"]["A.fq.gz"] == b"@r2\nAC\n+\nII\n@r3\nAC\n+\nII\n" and res["files"]["B.fq.gz"] == b""
    assert b"WARNING: Sequenced barcode ACGTACGC was an equally good match (1 mismatches) for samples A (ACGTACGT) and B (ACGTACGA)" in res["stderr"]
    assert res["stderr"].endswith(b"2 / 3 (66.7%) clusters carried a barcode matching one of the provided samples.\n")
    # UMI at 'U' positions of the sheet barcode, appended to both mates
    res = impl.demultiplex(b"X\tACGTUUUU\n", rec(b"u1 1:N", b"ACGTTTGA"), b"@u1 2:N BC:ACGTTTGA\nGG\n+\nII\n")
    assert res["files"]["X_1.fq.gz"] == b"@u1 1:N UMI:TTGA\nAC\n+\nII\n"
    assert res["files"]["X_2.fq.gz"] == b"@u1 2:N UMI:TTGA\nGG\n+\nII\n"
    # duplicate barcodes in the sheet -> every matching read is dropped (tie at distance 0)
    res = impl.demultiplex(b"P\tACGT\nQ\tACGT\n", rec(b"d", b"ACGT"))
    assert res["identified"] == 0 and res["files"] == {"P.fq.gz": b"", "Q.fq.gz": b""}
    # no BC field / wrong length / duplicate names
    assert impl.demultiplex(sheet, b"@x\nA\n+\nI\n")["stderr"].endswith(b"ERROR: No BC:xxxx field found.\n")
    assert impl.demultiplex(sheet, rec(b"x", b"ACG"))["exit_code"] == 255
    assert impl.demultiplex(b"A\tAC\nA\tGG\n", b"")["stderr"].endswith(b"ERROR: Sample A is listed multiple times in sample sheet.\n")
    # index-file route joins with '+', header untouched
    res = impl.demultiplex(b"S\tAC+GT\n", b"@r BC:zz\nAC\n+\nII\n", None, b"@i\nAC\n+\nII\n", b"@i\nGT\n+\nII\n")
    assert res["files"]["S.fq.gz"] == b"@r BC:zz\nAC\n+\nII\n"


@pytest.mark.parametrize("impl", [O, R], ids=["c", "py"])
def test_add_barcode_kats(impl):
    fqd = fq(b"II", b"AC", b"@a 1 ") + fq(b"II", b"GG", b"@b") + fq(b"II", b"TT", b"@c")
    bc = b"@x\nACGT+TTAA\n+\nIIIIIIIII\n@y\nGGGG\n+\nIIII\n"  # barcode file one record short: last barcode is reused
    code, out, err = impl.add_barcode(fqd, bc)
    assert code == 0
    assert out == b"@a 1 BC:ACGT+TTAA\nAC\n+\nII\n@b BC:GGGG\nGG\n+\nII\n@c BC:GGGG\nTT\n+\nII\n"
    code, out, err = impl.add_barcode(b">s\nACGT\nbad\n", b">b\nAA\n>b\nCC\n")
    assert code == 255 and out == b">s BC:AA\nACGT\nbad BC:CC\n" and err == b"ERROR: Invalid FASTQ line:\nbad\n\n"


def _golden_run(impl, c):
    op = c["op"]
    if op == "trim":
        return impl.trim_by_quality(GU.blob(c["input"]), c["min_baseq"])
    if op == "mask":
        return impl.mask_by_quality(GU.blob(c["input"]), c["min_baseq"])
    if op == "add_barcode":
        return impl.add_barcode(GU.blob(c["input"]), GU.blob(c["barcodes"]))
    raise AssertionError(op)


@pytest.mark.parametrize("impl", [O, R], ids=["c", "py"])
def test_golden_stream_ops(impl):
    n = 0
    for c in GU.cases():
        if c["op"] == "demux":
            continue
        code, out, err = _golden_run(impl, c)
        assert code == c["exit_code"], c["tag"]
        assert out == GU.blob(c["stdout"]), c["tag"]
        if code != 101:  # panic text is not part of the contract
            assert err == GU.blob(c["stderr"]), c["tag"]
        n += 1
    assert n > 50


@pytest.mark.parametrize("impl", [O, R], ids=["c", "py"])
def test_golden_demux(impl):
    for c in GU.cases("demux"):
        res = impl.demultiplex(GU.blob(c["sheet"]), GU.blob(c["r1"]), GU.blob(c["r2"]))
        assert res["exit_code"] == c["exit_code"], c["tag"]
        assert res["stderr"] == GU.blob(c["stderr"]), c["tag"]
        assert res["counts"] == c["counts"] and res["total"] == c["total"] and res["identified"] == c["identified"]
        assert set(res["files"]) == set(c["files"])
        for k, v in c["files"].items():
            assert res["files"][k] == GU.blob(v), (c["tag"], k)


def _mutate_bytes(rng, data):
    b = bytearray(data)
    for _ in range(rng.randrange(0, 6)):
        if not b:
            break
        i = rng.randrange(len(b))
        r = rng.random()
        if r < 0.3:
            b[i] = rng.choice(b"\n\r \t@+>ACGTN!#I5~\x00\x7f")
        elif r < 0.5:
            del b[i]
        elif r < 0.7:
            b.insert(i, rng.choice(b"\n \tACGT#I"))
        elif r < 0.8:
            b[i:i] = "é  ".encode("utf-8")  # valid multi-byte incl. Unicode whitespace
        elif r < 0.85:
            b[i] = rng.randrange(128, 256)  # usually invalid UTF-8
    return bytes(b)


def test_fuzz_c_vs_python_stream_ops():
    rng = random.Random(12345)
    for it in range(400):
        data = G.nasty_fastq(rng.randrange(1 << 30), rng.randrange(0, 12))
        if it % 2:
            data = _mutate_bytes(rng, data)
        q = rng.choice((0, 2, 10, 20, 30, 41, 93, 200, 255))
        for name in ("trim_by_quality", "mask_by_quality"):
            a = getattr(O, name)(data, q)
            b = getattr(R, name)(data, q)
            assert a[0] == b[0] and a[1] == b[1], (name, q, data)
            if a[0] != 101:
                assert a[2] == b[2], (name, q, data)
        bc = G.index_reads(rng.randrange(1 << 30), rng.randrange(0, 12), [b"ACGT", b"GG+TT"])
        if it % 3 == 0:
            bc = _mutate_bytes(rng, bc)
        a, b = O.add_barcode(data, bc), R.add_barcode(data, bc)
        assert a == b, (data, bc)


def test_fuzz_c_vs_python_demux():
    rng = random.Random(777)
    for it in range(150):
        S = rng.choice((1, 2, 3, 8, 20))
        L = rng.choice((4, 6, 8, 12))
        umi = rng.choice((0, 0, 3))
        sheet, bcs = G.make_sheet(rng.randrange(1 << 30), S, L, umi=umi, dual=rng.random() < 0.3,
                                  min_dist=rng.choice((0, 1, 2, 3)), wild_n=rng.choice((0, 0, 0.1)))
        n = rng.randrange(0, 25)
        if it % 2:
            r1, r2 = G.nasty_headers_pairs(rng.randrange(1 << 30), n, bcs)
        else:
            r1, r2 = G.clean_pairs(rng.randrange(1 << 30), n, bcs, p_sub=0.1, p_n=0.05, p_random=0.1, p_lower=0.02,
                                   bc_in_r2=rng.random() < 0.5)
        if it % 5 == 0:
            r1 = _mutate_bytes(rng, r1)
            r2 = _mutate_bytes(rng, r2)
        if it % 7 == 0:
            sheet = _mutate_bytes(rng, sheet)
        kw = {}
        if it % 11 == 0:
            half = [b[:len(b) // 2] for b in bcs]
            kw = {"index1": G.index_reads(it, n, half), "index2": G.index_reads(it + 1, n, [b[len(b) // 2 + 1:] for b in bcs])}
        for paired in (True, False):
            a = O.demultiplex(sheet, r1, r2 if paired else None, **kw)
            b = R.demultiplex(sheet, r1, r2 if paired else None, **kw)
            if a["exit_code"] == 101:
                assert b["exit_code"] == 101
                continue
            assert a == b, (sheet, r1, r2)
        if it % 13 == 0:
            a = O.demultiplex(sheet, r1, r2, dry_run=5)
            b = R.demultiplex(sheet, r1, r2, dry_run=5)
            assert a["exit_code"] == b["exit_code"] and a["counts"] == b["counts"]


def test_host_generator_shapes():
    """oracle/synth_host.c: deterministic, range-addressable (two half ranges concatenate to the whole), four
    lines per record, ' BC:' fields of the sheet's length."""
    import bench
    from oracle import pyoracle as O
    bcs = bench.make_sheet(24)
    a = O.synth_fastq(500, seed=5, first_pair=100, mate=1, barcodes=bcs)
    assert a == O.synth_fastq(500, seed=5, first_pair=100, mate=1, barcodes=bcs)
    assert a == O.synth_fastq(200, seed=5, first_pair=100, mate=1, barcodes=bcs) + \
        O.synth_fastq(300, seed=5, first_pair=300, mate=1, barcodes=bcs)
    lines = a.split(b"\n")
    assert len(lines) == 4 * 500 + 1 and all(len(x) == 150 for x in lines[1::4]) and all(x == b"+" for x in lines[2:-1:4])
    assert all(h.startswith(b"@SIM:1:FC:") and len(h[h.rfind(b" BC:") + 4:]) == 29 for h in lines[0:-1:4])
    b = O.synth_fastq(500, seed=5, first_pair=100, mate=2, barcodes=bcs)
    h1, h2 = lines[0:-1:4], b.split(b"\n")[0:-1:4]
    assert all(x.replace(b" 1:N", b" 2:N") == y for x, y in zip(h1, h2))  # mates share coordinates and barcode


def test_next_row_operators_kats():
    """Hand-derived expectations for the SURVEY 8(f) restatements (fasta_trim.rs, fasta_check.rs, fasta_statistics.rs,
    fasta_interleave.rs, fasta_deinterleave.rs, fasta_extract_dual_umi.rs)."""
    from oracle import pyoracle as O
    fq = b"@r1 x\nACGTACGT\n+r1\nIIIIIIII\n@r2\nAC\n+\nII\n"
    # trim: first+last < seq_len cuts both lines; otherwise both lines come out empty; '+' line normalised
    assert O.next_op(0, fq, x=2, y=1)[:3] == (0, b"@r1 x\nGTACG\n+\nIIIII\n@r2\n\n+\n\n", b"")
    assert O.next_op(0, b">s\nACGT\n", x=1)[:3] == (0, b">s\nCGT\n", b"")
    assert O.next_op(0, b"x\nAC\n")[:3] == (255, b"", b"ERROR: Invalid FASTA/FASTQ format encountered.\n")
    code, out, err, _ = O.next_op(0, b"@q\nACGTACGT\n+\nIII\n", x=1)  # &qual[1..8] on a 4-byte line: panic after the first print
    assert code == 101 and out == b"@q\nCGTACGT\n" and err
    # check: line number of the offending line and the (up to ten) lines read so far, each followed by a blank line
    assert O.next_op(1, fq)[:3] == (0, b"", b"")
    assert O.next_op(1, b"@r1\nAC\nX\nII\n")[:3] == (255, b"", b"ERROR: Missing quality header prefix '+' on line 3:\n@r1\n\nAC\n\nX\n\n\n\n")
    assert O.next_op(1, b">s\nAC\nzz\n")[:3] == (255, b"", b"ERROR: Missing header prefix '>' or '@' on line 3:\n>s\n\nAC\n\nzz\n\n\n\n")
    assert O.next_op(1, b"@r1\nAC\n")[2] == b"ERROR: Missing quality header prefix '+' on line 2:\n@r1\n\nAC\n\n\n\n"  # EOF: nothing more was read
    # statistics: fewer than 100 distinct barcodes -> the two header lines, then the slice panic
    code, out, err, _ = O.next_op(2, b"@a BC:ACGT+TT\nA\n+\nI\n>b BC:acgt\nA\n")
    assert code == 101 and out == b"Total sequence records: 2\nMost frequent sample barcodes:\n"
    many = b"".join(b"@r BC:%s\nA\n+\nI\n" % (bytes(b"ACGT"[(i >> s) & 3] for s in (0, 2, 4, 6))) * (1 + i % 3) for i in range(128))
    code, out, err, _ = O.next_op(2, many)
    lines = out.split(b"\n")
    assert code == 0 and lines[0] == b"Total sequence records: 255" and len(lines) == 103
    counts = [int(x.rsplit(b": ", 1)[1]) for x in lines[2:102]]
    assert counts == sorted(counts, reverse=True) and counts[0] == 3
    # interleave / deinterleave / extract dual umi
    a, b = b"@a/1\nAC\n+\nII\n@b/1\nGG\n+\nII\n", b"@a/2\nTT\n+\nII\n@b/2\nCC\n+\nII\n"
    il = O.next_op(3, a, b)
    assert il[:3] == (0, b"@a/1\nAC\n+\nII\n@a/2\nTT\n+\nII\n@b/1\nGG\n+\nII\n@b/2\nCC\n+\nII\n", b"")
    assert O.next_op(4, il[1]) == (0, a, b"", b)
    assert O.next_op(3, a, b[:13])[:3] == (255, a[:13] + b[:13] + a[13:], b"ERROR: Input files do not share a consistent format.\n")
    assert O.next_op(5, il[1], x=1)[1] == (b"@a/1 RX:A+T\nC\n+\nI\n@a/2 RX:A+T\nT\n+\nI\n@b/1 RX:G+C\nG\n+\nI\n@b/2 RX:G+C\nC\n+\nI\n")
    assert O.next_op(5, a[:13], x=0)[:3] == (255, b"", b"ERROR: Invalid FASTQ record found in input file.\n")


def test_fuzz_c_vs_python_next_rows():
    """The SURVEY 8(f) operators -- trim --first/--last, check, statistics, interleave, deinterleave, extract dual umi -- in
    the two independent restatements (C: fasta_oracle.c, Python: restatement.py) on clean, FASTA, interleaved, UTF-8 and
    malformed inputs: same exit code, stdout, second output, and stderr unless the reference would panic."""
    rng = random.Random(99)

    def same(op, a, b=None, x=0, y=0, ctx=None):
        c, p = O.next_op(op, a, b, x, y), R.next_op(op, a, b, x, y)
        assert c[0] == p[0], (ctx, op, c[0], p[0], c[2][-200:], p[2][-200:])
        assert c[1] == p[1] and c[3] == p[3], (ctx, op)
        if c[0] != 101:
            assert c[2] == p[2], (ctx, op, c[2][-200:], p[2][-200:])
        return c

    def fasta(seed, n):
        r = random.Random(seed)
        return b"".join(b">s%d d\n" % i + G.rand_seq(r, r.randrange(0, 60)) + b"\n" for i in range(n))

    sheet, bcs = G.make_sheet(3, 150, 8)
    for it in range(60):
        n = rng.choice((0, 1, 2, 7, 60, 400))
        p1, p2 = G.clean_pairs(rng.randrange(1 << 30), n, bcs)
        inter = same(3, p1, p2, ctx="interleave")[1]
        variants = [p1, G.nasty_fastq(rng.randrange(1 << 30), n), fasta(it, n), inter,
                    "@ré 日\nACGTé\n+\nIIIIII\n".encode() * 3, p1 + b"oops\n" + p2, p1[: len(p1) // 2], b"\n", b"@a", b">x\nAC"]
        for data in variants:
            same(0, data, None, rng.randrange(0, 9), rng.randrange(0, 9), ctx="trim fixed")
            same(1, data, ctx="check")
            same(2, data, ctx="statistics")
            same(3, data, rng.choice(variants), ctx="interleave")
            same(4, data, ctx="deinterleave")
            same(5, data, None, rng.randrange(0, 7), ctx="dual umi")
    # statistics with at least 100 distinct barcodes, ties included (tie order: barcode descending, section 2 of DESIGN.md)
    p1, _ = G.clean_pairs(5, 3000, bcs, p_sub=0.2)
    code, out, err, _ = same(2, p1, ctx="statistics table")
    assert code == 0 and out.count(b"\n") == 102


@pytest.mark.parametrize("impl", [O, R], ids=["c", "py"])
def test_golden_next_rows(impl):
    """Both restatements reproduce the committed fixtures of the SURVEY 8(f) operators (tests/golden/golden_next.json)."""
    n = 0
    for c in GU.next_cases():
        got = impl.next_op(c["op"], GU.next_blob(c["a"]), GU.next_blob(c["b"]), c["x"], c["y"])
        assert got[0] == c["exit_code"], c["tag"]
        assert got[1] == GU.next_blob(c["stdout"]) and got[3] == GU.next_blob(c["out2"]), (c["op"], c["tag"])
        if got[0] != 101:
            assert got[2] == GU.next_blob(c["stderr"]), (c["op"], c["tag"])
        n += 1
    assert n > 90
