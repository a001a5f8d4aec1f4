"""Drop-in check of the `fasta` host binary (-m gpu): same argv, same stdout bytes, same decompressed
output files and names, same stderr lines and exit status as the CPU oracle's CLI, on the same
seeded inputs -- plain, gzip-compressed and stdin inputs, single- and multi-batch (SK_BATCH_MB=1)."""
import gzip
import os
import random
import subprocess

import pytest

import fuzzgen as G

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FASTA = os.path.join(ROOT, "seqkit_b200", "fasta")


@pytest.fixture(scope="module")
def oracle_bin():
    from oracle import pyoracle
    pyoracle.build()
    p = os.path.join(ROOT, "oracle", "_build", "fasta_oracle")
    assert os.path.exists(p)
    return p


@pytest.fixture(scope="module", autouse=True)
def host_bin():
    if not os.path.exists(FASTA):
        subprocess.check_call(["make", "-s", "-C", ROOT, "seqkit_b200/fasta"])
    return FASTA


def run(binary, args, cwd, stdin=None, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([binary] + args, cwd=cwd, input=stdin, capture_output=True, env=e, timeout=300)
    return p.returncode, p.stdout, p.stderr


def same(a, b, ctx=None):
    assert a[0] == b[0], (ctx, a[0], b[0], a[2][-400:], b[2][-400:])
    assert a[1] == b[1], ctx
    if a[0] != 101:  # the text of a Rust panic is rustc-version dependent; only the status is contractual
        assert a[2] == b[2], (ctx, a[2][-400:], b[2][-400:])
    else:
        assert a[2] and b[2]


def gz_files(d):
    out = {}
    for f in sorted(os.listdir(d)):
        if f.endswith(".fq.gz"):
            out[f] = gzip.decompress(open(os.path.join(d, f), "rb").read())
    return out


def both(oracle_bin, tmp_path, args, files, stdin=None, env=None, ctx=None):
    res = []
    for tag, binary in (("ours", FASTA), ("oracle", oracle_bin)):
        d = tmp_path / tag
        d.mkdir(exist_ok=True)
        for f in os.listdir(d):
            os.remove(os.path.join(d, f))
        for name, data in files.items():
            if name.endswith(".gz"):
                data = gzip.compress(data)
            (d / name).write_bytes(data)
        res.append((run(binary, args, str(d), stdin, env if tag == "ours" else None), gz_files(str(d))))
    same(res[0][0], res[1][0], ctx)
    assert res[0][1] == res[1][1], ctx
    return res[0]


@pytest.mark.parametrize("op", ["trim", "mask"])
def test_stream_ops(oracle_bin, tmp_path, op):
    rng = random.Random(7)
    for it in range(10):
        n = rng.choice((0, 1, 50, 900))
        data = G.nasty_fastq(rng.randrange(1 << 30), n, fatal_ok=it % 3 == 0) if it % 2 else G.clean_fastq(it, n)
        q = str(rng.choice((0, 2, 20, 30, 41)))
        both(oracle_bin, tmp_path, [op, "by", "quality", "in.fq", q], {"in.fq": data}, ctx=(op, it))
    data = G.clean_fastq(5, 400)
    both(oracle_bin, tmp_path, [op, "by", "quality", "in.fq.gz", "20"], {"in.fq.gz": data}, ctx="gz")
    both(oracle_bin, tmp_path, [op, "by", "quality", "-", "20"], {}, stdin=data, ctx="stdin")


def test_stream_ops_multi_batch(oracle_bin, tmp_path):
    data = G.clean_fastq(11, 12000, qual_style="mix")  # ~2.5 MB -> several 1 MiB batches, two slots in flight
    assert len(data) > 2 << 20
    for op in ("trim", "mask"):
        both(oracle_bin, tmp_path, [op, "by", "quality", "in.fq", "20"], {"in.fq": data}, env={"SK_BATCH_MB": "1", "SK_GPUS": "8"}, ctx=op)
    bad = data + b"Xbroken\nACGT\n+\nIIII\n" + G.clean_fastq(12, 10)  # fatal in the last batch: earlier output stays
    for op in ("trim", "mask"):
        both(oracle_bin, tmp_path, [op, "by", "quality", "in.fq", "20"], {"in.fq": bad}, env={"SK_BATCH_MB": "1"}, ctx=op)


def test_argument_and_file_errors(oracle_bin, tmp_path):
    fq = G.clean_fastq(1, 3)
    both(oracle_bin, tmp_path, ["trim", "by", "quality", "missing.fq", "20"], {})
    both(oracle_bin, tmp_path, ["mask", "by", "quality", "in.fq", "300"], {"in.fq": fq})  # u8 parse -> panic status
    both(oracle_bin, tmp_path, ["trim", "by", "quality", "in.fq", "+7"], {"in.fq": fq})
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "missing.fq"], {"sheet.tsv": b"a\tACGT\n"})
    both(oracle_bin, tmp_path, ["demultiplex", "--dry-run=0", "sheet.tsv", "in.fq"], {"sheet.tsv": b"a\tACGT\n", "in.fq": fq})
    for args in (["trim", "by", "quality", "in.fq"], ["add", "barcode", "in.fq"], ["demultiplex", "sheet.tsv"]):
        a = run(FASTA, args, str(tmp_path))
        assert a[0] == 255 and a[2].startswith(b"ERROR: Invalid arguments.\n")
    a = run(FASTA, ["frobnicate"], str(tmp_path))
    assert a[0] == 0 and a[1] == b"" and b"fasta trim by quality <fastq_file> <min_baseq>" in a[2]


def test_add_barcode(oracle_bin, tmp_path):
    rng = random.Random(3)
    for it in range(6):
        n = rng.choice((1, 40, 700))
        reads = G.clean_fastq(100 + it, n)
        nb = n if it % 3 else max(1, n // 2)  # fewer barcode records: the last one is reused
        bc = G.index_reads(200 + it, nb, [b"ACGTACGTAC+TTGCATTGCA", b"GGGTTTAAAC+CCCAAATTTG"])
        both(oracle_bin, tmp_path, ["add", "barcode", "in.fq", "bc.fq"], {"in.fq": reads, "bc.fq": bc}, ctx=it)
    reads = G.clean_fastq(9, 9000)
    bc = G.index_reads(10, 4000, [b"ACGTACGT", b"TTTTCCCC"])
    both(oracle_bin, tmp_path, ["add", "barcode", "in.fq", "bc.fq.gz"], {"in.fq": reads, "bc.fq.gz": bc},
         env={"SK_BATCH_MB": "1"}, ctx="multi-batch + reuse")
    fa = b"".join(b">s%d\nACGTTGCA\n" % i for i in range(50))
    both(oracle_bin, tmp_path, ["add", "barcode", "in.fa", "bc.fa"], {"in.fa": fa, "bc.fa": fa}, ctx="fasta")
    both(oracle_bin, tmp_path, ["add", "barcode", "in.fq", "bc.fq"],
         {"in.fq": G.clean_fastq(1, 5) + b"oops\nA\n+\nI\n", "bc.fq": bc}, ctx="bad line")


def test_demultiplex(oracle_bin, tmp_path):
    rng = random.Random(21)
    for it in range(8):
        S = rng.choice((2, 8, 24, 96))
        sheet, bcs = G.make_sheet(it, S, rng.choice((8, 12, 20)), umi=rng.choice((0, 8)), dual=it % 2 == 0,
                                  min_dist=rng.choice((1, 3)))
        n = rng.choice((0, 1, 200, 1500))
        if it % 3 == 2:
            r1, r2 = G.nasty_headers_pairs(300 + it, n, bcs)
        else:
            r1, r2 = G.clean_pairs(300 + it, n, bcs, p_sub=0.05, p_n=0.02, p_random=0.1)
        files = {"sheet.tsv": sheet, "r1.fq": r1, "r2.fq": r2}
        both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq", "r2.fq"], files, ctx=("paired", it))
        both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq"], files, ctx=("single", it))


def test_demultiplex_multi_batch_gz_and_errors(oracle_bin, tmp_path):
    sheet, bcs = G.make_sheet(5, 48, 20, umi=8, dual=True)
    r1, r2 = G.clean_pairs(77, 9000, bcs, p_sub=0.03, p_random=0.05)
    files = {"sheet.tsv": sheet, "r1.fastq.gz": r1, "r2.fastq.gz": r2}
    ours, out_files = both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fastq.gz", "r2.fastq.gz"], files,
                           env={"SK_BATCH_MB": "1"}, ctx="multi-batch")
    assert len(out_files) == 96 and sum(len(v) for v in out_files.values()) > 1 << 20
    # a read without a BC field in a later batch: everything before it is written, then the reference's error
    bad1 = r1 + b"@nobc 1:N:0\nACGT\n+\nIIII\n"
    bad2 = r2 + b"@nobc 2:N:0\nACGT\n+\nIIII\n"
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq", "r2.fq"],
         {"sheet.tsv": sheet, "r1.fq": bad1, "r2.fq": bad2}, env={"SK_BATCH_MB": "1"}, ctx="no BC")
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq"],
         {"sheet.tsv": sheet + b"S001\tACGT\n", "r1.fq": r1}, ctx="sheet error after files were created")
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq"],
         {"sheet.tsv": sheet, "r1.fq": r1[:5000] + b"@x BC:ACGT\nA\n+\nI\n"}, ctx="barcode length")


def test_demultiplex_index_route_and_dry_run(oracle_bin, tmp_path):
    sheet, bcs = G.make_sheet(9, 12, 16, umi=0, dual=True)
    n = 800
    r1, r2 = G.clean_pairs(31, n, bcs, bc_in_r2=False)
    i1 = G.index_reads(32, n, [b.split(b"+")[0] for b in bcs], p_sub=0.03)
    i2 = G.index_reads(33, n, [b.split(b"+")[1] for b in bcs], p_sub=0.03)
    files = {"sheet.tsv": sheet, "r1.fq": r1, "r2.fq": r2, "i1.fq": i1, "i2.fq": i2}
    both(oracle_bin, tmp_path, ["demultiplex", "--index1=i1.fq", "--index2=i2.fq", "sheet.tsv", "r1.fq", "r2.fq"], files,
         ctx="index route")
    # dry run: >= 100 table entries (12 samples + many unmatched barcodes) ...
    sheet2, bcs2 = G.make_sheet(4, 12, 8)
    d1, _ = G.clean_pairs(41, 3000, bcs2, p_random=0.5)
    ours, _ = both(oracle_bin, tmp_path, ["demultiplex", "--dry-run=2500", "sheet.tsv", "r1.fq"],
                   {"sheet.tsv": sheet2, "r1.fq": d1}, env={"SK_BATCH_MB": "1"}, ctx="dry run")
    assert ours[0] == 0 and ours[1].count(b"\n") == 100
    # ... and the reference's panic when there are fewer than 100
    both(oracle_bin, tmp_path, ["demultiplex", "--dry-run=10", "sheet.tsv", "r1.fq"], {"sheet.tsv": sheet2, "r1.fq": d1},
         ctx="dry run, short table")


def test_fused_trim_extension_equals_shell_composition(oracle_bin, tmp_path):
    """--trim-by-quality=Q (extension) == fasta demultiplex sheet <(fasta trim by quality R1 Q) <(... R2 Q)."""
    sheet, bcs = G.make_sheet(2, 24, 20, umi=8, dual=True)
    r1, r2 = G.clean_pairs(55, 2500, bcs)
    d = tmp_path / "o"
    d.mkdir()
    for name, data in (("sheet.tsv", sheet), ("r1.fq", r1), ("r2.fq", r2)):
        (d / name).write_bytes(data)
    for m in ("1", "2"):
        t = run(oracle_bin, ["trim", "by", "quality", "r%s.fq" % m, "20"], str(d))
        assert t[0] == 0
        (d / ("t%s.fq" % m)).write_bytes(t[1])
    want = run(oracle_bin, ["demultiplex", "sheet.tsv", "t1.fq", "t2.fq"], str(d))
    want_files = gz_files(str(d))
    e = tmp_path / "g"
    e.mkdir()
    for name, data in (("sheet.tsv", sheet), ("r1.fq", r1), ("r2.fq", r2)):
        (e / name).write_bytes(data)
    got = run(FASTA, ["demultiplex", "--trim-by-quality=20", "sheet.tsv", "r1.fq", "r2.fq"], str(e))
    same(got, want)
    assert gz_files(str(e)) == want_files


def test_errors_in_early_and_middle_batches(oracle_bin, tmp_path):
    """A failing record, an ambiguous read and a dry run whose interesting record is the FIRST record of a
    middle batch of a multi-batch run (SK_BATCH_MB=1): the header text quoted in the messages and the dry-run
    walk come from the batch's own pinned buffer, which later batches must not have touched; and a .gz input
    that fails in its first batch must still exit (no gunzip child left blocked on a full pipe)."""
    sheet, bcs = G.make_sheet(5, 12, 8, umi=0, min_dist=1)
    r1, r2 = G.clean_pairs(77, 16000, bcs, p_sub=0.0, p_n=0.0, p_random=0.02, read_len=(100, 150))
    assert len(r1) > 4 << 20
    recs = r1.split(b"\n")
    # byte offset of the first record that starts beyond 2 MiB: with 1 MiB batches it sits in a middle batch
    pos, k = 0, 0
    while pos < (2 << 20) + 1000:
        pos += sum(len(x) + 1 for x in recs[4 * k:4 * k + 4])
        k += 1
    head, tail = b"\n".join(recs[:4 * k]) + b"\n", b"\n".join(recs[4 * k:])
    env = {"SK_BATCH_MB": "1"}
    bad = head + b"@nobc 1:N:0:1\nACGT\n+\nIIII\n" + tail
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq"], {"sheet.tsv": sheet, "r1.fq": bad}, env=env, ctx="no BC, middle")
    bad = head + b"Xnot a header BC:" + bcs[0] + b"\nACGT\n+\nIIII\n" + tail
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq"], {"sheet.tsv": sheet, "r1.fq": bad}, env=env, ctx="bad header, middle")
    bad = head + b"@short BC:ACG\nACGT\n+\nIIII\n" + tail
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq"], {"sheet.tsv": sheet, "r1.fq": bad}, env=env, ctx="barcode length, middle")
    # ambiguous reads (two sheet rows one mismatch apart on either side of the observed barcode)
    amb_sheet = b"A\tACGTACGT\nB\tACGTACGA\n" + sheet
    amb = head + b"@amb 1:N:0:1 BC:ACGTACGC\nACGT\n+\nIIII\n" + tail
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq"], {"sheet.tsv": amb_sheet, "r1.fq": amb}, env=env, ctx="ambiguous, middle")
    # dry run that covers three batches and a bit
    both(oracle_bin, tmp_path, ["demultiplex", "--dry-run=%d" % (k + 50), "sheet.tsv", "r1.fq"],
         {"sheet.tsv": G.make_sheet(4, 120, 8)[0], "r1.fq": r1}, env=env, ctx="dry run, middle")
    # .gz input, failure in the very first batch of a multi-batch file
    bad = b"@nobc 1:N:0:1\nACGT\n+\nIIII\n" + r1
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq.gz"], {"sheet.tsv": sheet, "r1.fq.gz": bad}, env=env, ctx="gz, first batch")
    for op in ("trim", "mask"):
        bad = G.clean_fastq(3, 10) + b"Xbroken\nACGT\n+\nIIII\n" + G.clean_fastq(11, 20000, qual_style="mix")
        both(oracle_bin, tmp_path, [op, "by", "quality", "in.fq.gz", "20"], {"in.fq.gz": bad}, env=env, ctx=(op, "gz early"))


def test_gzip_children_and_per_record_tables_give_the_same_files(oracle_bin, tmp_path):
    """SK_GZIP=child (one `gzip -c` child per file, the reference's GzipWriter) and SK_NO_COMPACT=1 (per-record
    slice tables instead of the device-side compaction) are the same drop-in: decompressed files identical."""
    sheet, bcs = G.make_sheet(6, 24, 20, umi=8, dual=True)
    r1, r2 = G.clean_pairs(78, 5000, bcs, p_sub=0.03, p_random=0.05)
    files = {"sheet.tsv": sheet, "r1.fq": r1, "r2.fq": r2}
    for env in ({"SK_GZIP": "child"}, {"SK_NO_COMPACT": "1"}, {"SK_GZIP": "child", "SK_NO_COMPACT": "1", "SK_BATCH_MB": "1"},
                {"SK_GZIP_THREADS": "3", "SK_BATCH_MB": "1"}):
        both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq", "r2.fq"], files, env=env, ctx=env)


def test_all_visible_gpus_give_the_single_gpu_bytes(oracle_bin, tmp_path):
    """Batches are dealt round-robin to the run's GPUs and consumed in batch order: the output does not depend on
    the number of GPUs (SK_GPUS=8 = every visible GPU, eight at most; one GPU visible: the same code path with one
    lane group).  Without SK_GPUS a run takes one GPU per 4 GiB of plain input."""
    sheet, bcs = G.make_sheet(7, 48, 20, umi=8, dual=True)
    r1, r2 = G.clean_pairs(79, 12000, bcs, p_sub=0.03, p_random=0.05)
    files = {"sheet.tsv": sheet, "r1.fq": r1, "r2.fq": r2}
    res = []
    for env in ({"SK_BATCH_MB": "1", "SK_GPUS": "8"}, {"SK_BATCH_MB": "1", "SK_GPUS": "1"}, {"SK_BATCH_MB": "1", "SK_DEVICE": "0"}):
        res.append(both(oracle_bin, tmp_path, ["demultiplex", "--trim-by-quality=20", "sheet.tsv", "r1.fq", "r2.fq"]
                        if False else ["demultiplex", "sheet.tsv", "r1.fq", "r2.fq"], files, env=env, ctx=env))
    assert res[0][1] == res[1][1] == res[2][1]
    data = G.clean_fastq(21, 30000, qual_style="mix")
    for op in ("trim", "mask"):
        both(oracle_bin, tmp_path, [op, "by", "quality", "in.fq", "20"], {"in.fq": data}, env={"SK_BATCH_MB": "1", "SK_GPUS": "8"}, ctx=op)


def test_demultiplex_long_and_utf8_records(oracle_bin, tmp_path):
    """Records no chunk engine frames -- 20 kb reads, UTF-8 and Unicode white space in headers -- go through the line
    engine inside the same `fasta demultiplex` run, batch by batch (fasta_demultiplex.rs:117-249)."""
    rng = random.Random(5)
    sheet, bcs = G.make_sheet(9, 12, 8, umi=4)
    p1, p2 = G.clean_pairs(41, 4000, bcs)
    extra1, extra2 = [], []
    for i in range(40):
        bc = G.observed_barcode(rng, bcs, p_sub=0.03)
        n = rng.choice((60, 9000, 20000))
        name = ("né%d 日" % i).encode() if i % 2 else b"long%d" % i
        tail = rng.choice((b"", "  ".encode(), b" z:9"))
        sq, q = G.rand_seq(rng, n), bytes(rng.choice(b"#5II") for _ in range(n))
        extra1.append(b"@" + name + b" 1 BC:" + bc + tail + b"\n" + sq + b"\n+\n" + q + b"\n")
        extra2.append(b"@" + name + b" 2 BC:" + bc + b"\n" + sq[::-1] + b"\n+\n" + q + b"\n")
    files = {"sheet.tsv": sheet, "r1.fq": p1 + b"".join(extra1) + p1, "r2.fq": p2 + b"".join(extra2) + p2}
    both(oracle_bin, tmp_path, ["demultiplex", "sheet.tsv", "r1.fq", "r2.fq"], files, env={"SK_BATCH_MB": "1"}, ctx="long+utf8")
