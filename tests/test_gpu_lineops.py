"""SURVEY.md section 8(f) rows on the line engine (-m gpu): `fasta trim --first/--last`, `fasta check`, `fasta statistics`,
`fasta interleave`, `fasta deinterleave`, `fasta extract dual umi` -- the `fasta` binary against the oracle's CLI
(same argv, stdout bytes, decompressed output files, stderr, exit status) and the raw C-ABI operator
(sk_line_op) against the oracle's functions, on clean, nasty, FASTA, truncated and multi-batch inputs."""
import collections
import gzip
import os
import random
import re
import subprocess

import pytest

import fuzzgen as G

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FASTA = os.path.join(ROOT, "seqkit_b200", "fasta")


@pytest.fixture(scope="module")
def oracle_bin():
    from oracle import pyoracle
    pyoracle.build(force=True)
    return os.path.join(ROOT, "oracle", "_build", "fasta_oracle")


@pytest.fixture(scope="module")
def O():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="module")
def eng():
    from seqkit_b200 import Engine
    e = Engine(max_stream_bytes=48 << 20, max_records=1 << 18, max_samples=0, aux_streams=False, line_ops=True)
    yield e
    e.close()


def run(binary, args, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([binary] + args, cwd=cwd, capture_output=True, env=e, timeout=300)
    return p.returncode, p.stdout, p.stderr


def both(oracle_bin, tmp_path, args, files, env=None, ctx=None):
    res = []
    for tag, binary in (("ours", FASTA), ("oracle", oracle_bin)):
        d = tmp_path / tag
        d.mkdir(exist_ok=True)
        for f in os.listdir(d):
            os.remove(os.path.join(d, f))
        for name, data in files.items():
            (d / name).write_bytes(gzip.compress(data) if name.endswith(".gz") else data)
        r = run(binary, args, str(d), env if tag == "ours" else None)
        outs = {f: gzip.decompress(open(os.path.join(d, f), "rb").read()) for f in sorted(os.listdir(d)) if f.endswith(".fq.gz") and f not in files}
        res.append((r, outs))
    (a, fa), (b, fb) = res
    assert a[0] == b[0], (ctx, a[0], b[0], a[2][-300:], b[2][-300:])
    assert a[1] == b[1], ctx
    if a[0] != 101:
        assert a[2] == b[2], (ctx, a[2][-300:], b[2][-300:])
    assert fa == fb, ctx
    return a


def fasta_of(seed, n):
    rng = random.Random(seed)
    return b"".join(b">s%d desc %d\n%s\n" % (i, rng.randrange(99), G.rand_seq(rng, rng.randrange(0, 120))) for i in range(n))


def test_raw_operator_against_the_oracle(eng, O):
    from seqkit_b200 import _lib as L
    rng = random.Random(5)
    for it in range(12):
        data = G.clean_fastq(it, rng.choice((0, 1, 7, 300, 4000)), read_len=(1, 160), qual_style="mix") if it % 3 else fasta_of(it, rng.choice((1, 50, 900)))
        for first, last in ((0, 0), (3, 0), (0, 5), (10, 10), (200, 0)):
            res, out, _ = eng.line_op(L.LOP_TRIM, data, x=first, y=last)
            want = O.next_op(0, data, x=first, y=last)
            assert res.status == 0 and want[0] == 0 and out == want[1], (it, first, last)
        res, out, _ = eng.line_op(L.LOP_CHECK, data)
        assert res.status == 0 and O.next_op(1, data)[0] == 0
        if data.count(b"\n") % (8 if data[:1] == b"@" else 4) == 0 and data:
            res, o1, o2 = eng.line_op(L.LOP_DEINTERLEAVE, data)
            want = O.next_op(4, data)
            assert res.status == 0 and (o1, o2) == (want[1], want[3]), it
            for fb in (0, 1):
                res, out, _ = eng.line_op(L.LOP_DUAL_UMI, data, x=fb)
                want = O.next_op(5, data, x=fb)
                if want[0] == 0:
                    assert res.status == 0 and out == want[1], (it, fb)
                else:
                    assert res.status in (L.DATA_SEQ_SHORT, L.DATA_QUAL_SHORT) and want[0] == 101
        res, out, _ = eng.line_op(L.LOP_INTERLEAVE, data, data)
        assert res.status == 0 and out == O.next_op(3, data, data)[1], it


def test_statistics_counts(eng):
    sheet, bcs = G.make_sheet(3, 200, 8)
    r1, _ = G.clean_pairs(9, 30000, bcs, p_sub=0.05, p_random=0.3)
    res, got = eng.statistics(r1)
    want = collections.Counter(m.group(0)[4:] for ln in r1.split(b"\n")[0::4] for m in [re.search(rb" BC:[ACGTNacgtn]+", ln)] if m)
    assert res.status == 0 and res.n_records == 30000 and got == dict(want)
    dual, _ = G.clean_pairs(10, 2000, G.make_sheet(4, 12, 16, dual=True)[1])  # 'ACGT+TTGA': the class has no '+'
    res, got = eng.statistics(dual)
    assert all(b"+" not in k and len(k) == 8 for k in got) and sum(got.values()) == 2000


def test_cli_trim_check_statistics(oracle_bin, tmp_path):
    rng = random.Random(11)
    sheet, bcs = G.make_sheet(5, 150, 8)
    for it in range(8):
        n = rng.choice((0, 1, 60, 2500))
        data = G.nasty_fastq(100 + it, n, fatal_ok=it % 2 == 0) if it % 3 == 2 else G.clean_fastq(it, n, read_len=(1, 160), qual_style="mix")
        files = {"in.fq": data}
        for args in (["trim", "in.fq"], ["trim", "--first=3", "in.fq"], ["trim", "--last=7", "--first=2", "in.fq"], ["trim", "--first=1000", "in.fq"]):
            both(oracle_bin, tmp_path, args, files, ctx=(it, args))
        both(oracle_bin, tmp_path, ["check", "in.fq"], files, ctx=(it, "check"))
    both(oracle_bin, tmp_path, ["trim", "--first=2", "in.fa"], {"in.fa": fasta_of(1, 400)}, ctx="fasta")
    both(oracle_bin, tmp_path, ["check", "in.fa"], {"in.fa": fasta_of(2, 400)}, ctx="fasta check")
    both(oracle_bin, tmp_path, ["trim", "--first=x", "in.fq"], {"in.fq": G.clean_fastq(1, 3)}, ctx="bad N")
    both(oracle_bin, tmp_path, ["trim", "--last=3", "missing.fq"], {}, ctx="missing")
    # check: every kind of failure, early, in a middle batch and at the end of the data
    good = G.clean_fastq(7, 9000, read_len=(100, 150))
    lines = good.split(b"\n")
    for where in (2, 4 * 4000 + 2, len(lines) - 3):
        bad = list(lines)
        bad[where] = b"-" + bad[where][1:]
        both(oracle_bin, tmp_path, ["check", "in.fq"], {"in.fq": b"\n".join(bad)}, env={"SK_BATCH_MB": "1"}, ctx=("no plus", where))
    for where in (0, 4 * 5000, 4 * 8999):
        bad = list(lines)
        bad[where] = b"X" + bad[where][1:]
        both(oracle_bin, tmp_path, ["check", "in.fq"], {"in.fq": b"\n".join(bad)}, env={"SK_BATCH_MB": "1"}, ctx=("bad header", where))
        both(oracle_bin, tmp_path, ["trim", "--first=1", "in.fq"], {"in.fq": b"\n".join(bad)}, env={"SK_BATCH_MB": "1"}, ctx=("trim bad header", where))
    both(oracle_bin, tmp_path, ["check", "in.fq"], {"in.fq": good[:-200]}, ctx="truncated")
    both(oracle_bin, tmp_path, ["trim", "--first=1", "in.fq.gz"], {"in.fq.gz": good}, env={"SK_BATCH_MB": "1"}, ctx="gz multi-batch")
    # a quality line shorter than the kept sequence: the reference panics after the record's first print
    both(oracle_bin, tmp_path, ["trim", "--first=1", "in.fq"], {"in.fq": G.clean_fastq(1, 5) + b"@q\nACGTACGT\n+\nIII\n" + G.clean_fastq(2, 5)}, ctx="qual short")
    # statistics: >= 100 distinct barcodes, ties and all; fewer than 100 -> the reference's panic; a bad header
    r1, _ = G.clean_pairs(31, 20000, bcs, p_sub=0.05, p_random=0.2)
    both(oracle_bin, tmp_path, ["statistics", "r1.fq"], {"r1.fq": r1}, env={"SK_BATCH_MB": "1"}, ctx="statistics")
    both(oracle_bin, tmp_path, ["statistics", "r1.fq"], {"r1.fq": G.clean_pairs(32, 500, bcs[:20], p_sub=0, p_n=0, p_random=0)[0]}, ctx="short table")
    both(oracle_bin, tmp_path, ["statistics", "r1.fq"], {"r1.fq": r1 + b"oops\nAC\n+\nII\n"}, env={"SK_BATCH_MB": "1"}, ctx="statistics bad header")


def test_cli_interleave_deinterleave_dual_umi(oracle_bin, tmp_path):
    sheet, bcs = G.make_sheet(6, 24, 8)
    r1, r2 = G.clean_pairs(41, 9000, bcs, read_len=(30, 150))
    files = {"r1.fq": r1, "r2.fq": r2}
    env = {"SK_BATCH_MB": "1"}
    a = both(oracle_bin, tmp_path, ["interleave", "r1.fq", "r2.fq"], files, env=env, ctx="interleave")
    inter = a[1]
    assert inter.count(b"\n") == 8 * 9000
    both(oracle_bin, tmp_path, ["deinterleave", "il.fq", "out"], {"il.fq": inter}, env=env, ctx="deinterleave")
    both(oracle_bin, tmp_path, ["deinterleave", "il.fq", "out"], {"il.fq": inter}, env={"SK_GZIP": "child"}, ctx="deinterleave, gzip children")
    for fb in ("0", "8", "30"):
        both(oracle_bin, tmp_path, ["extract", "dual", "umi", "--first-bases=" + fb, "il.fq"], {"il.fq": inter}, env=env, ctx=("dual umi", fb))
    both(oracle_bin, tmp_path, ["extract", "dual", "umi", "--first-bases=400", "il.fq"], {"il.fq": inter}, ctx="dual umi, N beyond the read")
    # FASTA pairs
    fa = fasta_of(3, 600)
    a = both(oracle_bin, tmp_path, ["interleave", "a.fa", "b.fa"], {"a.fa": fa, "b.fa": fa}, ctx="interleave fasta")
    both(oracle_bin, tmp_path, ["deinterleave", "il.fa", "p"], {"il.fa": a[1]}, ctx="deinterleave fasta")
    both(oracle_bin, tmp_path, ["extract", "dual", "umi", "il.fa"], {"il.fa": a[1]}, ctx="dual umi fasta")
    # failures: the second file shorter, of the other format, a bad line in the first; an odd number of records
    both(oracle_bin, tmp_path, ["interleave", "r1.fq", "r2.fq"], {"r1.fq": r1, "r2.fq": r2[:len(r2) // 2]}, env=env, ctx="second file short")
    both(oracle_bin, tmp_path, ["interleave", "r1.fq", "b.fa"], {"r1.fq": r1, "b.fa": fa}, ctx="formats differ")
    both(oracle_bin, tmp_path, ["interleave", "r1.fq", "r2.fq"], {"r1.fq": r1[:5000] + b"junk\n" + r1[5000:], "r2.fq": r2}, ctx="bad line")
    odd = b"\n".join(inter.split(b"\n")[:4 * 4001]) + b"\n"
    both(oracle_bin, tmp_path, ["deinterleave", "il.fq", "out"], {"il.fq": odd}, env=env, ctx="odd records")
    both(oracle_bin, tmp_path, ["extract", "dual", "umi", "--first-bases=4", "il.fq"], {"il.fq": odd}, env=env, ctx="odd records, umi")
    both(oracle_bin, tmp_path, ["extract", "dual", "umi", "il.fq"], {"il.fq": b"Xbad\nAC\n+\nII\n"}, ctx="bad header, umi")
    for args in (["interleave", "r1.fq"], ["deinterleave", "x"], ["check"], ["statistics"], ["extract", "dual", "umi"]):
        r = run(FASTA, args, str(tmp_path))
        assert r[0] == 255 and r[2].startswith(b"ERROR: Invalid arguments.\n"), args


def test_golden_next_rows_through_the_binary(tmp_path):
    """The committed fixtures of the SURVEY 8(f) operators (tests/golden/golden_next.json, from the Python restatement):
    the `fasta` binary reproduces exit status, stdout, stderr and -- for deinterleave -- the two output files."""
    import golden_util as GU
    argv = {0: lambda c: ["trim", "--first=%d" % c["x"], "--last=%d" % c["y"], "a.fq"], 1: lambda c: ["check", "a.fq"],
            2: lambda c: ["statistics", "a.fq"], 3: lambda c: ["interleave", "a.fq", "b.fq"],
            4: lambda c: ["deinterleave", "a.fq", "out"], 5: lambda c: ["extract", "dual", "umi", "--first-bases=%d" % c["x"], "a.fq"]}
    d = tmp_path / "g"
    d.mkdir()
    n = 0
    for c in GU.next_cases()[::3]:  # (a process start per case: every third keeps the suite short, every operator is in)
        for f in os.listdir(d):
            os.remove(os.path.join(d, f))
        (d / "a.fq").write_bytes(GU.next_blob(c["a"]))
        if c["b"] is not None:
            (d / "b.fq").write_bytes(GU.next_blob(c["b"]))
        code, out, err = run(FASTA, argv[c["op"]](c), str(d))
        ctx = (c["op"], c["tag"])
        assert code == c["exit_code"], (ctx, code, err[-300:])
        if c["op"] == 4:  # the restatement returns the two files in the stdout / second-output slots
            got1 = gzip.decompress((d / "out_1.fq.gz").read_bytes())
            got2 = gzip.decompress((d / "out_2.fq.gz").read_bytes())
            assert out == b"" and got1 == GU.next_blob(c["stdout"]) and got2 == GU.next_blob(c["out2"]), ctx
        else:
            assert out == GU.next_blob(c["stdout"]), ctx
        if code != 101:
            assert err == GU.next_blob(c["stderr"]), (ctx, err[-300:])
        n += 1
    assert n >= 33
