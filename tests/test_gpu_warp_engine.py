"""Parity tests for the warp engine (seqkit_b200/csrc/sk_warp.cu, the default path of header-route
demultiplex): every branch the engine takes for unusual records -- header pieces after the cut, no room
for the UMI tag, '+' lines too short for the in-place patch, tiles of several rounds, tiles it gives up --
must produce the oracle's bytes (fasta_demultiplex.rs:117-249, fasta_trim_by_quality.rs:28-48)."""
import os
import random

import pytest

import fuzzgen as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="module")
def eng():
    from seqkit_b200 import Engine
    e = Engine(max_stream_bytes=48 << 20, max_records=1 << 18, max_samples=512)
    yield e
    e.close()


def _cmp(a, b, ctx=None):
    assert a["exit_code"] == b["exit_code"], ctx
    if a["exit_code"] != 101:
        assert a["stderr"] == b["stderr"], ctx
    assert a["files"] == b["files"], ctx
    if a["exit_code"] == 0:
        assert a["counts"] == b["counts"] and a["total"] == b["total"] and a["identified"] == b["identified"], ctx


def _both(eng, O, sheet, r1, r2, ctx, want_engine=1):
    """plain and fused-trim demultiplex of (r1, r2) against the oracle; want_engine = sk_result.reserved"""
    _cmp(eng.demultiplex(sheet, r1, r2), O.demultiplex(sheet, r1, r2), (ctx, "plain"))
    if want_engine is not None:
        assert eng.last_result.reserved & 3 == want_engine, (ctx, eng.last_result.reserved)
    t1 = O.trim_by_quality(r1, 20)
    t2 = O.trim_by_quality(r2, 20) if r2 is not None else None
    if t1[0] == 0 and (t2 is None or t2[0] == 0):
        _cmp(eng.demultiplex(sheet, r1, r2, fused_trim=20), O.demultiplex(sheet, t1[1], None if t2 is None else t2[1]),
             (ctx, "fused"))
        if want_engine is not None:
            assert eng.last_result.reserved & 3 == want_engine, (ctx, eng.last_result.reserved)


def _reads(seed, n, bcs, read_len=(100, 160), header_tail=(b"",), plus=(b"+",), qual_style="decay", umi_fill=b"ACGT"):
    """n records whose headers carry ' BC:<sheet barcode with U filled>' followed by one of header_tail"""
    rng = random.Random(seed)
    r1, r2 = [], []
    for i in range(n):
        L = rng.randrange(read_len[0], read_len[1] + 1)
        bc = bytes(rng.choice(umi_fill) if ch in b"UN" else ch for ch in rng.choice(bcs))
        if rng.random() < 0.05:
            bc = bytes(rng.choice(b"ACGT") for _ in bc)
        tail = rng.choice(header_tail)
        for mate, dst in ((1, r1), (2, r2)):
            hdr = b"@SIM:%d:%d %d:N:0 BC:%s%s" % (seed, i, mate, bc, tail)
            dst.append(hdr + b"\n" + G.rand_seq(rng, L) + b"\n" + rng.choice(plus) + b"\n" + G.rand_qual(rng, L, qual_style) + b"\n")
    return b"".join(r1), b"".join(r2)


def test_regular_tiles_stay_on_the_warp_engine(eng, O):
    sheet, bcs = G.make_sheet(21, 48, 8, umi=6)
    r1, r2 = _reads(1, 9000, bcs)
    _both(eng, O, sheet, r1, r2, "regular")
    _both(eng, O, sheet, r1, None, "single-end")


def test_header_piece_after_the_barcode(eng, O):
    """' BC:x' in the middle of the header: the kept header is two pieces (fasta_demultiplex.rs:145), which
    the in-place patch cannot express -- the record takes the byte-wise header path."""
    sheet, bcs = G.make_sheet(22, 24, 8, umi=4)
    r1, r2 = _reads(2, 6000, bcs, header_tail=(b"", b" extra", b" x y z", b"\t", b"  "))
    _both(eng, O, sheet, r1, r2, "mid-header")


def test_umi_shapes(eng, O):
    """A sheet of wildcards only (' UMI:' + L characters is one byte longer than the deleted ' BC:' + L; the
    pigeonhole index cannot represent it, so this one runs on the general engine), a UMI longer than two
    registers, and U positions that are not one run."""
    sheet = b"all\tUUUUUUUU\n"
    r1, r2 = _reads(3, 5000, [b"UUUUUUUU"])
    _both(eng, O, sheet, r1, r2, "all-U", want_engine=None)
    sheet2 = b"a\tACUUUUUUUUUU\nb\tTGUUUUUUUUUU\n"  # ten U: the UMI does not fit two registers
    r1, r2 = _reads(4, 5000, [b"ACUUUUUUUUUU", b"TGUUUUUUUUUU"])
    _both(eng, O, sheet2, r1, r2, "ten-U")
    sheet3 = b"a\tAUCUGUTU\nb\tTUGUCUAU\n"  # U positions that are not one run
    r1, r2 = _reads(5, 5000, [b"AUCUGUTU", b"TUGUCUAU"])
    _both(eng, O, sheet3, r1, r2, "scattered-U")


def test_plus_lines_and_tiny_records(eng, O):
    """'+' lines of every shape (bare, with text, empty) and records too short for the patched literals."""
    sheet, bcs = G.make_sheet(23, 16, 6, umi=3)
    r1, r2 = _reads(6, 7000, bcs, read_len=(40, 120), plus=(b"+", b"+SIM", b"", b"+ "), qual_style="mix")
    _both(eng, O, sheet, r1, r2, "plus-shapes")
    # mostly ordinary records with tiny ones in between (a tile of tiny records only is too dense, below)
    rng = random.Random(70)
    a1, a2 = _reads(7, 3000, bcs, read_len=(0, 6), plus=(b"+", b""), qual_style="bad")
    b1, b2 = _reads(71, 3000, bcs, read_len=(120, 150), plus=(b"+", b""), qual_style="mix")
    split = lambda d: [b"\n".join(x) + b"\n" for x in zip(*[iter(d.split(b"\n")[:-1])] * 4)]
    mix = [(x, y) for x, y in zip(split(a1), split(a2))] + [(x, y) for x, y in zip(split(b1), split(b2))] * 3
    rng.shuffle(mix)
    _both(eng, O, sheet, b"".join(x for x, _ in mix), b"".join(y for _, y in mix), "tiny-in-between")


def test_tiles_of_several_rounds_and_dense_tiles(O, monkeypatch):
    """With the tile pinned at 29 lanes: ~150-byte records make three rounds of 32 per tile; ~80-byte records
    are more than 128 per tile, the engine gives the tile up and the operator is re-run on the general
    engine (sk_result.reserved == 2)."""
    from seqkit_b200 import Engine
    monkeypatch.setenv("SK_TILE_LANES", "29")
    sheet, bcs = G.make_sheet(24, 32, 8, umi=0)
    with Engine(max_stream_bytes=16 << 20, max_records=1 << 17, max_samples=64) as e:
        r1, r2 = _reads(8, 20000, bcs, read_len=(40, 60))
        _both(e, O, sheet, r1, r2, "multi-round")
        rng = random.Random(9)
        recs = [b"@r BC:%s\n%s\n+\n%s\n" % (rng.choice(bcs), G.rand_seq(rng, 30), b"I" * 30) for _ in range(30000)]
        data = b"".join(recs)
        _both(e, O, sheet, data, data, "dense", want_engine=2)


def test_tile_size_follows_the_record_size(O):
    """Without SK_TILE_LANES the tile is sized for about 31 records of the data at hand (sk_api.cu:
    choose_tile_lanes): 60 bp reads get a smaller tile than 150 bp reads, both stay on the warp engine in
    one round per tile, and a change of record size between calls is picked up."""
    from seqkit_b200 import Engine
    assert "SK_TILE_LANES" not in os.environ
    sheet, bcs = G.make_sheet(28, 24, 8, umi=4)
    with Engine(max_stream_bytes=16 << 20, max_records=1 << 16, max_samples=64) as e:
        tiles = {}
        for label, rl in (("short", (60, 60)), ("long", (150, 150)), ("short-again", (60, 60))):
            r1, r2 = _reads(13, 8000, bcs, read_len=rl)
            for _ in range(2):  # the second call has the first one's measured record size
                _both(e, O, sheet, r1, r2, label)
            tile = len(r1) / (e.last_result.n_chunks[0] / 4.0)  # bytes per tile (4 slice-table rows each)
            tiles[label] = tile / (len(r1) / 8000.0)              # records per tile
        # a whole number of 400-byte lanes that holds at most 31.1 records, never more than 29 lanes
        assert all(28.0 < v <= 31.2 for v in tiles.values()), tiles


def test_ragged_ends(eng, O):
    sheet, bcs = G.make_sheet(25, 16, 8, umi=4)
    r1, r2 = _reads(10, 3000, bcs)
    for cut in (1, 2, 7, 160, 171, 330):
        _both(eng, O, sheet, r1[:-cut], r2, ("cut", cut), want_engine=None)
        _both(eng, O, sheet, r1[:-cut], None, ("cut-single", cut), want_engine=None)
    _both(eng, O, sheet, r1 + b"\n", r2, "blank-line", want_engine=None)
    _both(eng, O, sheet, r1[: len(r1) // 2], r2, "short-mate-1", want_engine=None)


@pytest.mark.parametrize("lanes", [8, 20, 27, 30])
def test_other_tile_sizes(O, lanes, monkeypatch):
    """SK_TILE_LANES: the tile may be anything from 8 to 30 lanes of 400 bytes (the rest is overhang)."""
    from seqkit_b200 import Engine
    monkeypatch.setenv("SK_TILE_LANES", str(lanes))
    sheet, bcs = G.make_sheet(26, 24, 8, umi=4)
    r1, r2 = _reads(11, 8000, bcs, read_len=(60, 200) if lanes < 30 else (100, 150))
    with Engine(max_stream_bytes=16 << 20, max_records=1 << 16, max_samples=64) as e:
        _both(e, O, sheet, r1, r2, ("lanes", lanes))


def test_general_engine_is_the_second_cuda_opinion(O, monkeypatch):
    """SK_NO_WARP=1 routes demultiplex through the general engine (sk_kernels.cu; sk_result.reserved bit 0 clear),
    the engine every re-run lands on: same bytes.  (Round 1's third engine, the lean one, was retired.)"""
    from seqkit_b200 import Engine
    monkeypatch.setenv("SK_NO_WARP", "1")
    sheet, bcs = G.make_sheet(27, 24, 8, umi=4)
    r1, r2 = _reads(12, 6000, bcs)
    with Engine(max_stream_bytes=16 << 20, max_records=1 << 16, max_samples=64) as e:
        _both(e, O, sheet, r1, r2, "general", want_engine=0)
        for blob in (r1, G.nasty_fastq(3, 800)):
            for op, ref in ((e.trim_by_quality, O.trim_by_quality), (e.mask_by_quality, O.mask_by_quality)):
                got, want = op(blob, 20), ref(blob, 20)
                assert got[0] == want[0] and got[1] == want[1]


def test_shards_concatenate_to_the_single_stream_output(eng, O):
    """SURVEY 8e on one GPU: the job cut into contiguous record ranges (seqkit_b200/shard.py), each range
    demultiplexed on its own, files appended in range order and counters summed = the whole job at once."""
    from seqkit_b200 import shard
    sheet, bcs = G.make_sheet(29, 24, 8, umi=4)
    r1, r2 = _reads(14, 9000, bcs)
    whole = eng.demultiplex(sheet, r1, r2, fused_trim=20)
    parts = [eng.demultiplex(sheet, a, b, fused_trim=20) for a, b in zip(shard.split_records(r1, 3), shard.split_records(r2, 3))]
    assert all(p["exit_code"] == 0 for p in parts) and whole["exit_code"] == 0
    assert shard.merge_files([p["files"] for p in parts]) == whole["files"]
    assert [sum(c) for c in zip(*[p["counts"] for p in parts])] == whole["counts"]
    assert sum(p["total"] for p in parts) == whole["total"] and sum(p["identified"] for p in parts) == whole["identified"]
    want = O.demultiplex(sheet, O.trim_by_quality(r1, 20)[1], O.trim_by_quality(r2, 20)[1])
    assert whole["files"] == want["files"] and whole["counts"] == want["counts"]


@pytest.mark.parametrize("gather", [False, True])
def test_trim_and_mask_on_the_warp_engine(O, monkeypatch, gather):
    """SK_WARP_STREAM=1: trim and mask by quality through the warp engine (second look-back on output bytes,
    in-place mask, two runs per record; by default only mask takes it).  The bytes must be the oracle's:
    regular data, uneven record sizes (tiles of several rounds), every nasty record shape, failing records
    (the host replays the records before them)."""
    from seqkit_b200 import Engine
    monkeypatch.setenv("SK_WARP_STREAM", "1")
    # trim: unordered tiles + scan + gather (the default), or the in-order look-back on output bytes
    monkeypatch.setenv("SK_TRIM_GATHER", "1" if gather else "0")
    with Engine(max_stream_bytes=16 << 20, max_records=1 << 17, max_samples=64) as e:
        blobs = [("decay", G.clean_fastq(41, 9000, read_len=(150, 150))), ("uneven", G.clean_fastq(42, 12000, read_len=(20, 160), qual_style="mix")),
                 ("short", G.clean_fastq(43, 20000, read_len=(30, 40), qual_style="bad"))]
        blobs += [("nasty%d" % k, G.nasty_fastq(50 + k, 1500)) for k in range(6)]
        blobs += [("ragged", G.clean_fastq(44, 500)[:-9]), ("empty", b""), ("one", b"@r\nACGT\n+\nII#I\n"), ("blank", G.clean_fastq(45, 300) + b"\n")]
        for label, blob in blobs:
            for q in (20, 0, 2, 41, 255):
                got, want = e.trim_by_quality(blob, q), O.trim_by_quality(blob, q)
                assert got[0] == want[0] and got[1] == want[1], ("trim", label, q)
                if got[0] != 101:
                    assert got[2] == want[2], ("trim", label, q)
                got, want = e.mask_by_quality(blob, q), O.mask_by_quality(blob, q)
                assert got[0] == want[0] and got[1] == want[1], ("mask", label, q)
                if got[0] != 101:
                    assert got[2] == want[2], ("mask", label, q)
        assert (e.last_result.reserved & 3) in (1, 2)


def test_mask_in_place_and_its_ordered_form(O, monkeypatch):
    """Mask by quality of a regular file (bare '+' lines, final newline) keeps every record's length: the warp
    engine writes each tile at its input offsets without a look-back on output bytes (sk_result.reserved == 1).
    A single record that changes its length -- a named '+' line, a missing final newline, CRLF is fine (same
    length) -- or fails (fasta_mask_by_quality.rs:21-23,35-37) makes sk_wait run the ordered form (bit 2); the
    bytes are the oracle's either way, and with the in-place form switched off."""
    from seqkit_b200 import Engine
    reg = G.clean_fastq(61, 6000, read_len=(150, 150))
    recs = reg.split(b"\n")
    named = list(recs)
    named[4 * 3000 + 2] = b"+named line"
    cases = [("regular", reg, 1), ("named plus", b"\n".join(named), 5), ("no final newline", reg[:-1], 5),
             # failing records: the ordered form reports them, then the host replays the records before them with a
             # record limit (ordered from the start: bits of the replay)
             ("bad header", reg + b"oops\nACGT\n+\nIIII\n" + reg, 1),
             ("length mismatch", reg + b"@x\nACGT\n+\nIII\n", 1), ("empty", b"", None)]
    for switch in ("1", "0"):
        monkeypatch.setenv("SK_MASK_INPLACE", switch)
        with Engine(max_stream_bytes=16 << 20, max_records=1 << 17, max_samples=64) as e:
            for label, blob, bits in cases:
                for q in (20, 41):
                    got, want = e.mask_by_quality(blob, q), O.mask_by_quality(blob, q)
                    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2], (label, q, switch)
                    if bits is not None:
                        assert e.last_result.reserved == (bits if switch == "1" else 1), (label, q, switch, e.last_result.reserved)


def test_add_barcode_on_the_warp_engine(eng, O):
    """fasta add barcode (fasta_add_barcode.rs:19-44) on the warp engine: barcodes of one length put every record at
    its input offset plus a multiple of that growth (first form); any other shape -- barcodes of several lengths, a
    header that ends in white space, a barcode file that runs out or yields nothing -- repeats the pass with an
    ordered look-back on output bytes (sk_result.reserved bit 2).  The barcode records come from the global line
    table, so index reads far denser than a chunk engine's record slots are fine."""
    def run(reads, bc, ctx, want=None):
        got, exp = eng.add_barcode(reads, bc), O.add_barcode(reads, bc)
        assert got[0] == exp[0], ctx
        assert got[1] == exp[1], ctx
        if exp[0] != 101:
            assert got[2] == exp[2], ctx
        if want is not None:
            assert eng.last_result.reserved & 7 == want, (ctx, eng.last_result.reserved)

    n = 60000
    reads = G.clean_fastq(11, n, read_len=(100, 151))
    uni = G.index_reads(12, n, [b"ACGTACGT", b"TTGGCCAA"], p_random=0.2, p_n=0.05)
    run(reads, uni, "one length", want=1)
    dual = G.index_reads(13, n, [b"ACGTACGT+TTGGCCAA"])
    run(reads, dual, "dual index, one length", want=1)
    run(reads, G.index_reads(14, n, [b"ACGT", b"GG+TT", b"ACGTACGTAC"]), "three lengths", want=5)
    run(reads, G.index_reads(15, n - 1000, [b"ACGTACGT"]), "barcode file runs out: the last one is reused", want=1)
    run(reads, G.index_reads(16, n // 2, [b"ACGT", b"ACGTAC"]), "runs out, two lengths", want=5)
    run(reads, b"", "no barcodes at all", want=1)
    run(reads, b"junk\nlines\n", "not a FASTA/FASTQ file: empty barcodes", want=1)
    run(reads, uni[:-9], "last barcode record cut short", want=1)
    fa_bc = b"".join(b">b%d\nACGTAC\n" % i for i in range(n))
    run(reads, fa_bc, "FASTA barcode file", want=1)
    # a header that ends in white space shrinks before the tag goes on (:33)
    ws = reads.replace(b"\n", b" \t\n", 1)
    run(ws, uni, "one header ends in white space", want=5)
    # tiny barcode records: ~2000 per 32 KiB
    tiny = b"".join(b"@\nAC\n+\nII\n" for _ in range(n))
    run(reads, tiny, "dense barcode file", want=1)
    # records shorter than their new header line; empty lines
    short = (b"@" + b"h" * 89 + b"\nA\n+\nI\n@" + b"g" * 89 + b"\n\n+\n\n") * 1000 + G.clean_fastq(25, 300)
    run(short, G.index_reads(17, 2300, [b"ACGTACGTACGTACGTACGT"]), "short records", want=1)
    # failures: a line that is neither '@' nor '>' (message after the BC'd header), FASTA record among FASTQ
    bad = G.clean_fastq(18, 3000) + b"oops\nAC\n+\nII\n" + G.clean_fastq(19, 100)
    run(bad, G.index_reads(20, 4000, [b"ACGTACGT"]), "bad line")
    # several rounds per tile (short records) and a record across tiles
    run(G.clean_fastq(21, 20000, read_len=(5, 30)), G.index_reads(22, 20000, [b"ACGTACGT"]), "short reads", want=1)
    run(G.clean_fastq(23, 5000, read_len=(400, 700)), G.index_reads(24, 5000, [b"ACGTACGT"]), "long reads", want=1)
