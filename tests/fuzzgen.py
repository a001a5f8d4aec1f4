"""Seeded FASTQ generators for the differential tests (oracle vs restatement vs CUDA path).

`clean_*` produce well-formed Illumina-style input; `nasty_*` inject the edge cases listed in
SURVEY.md section 8c (missing final newline, CRLF, empty seq/qual, quality bytes < 33, '@'/'+'-leading
quality lines, N / lower-case / '+' in barcodes, truncated last record, trailing blank line, ...).
"""
from __future__ import annotations

import random

BASES = b"ACGT"


def rand_seq(rng: random.Random, n: int, p_n: float = 0.01) -> bytes:
    return bytes(78 if rng.random() < p_n else BASES[rng.randrange(4)] for _ in range(n))


def rand_qual(rng: random.Random, n: int, style: str = "decay") -> bytes:
    if style == "uniform":
        return bytes(33 + rng.randrange(0, 42) for _ in range(n))
    if style == "bins":
        return bytes(33 + rng.choice((2, 11, 25, 37)) for _ in range(n))
    if style == "good":
        return bytes(33 + rng.randrange(30, 41) for _ in range(n))
    if style == "bad":
        return bytes(33 + rng.randrange(0, 12) for _ in range(n))
    # 3'-decaying profile with optional crash position
    crash = rng.randrange(n) if n and rng.random() < 0.15 else n
    scale = rng.choice((0, 2, 6, 12))
    out = bytearray()
    for i in range(n):
        if i >= crash:
            q = 2
        else:
            x = i / max(n - 1, 1)
            q = int(round(38 - 30 * x ** 3 + rng.uniform(-scale, scale)))
            q = min(41, max(2, q))
        out.append(33 + q)
    return bytes(out)


def header(rng: random.Random, i: int, mate: int = 1, bc: bytes | None = None) -> bytes:
    h = b"@SIM:1:FC:%d:%d:%d:%d %d:N:0:1" % (rng.randrange(1, 9), rng.randrange(1101, 2679), rng.randrange(1000, 30000),
                                           rng.randrange(1000, 30000), mate)
    if rng.random() < 0.3:
        h = b"@r%d" % i
    if bc is not None:
        h += b" BC:" + bc
    return h


def clean_fastq(seed: int, n: int, read_len=(20, 160), qual_style: str = "decay") -> bytes:
    rng = random.Random(seed)
    parts = []
    for i in range(n):
        L = rng.randrange(read_len[0], read_len[1] + 1)
        style = qual_style if qual_style != "mix" else rng.choice(("decay", "uniform", "bins", "good", "bad"))
        parts.append(header(rng, i) + b"\n" + rand_seq(rng, L) + b"\n+\n" + rand_qual(rng, L, style) + b"\n")
    return b"".join(parts)


def nasty_fastq(seed: int, n: int, fatal_ok: bool = True) -> bytes:
    """Mostly valid records with per-record edge-case mutations.  With fatal_ok the stream may
    contain records that make the reference stop (non-'@' header, length mismatch, ...)."""
    rng = random.Random(seed)
    parts = []
    for i in range(n):
        L = rng.choice((0, 1, 2, 5, 31, 32, 33, 50, 150, 151, 300)) if rng.random() < 0.5 else rng.randrange(0, 200)
        hdr = header(rng, i)
        seq = rand_seq(rng, L, 0.05)
        qual = rand_qual(rng, L, rng.choice(("decay", "uniform", "bins", "good", "bad")))
        plus = b"+"
        eol = b"\n"
        r = rng.random()
        if r < 0.05:
            eol = b"\r\n"
        elif r < 0.10 and L:
            qual = bytes(rng.randrange(0, 33) if rng.random() < 0.3 else c for c in qual)  # bytes < 33 wrap
            qual = qual.replace(b"\n", b"!")
        elif r < 0.15 and L:
            qual = rng.choice((b"@", b"+")) + qual[1:]
        elif r < 0.20:
            plus = b"+" + hdr[1:]
        elif r < 0.24:
            hdr += rng.choice((b" ", b"\t", b"  \t ", b" \x0b\x0c"))
        elif r < 0.27 and L:
            qual = qual[:-1] + rng.choice((b" ", b"\t"))  # trailing whitespace inside the quality line
        elif r < 0.30 and L > 2:
            qual = qual + b"  "  # longer qual than seq (whitespace tail)
        elif r < 0.32 and fatal_ok and L > 3:
            seq = seq[:-2]  # length mismatch: mask fatal, trim may panic
        elif r < 0.33 and fatal_ok:
            hdr = b"X" + hdr[1:]
        elif r < 0.34 and fatal_ok:
            hdr = b""  # blank header line
        elif r < 0.36 and L:
            qual = bytes(rng.randrange(33, 127) for _ in range(L))
        parts.append(hdr + eol + seq + eol + plus + eol + qual + eol)
    data = b"".join(parts)
    r = rng.random()
    if r < 0.15 and data.endswith(b"\n"):
        data = data[:-1]  # missing final newline
    elif r < 0.25:
        cut = rng.randrange(0, min(len(data), 400) + 1)
        data = data[:len(data) - cut]  # truncated last record(s)
    elif r < 0.30 and fatal_ok:
        data += b"\n"  # trailing blank line -> fatal after all records
    return data


def make_sheet(seed: int, n_samples: int, bc_len: int, umi: int = 0, dual: bool = False, min_dist: int = 3,
               wild_n: float = 0.0):
    """Returns (sheet_bytes, [barcode bytes]) -- barcodes pairwise Hamming >= min_dist on the literal part."""
    rng = random.Random(seed)
    codes: list[bytes] = []
    tries = 0
    while len(codes) < n_samples:
        c = bytes(BASES[rng.randrange(4)] for _ in range(bc_len))
        tries += 1
        if tries > 200000:
            min_dist = max(0, min_dist - 1)
            tries = 0
        if all(sum(a != b for a, b in zip(c, d)) >= min_dist for d in codes):
            codes.append(c)
    bcs = []
    for c in codes:
        if dual:
            h = bc_len // 2
            c = c[:h] + b"+" + c[h:]
        if wild_n:
            c = bytes(78 if (rng.random() < wild_n and ch != 43) else ch for ch in c)
        bcs.append(c + b"U" * umi)
    lines = [b"# sample\tbarcode\n"]
    for i, b in enumerate(bcs):
        lines.append(b"S%03d\t%s\n" % (i, b))
    return b"".join(lines), bcs


def observed_barcode(rng: random.Random, bcs: list[bytes], p_sub=0.02, p_n=0.01, p_random=0.05, p_lower=0.0) -> bytes:
    b = bytearray(rng.choice(bcs))
    rnd = rng.random() < p_random
    for k in range(len(b)):
        if b[k] == 43:  # '+'
            continue
        if b[k] in (85, 78) or rnd:  # U / N wildcard positions or fully random barcode
            b[k] = BASES[rng.randrange(4)]
        if rng.random() < p_sub:
            b[k] = BASES[rng.randrange(4)]
        if rng.random() < p_n:
            b[k] = 78
        if p_lower and rng.random() < p_lower:
            b[k] = bytes([b[k]]).lower()[0]
    return bytes(b)


def clean_pairs(seed: int, n: int, bcs: list[bytes], read_len=(30, 151), bc_in_r2: bool = True, qual_style="decay",
                **obs_kw):
    """Returns (r1, r2) with ' BC:<observed>' in the headers (as `fasta add barcode` produces)."""
    rng = random.Random(seed)
    p1, p2 = [], []
    for i in range(n):
        bc = observed_barcode(rng, bcs, **obs_kw)
        base = header(rng, i)
        L1 = rng.randrange(read_len[0], read_len[1] + 1)
        L2 = rng.randrange(read_len[0], read_len[1] + 1)
        h1 = base + b" BC:" + bc
        h2 = base.replace(b" 1:N", b" 2:N") + (b" BC:" + bc if bc_in_r2 else b"")
        p1.append(h1 + b"\n" + rand_seq(rng, L1) + b"\n+\n" + rand_qual(rng, L1, qual_style) + b"\n")
        p2.append(h2 + b"\n" + rand_seq(rng, L2) + b"\n+\n" + rand_qual(rng, L2, qual_style) + b"\n")
    return b"".join(p1), b"".join(p2)


def index_reads(seed: int, n: int, bcs_part: list[bytes], **obs_kw) -> bytes:
    rng = random.Random(seed)
    parts = []
    for i in range(n):
        bc = observed_barcode(rng, bcs_part, **obs_kw)
        parts.append(b"@i%d\n" % i + bc + b"\n+\n" + b"I" * len(bc) + b"\n")
    return b"".join(parts)


def nasty_headers_pairs(seed: int, n: int, bcs: list[bytes]):
    """Pairs whose headers stress the BC regex / trim_end / drain logic."""
    rng = random.Random(seed)
    p1, p2 = [], []
    for i in range(n):
        bc = observed_barcode(rng, bcs, p_sub=0.05, p_n=0.03, p_random=0.1, p_lower=0.02)
        base = b"@q%d" % i
        r = rng.random()
        if r < 0.15:
            h1 = base + b" BC:" + bc + b" extra:1"  # BC mid-header
        elif r < 0.25:
            h1 = base + b"  BC:" + bc + b"  \t"  # whitespace around
        elif r < 0.32:
            h1 = base + b" BC:?? BC:" + bc  # first ' BC:' not followed by a class char
        elif r < 0.38:
            h1 = base + b" BC:" + bc + b" BC:" + bc  # two fields, only the first is removed
        elif r < 0.43:
            h1 = base + b" \t BC:" + bc  # prefix ends in whitespace -> trim_end eats into the prefix
        elif r < 0.46:
            h1 = base + b" BC:" + bc + b"x"  # class run stops at 'x'
        else:
            h1 = base + b" 1:N:0 BC:" + bc
        r = rng.random()
        if r < 0.4:
            h2 = base + b" 2:N:0 BC:" + bc
        elif r < 0.6:
            h2 = base + b" 2:N:0"
        elif r < 0.7:
            h2 = base + b" BC:" + bc[:3] + b" tail "
        else:
            h2 = base + b" BC:" + bc + b"\t"
        L1, L2 = rng.randrange(1, 60), rng.randrange(1, 60)
        p1.append(h1 + b"\n" + rand_seq(rng, L1) + b"\n+\n" + rand_qual(rng, L1, "uniform") + b"\n")
        p2.append(h2 + b"\n" + rand_seq(rng, L2) + b"\n+\n" + rand_qual(rng, L2, "uniform") + b"\n")
    return b"".join(p1), b"".join(p2)
