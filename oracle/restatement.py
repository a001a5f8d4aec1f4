"""Second, independently written restatement of the reference's per-read FASTQ batch path.

TEST INFRASTRUCTURE ONLY (same rule as fasta_oracle.c): imported by tests/ and by
tests/golden/make_golden.py, never by the product path.

PARITY UNPINNED: the reference (Rust) has no tests or fixtures on this path and cannot be
built in this image.  This module exists so that the C oracle is not the *only* reading of
the Rust source: it is written against the same cited lines but with Python's own string
machinery (`str`, `re`) standing in for Rust's `String`, `regex` and `str::trim_end`, and
the two are fuzzed against each other in tests/test_oracle.py.

Each function mirrors one `main()` of the reference and returns (exit_code, stdout, stderr[, files]).
Paths are relative to /root/reference/src/.
"""
from __future__ import annotations

import re

# char::is_whitespace == Unicode White_Space (NOT Python's str.isspace, which adds U+001C..1F)
_WS = "\t\n\x0b\x0c\r \x85\xa0\u1680" + "".join(chr(c) for c in range(0x2000, 0x200B)) + "\u2028\u2029\u202f\u205f\u3000"
_BC_RE = re.compile(r" BC:[ACGTNacgtn+]+")  # fasta_demultiplex.rs:38


class _Exit(Exception):
    def __init__(self, code):
        self.code = code


class _Proc:
    def __init__(self):
        self.out = []
        self.err = []

    def error(self, msg):  # common.rs:11-16
        self.err.append("ERROR: " + msg + "\n")
        raise _Exit(255)

    def panic(self, msg):
        self.err.append("thread 'main' panicked: " + msg + "\n")
        raise _Exit(101)


class _Reader:
    """FileReader over bytes, common.rs:83-112."""

    def __init__(self, proc, data: bytes):
        self.proc, self.data, self.pos = proc, data, 0

    def read_line(self):
        """Returns (ok, line). line is '' at EOF (line.clear())."""
        if self.pos >= len(self.data):
            return False, ""
        nl = self.data.find(b"\n", self.pos)
        end = len(self.data) if nl < 0 else nl + 1
        raw = self.data[self.pos:end]
        try:
            s = raw.decode("utf-8")  # strict: same acceptance as str::from_utf8
        except UnicodeDecodeError:
            self.proc.error("I/O error while reading from file.")
        self.pos = end
        return True, s


def _trim_end(s: str) -> str:
    return s.rstrip(_WS)


def _trim(s: str) -> str:
    return s.strip(_WS)


def _enc(parts) -> bytes:
    return "".join(parts).encode("utf-8")


def _byte_slice(proc, s: str, k: int, what: str) -> str:
    """&s[..k] with k a BYTE index: panics when out of range or off a char boundary."""
    b = s.encode("utf-8")
    if k > len(b):
        proc.panic("byte index out of range " + what)
    if k < len(b) and (b[k] & 0xC0) == 0x80:
        proc.panic("byte index is not a char boundary " + what)
    return b[:k].decode("utf-8")


def trim_by_quality(data: bytes, min_baseq: int):
    """fasta_trim_by_quality.rs:10-50"""
    P = _Proc()
    f = _Reader(P, data)
    code = 0
    try:
        while True:
            ok, line = f.read_line()  # :19
            if not ok:
                break
            if not line.startswith("@"):  # :20-22
                P.error("Invalid FASTQ format encountered.")
            P.out.append(line)  # :23
            _, seq = f.read_line()
            _, _plus = f.read_line()
            _, qual = f.read_line()
            qb = qual.encode("utf-8")
            total = -50  # :28
            lowest_total = total
            k = len(_trim_end(qual).encode("utf-8"))  # :31 (byte length)
            lowest_k = k
            while k > 0:  # :33-42
                k -= 1
                total += ((qb[k] - 33) & 0xFF) - min_baseq  # :35, wrapping u8 subtraction
                if total > 0:
                    break
                if total < lowest_total:
                    lowest_total = total
                    lowest_k = k
            if lowest_k == 0:  # :44-45
                P.out.append("N\n+\n!\n")
            else:  # :47
                a = _byte_slice(P, seq, lowest_k, "seq")
                b = _byte_slice(P, qual, lowest_k, "qual")
                P.out.append(a + "\n+\n" + b + "\n")
    except _Exit as e:
        code = e.code
    return code, _enc(P.out), _enc(P.err)


def mask_by_quality(data: bytes, min_baseq: int):
    """fasta_mask_by_quality.rs:11-47"""
    P = _Proc()
    f = _Reader(P, data)
    code = 0
    try:
        while True:
            ok, line = f.read_line()  # :20
            if not ok:
                break
            if not line.startswith("@"):
                P.error("Invalid FASTQ format encountered.")
            output = [line]  # :25-26
            _, seq = f.read_line()
            _, _plus = f.read_line()
            _, bq = f.read_line()
            if seq.endswith("\n"):  # :32
                seq = seq[:-1]
            if bq.endswith("\n"):  # :33
                bq = bq[:-1]
            if len(seq.encode("utf-8")) != len(bq.encode("utf-8")):  # :35-37 (String::len is bytes)
                P.error("Read sequence and base qualities are of different length.")
            for base, q in zip(seq, bq):  # :40-43 (chars)
                output.append("N" if ((ord(q) & 0xFF) - 33) & 0xFF < min_baseq else base)
            output.append("\n+\n" + bq + "\n")  # :44
            P.out.append("".join(output))  # :45
    except _Exit as e:
        code = e.code
    return code, _enc(P.out), _enc(P.err)


def add_barcode(fastq: bytes, barcodes: bytes):
    """fasta_add_barcode.rs:11-45"""
    P = _Proc()
    fq = _Reader(P, fastq)
    bf = _Reader(P, barcodes)
    code = 0
    barcode = ""
    try:
        while True:
            _, header = bf.read_line()  # :20
            if header.startswith("@"):  # :21-24
                _, barcode = bf.read_line()
                bf.read_line()
                bf.read_line()
            elif header.startswith(">"):  # :25-27
                _, barcode = bf.read_line()
            ok, header = fq.read_line()  # :29
            if not ok:
                break
            P.out.append(_trim_end(header) + " BC:" + _trim_end(barcode) + "\n")  # :33
            if header.startswith("@"):
                for _ in range(3):
                    P.out.append(fq.read_line()[1])
            elif header.startswith(">"):
                P.out.append(fq.read_line()[1])
            else:
                P.error("Invalid FASTQ line:\n" + header)
    except _Exit as e:
        code = e.code
    return code, _enc(P.out), _enc(P.err)


def _barcode_diff(observed: bytes, candidate: bytes) -> int:
    """fasta_demultiplex.rs:269-277"""
    assert len(observed) == len(candidate)
    mm = 0
    for o, c in zip(observed, candidate):
        if c == 0x4E or c == 0x55:  # 'N' / 'U'
            continue
        if o != c:
            mm += 1
    return mm


def demultiplex(sheet: bytes, fastq_1: bytes, fastq_2: bytes | None = None, index1: bytes | None = None,
                index2: bytes | None = None, dry_run: int = 0):
    """fasta_demultiplex.rs:30-265.

    Returns dict(exit_code, stdout, stderr, files={filename: decompressed bytes}, counts=[...],
    total, identified)."""
    P = _Proc()
    files: dict[str, list[str]] = {}
    samples = []  # dicts: name, barcode, output(list of file keys), total_reads
    total_reads = identified_reads = 0
    code = 0
    try:
        fastq = [_Reader(P, fastq_1)]
        if fastq_2 is not None:
            fastq.append(_Reader(P, fastq_2))
        paired_end = len(fastq) == 2
        index_fastq = [_Reader(P, x) for x in (index1, index2) if x is not None]
        P.err.append("Reading sample sheet...\n")  # :58
        sh = _Reader(P, sheet)
        barcode_len = 0
        while True:  # :63-95
            ok, line = sh.read_line()
            if not ok:
                break
            if line.startswith("#"):
                continue
            cols = _trim(line).split("\t")
            if len(cols) < 2:
                continue
            name = cols[0]
            if cols[1] == "":
                P.error("Sample %s has no barcode." % name)
            blen = len(cols[1].encode("utf-8"))
            if barcode_len == 0:
                barcode_len = blen
            elif blen != barcode_len:
                P.error("Barcodes in sample sheet must all be of same length.")
            outputs = []
            if dry_run > 0:
                pass
            elif paired_end:
                outputs = [name + "_1.fq.gz", name + "_2.fq.gz"]
            else:
                outputs = [name + ".fq.gz"]
            for o in outputs:
                files[o] = []  # File::create truncates: later duplicate names share the path
            samples.append({"name": name, "barcode": cols[1], "output": outputs, "total_reads": 0})
        for s in range(len(samples)):  # :98-104
            for k in range(s + 1, len(samples)):
                if samples[s]["name"] == samples[k]["name"]:
                    P.error("Sample %s is listed multiple times in sample sheet." % samples[s]["name"])
        P.err.append("Starting demultiplexing in %s end mode...\n" % ("paired" if paired_end else "single"))
        extra: dict[str, int] = {}
        while True:  # :117-249
            ok, header = fastq[0].read_line()
            if not ok:
                break
            if not header.startswith("@"):
                P.error("Invalid FASTQ header line:\n" + header)
            barcode = ""
            if index_fastq:  # :126-136
                for ifq in index_fastq:
                    if barcode != "":
                        barcode += "+"
                    _, line = ifq.read_line()
                    if not line.startswith("@"):
                        P.panic("assertion failed: line.starts_with('@')")
                    _, line = ifq.read_line()
                    barcode += _trim_end(line)
                    _, line = ifq.read_line()
                    if not line.startswith("+"):
                        P.panic("assertion failed: line.starts_with('+')")
                    ifq.read_line()
            else:  # :138-146
                hit = _BC_RE.search(header)
                if hit is None:
                    P.error("No BC:xxxx field found.")
                barcode += header[hit.start() + 4:hit.end()]
                header = header[:hit.start()] + header[hit.end():]
            bcb = barcode.encode("utf-8")
            if len(bcb) != barcode_len:  # :148-150
                P.error("Sequenced barcode %s is of different length (%d nt) than barcodes in the sample sheet (%d nt)."
                        % (barcode, len(bcb), barcode_len))
            best = eq = 0
            lowest = None  # usize::MAX
            for s, smp in enumerate(samples):  # :157-166
                d = _barcode_diff(bcb, smp["barcode"].encode("utf-8"))
                if lowest is None or d < lowest:
                    lowest, best, eq = d, s, s
                elif d == lowest:
                    eq = s
            total_reads += 1
            write_read_out = False
            if lowest is not None and lowest <= 1:  # :172
                if best == eq:
                    identified_reads += 1
                    samples[best]["total_reads"] += 1
                    write_read_out = not (dry_run > 0)
                else:
                    P.err.append(
                        "WARNING: Sequenced barcode %s was an equally good match (%d mismatches) for samples %s (%s) "
                        "and %s (%s), and was therefore not assigned to any sample.\n"
                        % (barcode, lowest, samples[best]["name"], samples[best]["barcode"], samples[eq]["name"],
                           samples[eq]["barcode"]))
            elif dry_run > 0:
                extra[barcode] = extra.get(barcode, 0) + 1
            if write_read_out:
                smp = samples[best]
                umi = "".join(o for c, o in zip(smp["barcode"], barcode) if c == "U")  # :200-203
                o0 = files[smp["output"][0]]
                o0.append(_trim_end(header) + ((" UMI:" + umi) if umi else "") + "\n")  # :206-208
                for _ in range(3):
                    o0.append(fastq[0].read_line()[1])
                if paired_end:
                    _, line = fastq[1].read_line()
                    if not index_fastq:  # :219-227
                        hit = _BC_RE.search(line)
                        if hit is not None and hit.end() > 0:
                            line = line[:hit.start()] + line[hit.end():]
                    o1 = files[smp["output"][1]]
                    o1.append(_trim_end(line) + ((" UMI:" + umi) if umi else "") + "\n")
                    for _ in range(3):
                        o1.append(fastq[1].read_line()[1])
            else:
                for _ in range(3):
                    fastq[0].read_line()
                if paired_end:
                    for _ in range(4):
                        fastq[1].read_line()
            if dry_run > 0 and total_reads >= dry_run:
                break
        if dry_run > 0:  # :251-261
            P.err.append("Dry run completed with %d clusters. Barcodes found:\n" % total_reads)
            entries = [(s["name"], s["total_reads"]) for s in samples] + list(extra.items())
            entries.sort(key=lambda x: x[1])
            entries.reverse()
            if len(entries) < 100:
                P.panic("range end index 100 out of range for slice")
            for name, count in entries[:100]:
                P.out.append("- %s: %d\n" % (name, count))
        if total_reads == 0:
            pct = "NaN"
        else:
            pct = "%.1f" % (identified_reads / total_reads * 100.0)
        P.err.append("%d / %d (%s%%) clusters carried a barcode matching one of the provided samples.\n"
                     % (identified_reads, total_reads, pct))
    except _Exit as e:
        code = e.code
    return {
        "exit_code": code,
        "stdout": _enc(P.out),
        "stderr": _enc(P.err),
        "files": {k: _enc(v) for k, v in files.items()},
        "counts": [s["total_reads"] for s in samples],
        "names": [s["name"] for s in samples],
        "total": total_reads,
        "identified": identified_reads,
    }


# ---------------------------------------------------------------------------------------------------------
# SURVEY.md section 8(f) operators (the rows around the hot path).  Same return shape as pyoracle.next_op:
# (exit_code, stdout, stderr, second_output).
# ---------------------------------------------------------------------------------------------------------
def _slice(proc, s: str, a: int, b: int | None, what: str) -> str:
    """&s[a..b] (b None: &s[a..]) with BYTE indices: panics when out of range or off a char boundary."""
    raw = s.encode("utf-8")
    if b is None:
        b = len(raw)
        if a > b:
            proc.panic("byte index out of range " + what)
    if a > b or b > len(raw):
        proc.panic("byte index out of range " + what)
    for k in (a, b):
        if k < len(raw) and (raw[k] & 0xC0) == 0x80:
            proc.panic("byte index is not a char boundary " + what)
    return raw[a:b].decode("utf-8")


def trim_fixed(data: bytes, remove_first: int, remove_last: int):
    """fasta_trim.rs:24-47."""
    P = _Proc()
    f = _Reader(P, data)
    try:
        while True:
            ok, line = f.read_line()  # :27
            if not ok:
                break
            if not line.startswith(">") and not line.startswith("@"):  # :28-30
                P.error("Invalid FASTA/FASTQ format encountered.")
            _, seq = f.read_line()  # :32
            seq_len = len(_trim_end(seq).encode("utf-8"))  # :33 (a byte length)
            cut = remove_first + remove_last < seq_len
            if cut:  # :34-38
                P.out.append(line + _slice(P, seq, remove_first, seq_len - remove_last, "of seq (fasta_trim.rs:35)") + "\n")
            else:
                P.out.append(line + "\n")
            if line.startswith("@"):  # :40-47
                f.read_line()
                _, qual = f.read_line()
                if cut:
                    P.out.append("+\n" + _slice(P, qual, remove_first, seq_len - remove_last, "of qual (fasta_trim.rs:44)") + "\n")
                else:
                    P.out.append("+\n\n")
    except _Exit as e:
        return e.code, _enc(P.out), _enc(P.err), b""
    return 0, _enc(P.out), _enc(P.err), b""


def check(data: bytes):
    """fasta_check.rs:14-70."""
    P = _Proc()
    f = _Reader(P, data)
    prev, lines_read = [], 0

    def read():  # ReaderWithMemory::read_line (:30-38)
        nonlocal lines_read
        ok, line = f.read_line()
        if not ok:
            return False, line
        prev.append(line)
        if len(prev) > 10:
            prev.pop(0)
        lines_read += 1
        return True, line

    def history():  # :40-46
        return "".join(l + "\n" for l in prev)

    try:
        while True:
            ok, line = read()
            if not ok:
                break
            if line.startswith(">"):
                read()
            elif line.startswith("@"):
                read()
                _, line = read()
                if not line.startswith("+"):  # :58-61
                    P.error("Missing quality header prefix '+' on line %d:\n%s\n" % (lines_read, history()))
                read()
            else:  # :64-67
                P.error("Missing header prefix '>' or '@' on line %d:\n%s\n" % (lines_read, history()))
    except _Exit as e:
        return e.code, _enc(P.out), _enc(P.err), b""
    return 0, _enc(P.out), _enc(P.err), b""


_STAT_RE = re.compile(r" BC:[ACGTNacgtn]+")  # fasta_statistics.rs:17 (no '+')


def statistics(data: bytes):
    """fasta_statistics.rs:13-52.  Barcodes of equal count: barcode descending (the reference prints its HashMap's
    order there; the C oracle and the product fix the same order)."""
    P = _Proc()
    f = _Reader(P, data)
    total, seen = 0, {}
    try:
        while True:
            ok, line = f.read_line()
            if not ok:
                break
            m = _STAT_RE.search(line)  # :25-28
            if m:
                bc = line[m.start() + 4:m.end()]
                seen[bc] = seen.get(bc, 0) + 1
            if line.startswith("@"):  # :31-37
                for _ in range(3):
                    f.read_line()
            elif line.startswith(">"):
                f.read_line()
            else:
                P.error("Invalid FASTQ header:\n" + line)
            total += 1
        P.out.append("Total sequence records: %d\n" % total)  # :42
        P.out.append("Most frequent sample barcodes:\n")      # :44
        entries = sorted(seen.items(), key=lambda kv: (kv[1], kv[0].encode("utf-8")), reverse=True)
        if len(entries) < 100:  # &entries[0..100] (:50)
            P.panic("range end index 100 out of range for slice (fasta_statistics.rs:50)")
        for bc, n in entries[:100]:
            P.out.append("- %s: %d\n" % (bc, n))
    except _Exit as e:
        return e.code, _enc(P.out), _enc(P.err), b""
    return 0, _enc(P.out), _enc(P.err), b""


def interleave(a: bytes, b: bytes):
    """fasta_interleave.rs:14-35."""
    P = _Proc()
    f1, f2 = _Reader(P, a), _Reader(P, b)
    try:
        while True:
            ok, line = f1.read_line()
            if not ok:
                break
            if line.startswith("@"):
                lines = 4
            elif line.startswith(">"):
                lines = 2
            else:
                P.error("Line is not FASTA/FASTQ format: " + line)  # :19
            P.out.append(line)
            for _ in range(lines - 1):
                P.out.append(f1.read_line()[1])
            _, line = f2.read_line()
            if (lines == 4 and not line.startswith("@")) or (lines == 2 and not line.startswith(">")):  # :26-29
                P.error("Input files do not share a consistent format.")
            P.out.append(line)
            for _ in range(lines - 1):
                P.out.append(f2.read_line()[1])
    except _Exit as e:
        return e.code, _enc(P.out), _enc(P.err), b""
    return 0, _enc(P.out), _enc(P.err), b""


def deinterleave(data: bytes):
    """fasta_deinterleave.rs:14-39: stdout stays empty, the two outputs are <prefix>_1.fq.gz (first of the return's
    outputs, in the stdout slot) and <prefix>_2.fq.gz (second output)."""
    P = _Proc()
    f = _Reader(P, data)
    out2 = []
    try:
        while True:
            ok, line = f.read_line()
            if not ok:
                break
            if line.startswith("@"):
                lines = 4
            elif line.startswith(">"):
                lines = 2
            else:
                P.error("Line is not FASTA/FASTQ format: " + line)  # :23
            P.out.append(line)
            for _ in range(lines - 1):
                P.out.append(f.read_line()[1])
            _, line = f.read_line()
            if (lines == 4 and not line.startswith("@")) or (lines == 2 and not line.startswith(">")):  # :30-33
                P.error("Interleaved FASTA records are not in consistent format.")
            out2.append(line)
            for _ in range(lines - 1):
                out2.append(f.read_line()[1])
    except _Exit as e:
        return e.code, _enc(P.out), _enc(P.err), _enc(out2)
    return 0, _enc(P.out), _enc(P.err), _enc(out2)


def extract_dual_umi(data: bytes, first_bases: int):
    """fasta_extract_dual_umi.rs:14-72."""
    P = _Proc()
    f = _Reader(P, data)
    try:
        while True:
            ok, h1 = f.read_line()
            if not ok:
                break
            if h1.startswith("@"):
                fq = True
            elif h1.startswith(">"):
                fq = False
            else:
                P.error("Header is not valid FASTA/FASTQ:\n" + h1)  # :33
            q1 = q2 = ""
            if fq:  # :35-45
                s1 = f.read_line()[1]
                f.read_line()
                q1 = f.read_line()[1]
                h2 = f.read_line()[1]
                s2 = f.read_line()[1]
                f.read_line()
                q2 = f.read_line()[1]
                if not h2.startswith("@"):
                    P.error("Invalid FASTQ record found in input file.")
            else:  # :46-52
                s1 = f.read_line()[1]
                h2 = f.read_line()[1]
                s2 = f.read_line()[1]
                if not h2.startswith(">"):
                    P.error("Invalid FASTA record found in input file.")
            umi = _slice(P, s1, 0, first_bases, "of seq_1 (fasta_extract_dual_umi.rs:55)") + "+" + \
                _slice(P, s2, 0, first_bases, "of seq_2 (fasta_extract_dual_umi.rs:57)")
            if fq:  # :59-64
                P.out.append("%s RX:%s\n%s+\n%s%s RX:%s\n%s+\n%s" % (
                    _trim_end(h1), umi, _slice(P, s1, first_bases, None, "of seq_1"), _slice(P, q1, first_bases, None, "of qual_1"),
                    _trim_end(h2), umi, _slice(P, s2, first_bases, None, "of seq_2"), _slice(P, q2, first_bases, None, "of qual_2")))
            else:  # :65-69
                P.out.append("%s RX:%s\n%s%s RX:%s\n%s" % (
                    _trim_end(h1), umi, _slice(P, s1, first_bases, None, "of seq_1"),
                    _trim_end(h2), umi, _slice(P, s2, first_bases, None, "of seq_2")))
    except _Exit as e:
        return e.code, _enc(P.out), _enc(P.err), b""
    return 0, _enc(P.out), _enc(P.err), b""


def next_op(op: int, a: bytes, b: bytes | None = None, x: int = 0, y: int = 0):
    """Same numbering as pyoracle.next_op: 0 trim --first=x --last=y, 1 check, 2 statistics, 3 interleave(a, b),
    4 deinterleave, 5 extract dual umi --first-bases=x."""
    if op == 0:
        return trim_fixed(a, x, y)
    if op == 1:
        return check(a)
    if op == 2:
        return statistics(a)
    if op == 3:
        return interleave(a, b or b"")
    if op == 4:
        return deinterleave(a)
    return extract_dual_umi(a, x)
