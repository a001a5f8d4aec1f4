"""Python mirror of the reference's operator entry points on top of the C ABI.

Each public method corresponds to one `main()` of the reference and returns what that process
would have produced -- (exit_code, stdout, stderr) or, for demultiplex, a dict with the
decompressed content of every output file -- so the parity tests read like the reference's CLI:

    fasta trim by quality <fastq> <min_baseq>      -> Engine.trim_by_quality      (fasta_trim_by_quality.rs:10-50)
    fasta mask by quality <fastq> <min_baseq>      -> Engine.mask_by_quality      (fasta_mask_by_quality.rs:11-47)
    fasta add barcode <fastq> <barcode_file>       -> Engine.add_barcode          (fasta_add_barcode.rs:11-45)
    fasta demultiplex [...] <sheet> <fq1> [<fq2>]  -> Engine.demultiplex          (fasta_demultiplex.rs:30-265)

All per-read arithmetic runs in the CUDA kernels behind libseqkit_b200.so.  The code here only
moves bytes, parses the sample sheet, and formats the reference's messages (which quote header
text); it never computes an operator result on the CPU.
"""
from __future__ import annotations

import ctypes as C
import re

from . import _lib as L

_WS = b"\t\n\x0b\x0c\r "  # ASCII subset of Unicode White_Space (non-ASCII input is refused)
_BC_RE = re.compile(rb" BC:[ACGTNacgtn+]+")


class Unsupported(L.SkError):
    """Input this implementation refuses instead of guessing."""


class DemuxResult(dict):
    pass


def _check(ctx, rc, what):
    if rc != 0:
        raise L.SkError("%s failed (%d): %s" % (what, rc, (L.lib().sk_last_error(ctx) or b"").decode()))


class Engine:
    def __init__(self, device: int = 0, max_stream_bytes: int = 32 << 20, max_records: int = 1 << 19,
                 n_slots: int = 1, max_samples: int = 1024, aux_streams: bool = True, compact: bool = True,
                 line_ops: bool = False):
        self.lib = L.lib()
        lim = L.Limits(max_stream_bytes, max_records, n_slots, max_samples, 1 if aux_streams else 0,
                       0x200 if line_ops else 0)
        ctx = C.c_void_p()
        rc = self.lib.sk_ctx_create(device, C.byref(lim), C.byref(ctx))
        if rc != 0:
            raise L.SkError("sk_ctx_create failed (%d): %s" % (rc, (self.lib.sk_last_error(None) or b"").decode()))
        self.ctx = ctx
        self.max_records = max_records
        # per-sample output streams come from the device-side compaction (sk_demux_compact); False: from the
        # per-record slice tables and the host helper sk_demux_gather
        self.compact = compact
        self.S = 0
        self.L = 0

    def close(self):
        if self.ctx:
            self.lib.sk_ctx_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ raw calls
    def upload(self, which: int, data: bytes | None, slot: int = 0):
        data = data or b""
        _check(self.ctx, self.lib.sk_upload(self.ctx, slot, which, data, len(data)), "sk_upload")

    def wait(self, slot: int = 0) -> L.Result:
        res = L.Result()
        _check(self.ctx, self.lib.sk_wait(self.ctx, slot, C.byref(res)), "sk_wait")
        self.last_result = res  # .reserved: bit0 = warp engine ran, bit1 = re-run on the general engine, bit4 = line engine
        return res

    def fetch_out(self, which: int, n: int, slot: int = 0) -> bytes:
        if n == 0:
            return b""
        buf = C.create_string_buffer(n)
        _check(self.ctx, self.lib.sk_download_out(self.ctx, slot, which, buf, n), "sk_download_out")
        self.wait(slot)
        return buf.raw[:n]

    def set_sheet(self, barcodes: list[bytes]):
        S = len(barcodes)
        Lb = len(barcodes[0]) if S else 0
        assert all(len(b) == Lb for b in barcodes)
        rc = self.lib.sk_set_sheet(self.ctx, b"".join(barcodes), S, Lb)
        if rc == -6:
            raise Unsupported((self.lib.sk_last_error(self.ctx) or b"").decode())
        _check(self.ctx, rc, "sk_set_sheet")
        self.S, self.L = S, Lb

    @staticmethod
    def _refuse(res):
        names = {L.DATA_NON_ASCII: "non-ASCII bases or qualities, or invalid UTF-8",
                 L.DATA_RECORD_TOO_LONG: "record beyond the operator's length limit",
                 L.DATA_CHUNK_TOO_DENSE: "more records per chunk than the operator's engine has slots for",
                 L.DATA_MIXED_FORMAT: "mixed FASTA/FASTQ records", L.DATA_OUT_OVERFLOW: "output capacity exceeded",
                 L.DATA_TRUNCATED_FUSED: "fused trim+demux on a truncated header",
                 L.DATA_TOO_MANY_RECORDS: "more records in the batch than the context's max_records"}
        if res.status in names:
            raise Unsupported("%s (record %d)" % (names[res.status], res.err_record))

    # ------------------------------------------------------------------ trim / mask
    def _stream_op(self, fn, data: bytes, min_baseq: int):
        """Runs the op; on a data error replays the records before it (the reference has already
        printed them when it stops).  Returns (res_first, out_bytes, offset_of_failing_record)."""
        self.upload(L.IN_R1, data)
        _check(self.ctx, fn(self.ctx, 0, min_baseq, 0), "operator")
        res = self.wait()
        self._refuse(res)
        if res.status == L.DATA_OK:
            return res, self.fetch_out(0, res.out_bytes[0]), len(data)
        if res.err_record == 0:
            return res, b"", 0
        _check(self.ctx, fn(self.ctx, 0, min_baseq, res.err_record), "operator (replay)")
        rep = self.wait()
        return res, self.fetch_out(0, rep.out_bytes[0]), rep.consumed[0]

    def trim_by_quality(self, data: bytes, min_baseq: int):
        res, out, off = self._stream_op(self.lib.sk_trim_by_quality, data, min_baseq)
        if res.status == L.DATA_OK:
            return 0, out, b""
        if res.status == L.DATA_BAD_HEADER:  # fasta_trim_by_quality.rs:20-22
            return 255, out, b"ERROR: Invalid FASTQ format encountered.\n"
        if res.status == L.DATA_SEQ_SHORT:  # :47 slice panic; the header (:23) is already out
            nl = data.find(b"\n", off)
            hdr = data[off:] if nl < 0 else data[off:nl + 1]
            return 101, out + hdr, b"thread 'main' panicked: byte index out of range of seq (fasta_trim_by_quality.rs:47)\n"
        raise L.SkError("unexpected status %d" % res.status)

    def mask_by_quality(self, data: bytes, min_baseq: int):
        res, out, _ = self._stream_op(self.lib.sk_mask_by_quality, data, min_baseq)
        if res.status == L.DATA_OK:
            return 0, out, b""
        if res.status == L.DATA_BAD_HEADER:  # fasta_mask_by_quality.rs:21-23
            return 255, out, b"ERROR: Invalid FASTQ format encountered.\n"
        if res.status == L.DATA_LEN_MISMATCH:  # :35-37
            return 255, out, b"ERROR: Read sequence and base qualities are of different length.\n"
        raise L.SkError("unexpected status %d" % res.status)

    # ------------------------------------------------------------------ add barcode
    def add_barcode(self, fastq: bytes, barcodes: bytes):
        self.upload(L.IN_R1, fastq)
        self.upload(L.IN_AUX1, barcodes)
        _check(self.ctx, self.lib.sk_add_barcode(self.ctx, 0, 0), "sk_add_barcode")
        res = self.wait()
        self._refuse(res)
        if res.status == L.DATA_OK:
            return 0, self.fetch_out(0, res.out_bytes[0]), b""
        if res.status != L.DATA_BAD_FASTX_LINE:
            raise L.SkError("unexpected status %d" % res.status)
        out, off = b"", 0
        if res.err_record:
            _check(self.ctx, self.lib.sk_add_barcode(self.ctx, 0, res.err_record), "sk_add_barcode (replay)")
            rep = self.wait()
            out, off = self.fetch_out(0, rep.out_bytes[0]), rep.consumed[0]
        # The reference prints the BC'd header line and only then rejects it (fasta_add_barcode.rs:33,41-43).
        nl = fastq.find(b"\n", off)
        hdr = fastq[off:] if nl < 0 else fastq[off:nl + 1]
        bc = self._barcode_text(barcodes, res.err_record)
        out += hdr.rstrip(_WS) + b" BC:" + bc + b"\n"
        return 255, out, b"ERROR: Invalid FASTQ line:\n" + hdr + b"\n"

    @staticmethod
    def _barcode_text(barcodes: bytes, rec: int) -> bytes:
        """Barcode used by iteration `rec` (only needed to quote it in an error path)."""
        if not barcodes or barcodes[:1] not in (b"@", b">"):
            return b""
        lpr = 4 if barcodes[:1] == b"@" else 2
        lines = barcodes.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        nrec = (len(lines) + lpr - 1) // lpr
        i = min(rec, nrec - 1)
        j = i * lpr + 1
        return lines[j].rstrip(_WS) if j < len(lines) else b""

    # ------------------------------------------------------------------ demultiplex
    @staticmethod
    def parse_sheet(sheet: bytes):
        """fasta_demultiplex.rs:63-104.  Returns (names, barcodes, error_message_or_None)."""
        names, bcs = [], []
        barcode_len = 0
        pos = 0
        while pos < len(sheet):  # read_line: through '\n' inclusive (common.rs:106-112)
            nl = sheet.find(b"\n", pos)
            end = len(sheet) if nl < 0 else nl + 1
            line, pos = sheet[pos:end], end
            if line.startswith(b"#"):
                continue
            cols = line.strip(_WS).split(b"\t")
            if len(cols) < 2:
                continue
            if cols[1] == b"":
                return names, bcs, b"Sample %s has no barcode." % cols[0]
            if barcode_len == 0:
                barcode_len = len(cols[1])
            elif len(cols[1]) != barcode_len:
                return names, bcs, b"Barcodes in sample sheet must all be of same length."
            names.append(cols[0])
            bcs.append(cols[1])
        for s in range(len(names)):
            if names[s] in names[s + 1:]:
                return names, bcs, b"Sample %s is listed multiple times in sample sheet." % names[s]
        return names, bcs, None

    def _demux_call(self, opts):
        _check(self.ctx, self.lib.sk_demultiplex(self.ctx, 0, C.byref(opts)), "sk_demultiplex")
        if self.compact and not opts.no_output:
            _check(self.ctx, self.lib.sk_demux_compact(self.ctx, 0), "sk_demux_compact")
        res = self.wait()
        self._refuse(res)
        return res

    def _gather_files(self, res, names, paired):
        files = {}
        S = self.S
        if res.reserved & 8:  # compacted on the device: one slice per sample and mate
            for mate in range(2 if paired else 1):
                n = res.out_extent[mate]
                buf = C.create_string_buffer(max(n, 1))
                sl = (L.Slice * (S + 1))()
                _check(self.ctx, self.lib.sk_download_compact(self.ctx, 0, mate, buf, n), "sk_download_compact")
                _check(self.ctx, self.lib.sk_download_slices(self.ctx, 0, mate, sl), "sk_download_slices")
                self.wait()
                assert sl[S].offset == n and sl[S].len == res.out_bytes[mate], \
                    "slice table and outcome block disagree: mate %d extent %d / %d, payload %d / %d, reserved %d" % (
                        mate, sl[S].offset, n, sl[S].len, res.out_bytes[mate], res.reserved)
                raw = buf.raw
                for s, name in enumerate(names):
                    key = (name + (b"_%d.fq.gz" % (mate + 1) if paired else b".fq.gz")).decode()
                    files[key] = raw[sl[s].offset:sl[s].offset + sl[s].len]
            return files
        for mate in range(2 if paired else 1):
            n = res.out_extent[mate]
            out = self.fetch_out(mate, n)
            nc = res.n_chunks[mate]
            rows = (L.ChunkRow * max(nc, 1))()
            groups = (L.Group * max(res.n_records, 1))()
            _check(self.ctx, self.lib.sk_download_demux_tables(self.ctx, 0, mate, rows, groups, res.n_records), "tables")
            self.wait()
            outbuf = C.create_string_buffer(out, max(len(out), 1))
            for s, name in enumerate(names):
                need = self.lib.sk_demux_gather(outbuf, rows, groups, nc, s, None, 0)
                dst = C.create_string_buffer(max(need, 1))
                self.lib.sk_demux_gather(outbuf, rows, groups, nc, s, dst, need)
                key = (name + (b"_%d.fq.gz" % (mate + 1) if paired else b".fq.gz")).decode()
                files[key] = dst.raw[:need]
        return files

    def demultiplex(self, sheet: bytes, fastq_1: bytes, fastq_2: bytes | None = None, index1: bytes | None = None,
                    index2: bytes | None = None, dry_run: int = 0, fused_trim: int | None = None) -> DemuxResult:
        err = [b"Reading sample sheet...\n"]  # :58
        names, bcs, sheet_err = self.parse_sheet(sheet)
        paired = fastq_2 is not None
        R = DemuxResult(exit_code=0, stdout=b"", stderr=b"", files={}, counts=[0] * len(names),
                        names=[n.decode() for n in names], total=0, identified=0)
        if dry_run == 0:
            for n in names:  # GzipWriter::with_method creates the files while the sheet is read (:79-87)
                for key in ([n + b"_1.fq.gz", n + b"_2.fq.gz"] if paired else [n + b".fq.gz"]):
                    R["files"][key.decode()] = b""
        if sheet_err is not None:
            R["exit_code"] = 255
            R["stderr"] = b"".join(err) + b"ERROR: " + sheet_err + b"\n"
            return R
        err.append(b"Starting demultiplexing in %s end mode...\n" % (b"paired" if paired else b"single"))  # :106
        if not names:
            # barcode_len stays 0: the first read fails the length check (:148) or, on the index route with
            # empty index reads, nothing ever matches.  Keep the sheet non-empty on the device.
            raise Unsupported("empty sample sheet")
        self.set_sheet(bcs)
        self.upload(L.IN_R1, fastq_1)
        self.upload(L.IN_R2, fastq_2)
        use_index = 0
        if index1 is not None:
            self.upload(L.IN_AUX1, index1)
            use_index |= 1
        if index2 is not None:
            self.upload(L.IN_AUX2, index2)
            use_index |= 2
        opts = L.DemuxOpts(-1 if fused_trim is None else fused_trim, use_index, dry_run, 1 if dry_run else 0, 0)
        res = self._demux_call(opts)
        first = res
        if res.status != L.DATA_OK:
            if res.err_record == 0:
                res = None
            else:
                opts.rec_limit = res.err_record
                res = self._demux_call(opts)
        if res is not None:
            counts = (C.c_uint64 * (self.S + 2))()
            _check(self.ctx, self.lib.sk_download_counts(self.ctx, 0, counts), "counts")
            R["counts"] = list(counts[:self.S])
            R["total"], R["identified"] = counts[self.S], counts[self.S + 1]
            if dry_run == 0:
                R["files"].update(self._gather_files(res, names, paired))
            # WARNING lines in record order (:184-188)
            ev = (L.Event * max(res.n_events, 1))()
            n_ev = self.lib.sk_download_events(self.ctx, 0, ev, res.n_events)
            for e in ev[:max(n_ev, 0)]:
                if use_index:
                    parts = []
                    for src, off in ((index1 if index1 is not None else index2, e.bc_off),
                                     (index2, e.bc_off2)):
                        if off != 0xFFFFFFFF and src is not None:
                            nl = src.find(b"\n", off)
                            parts.append((src[off:] if nl < 0 else src[off:nl]).rstrip(_WS))
                    bc = parts[0] if parts else b""
                    for p in parts[1:]:
                        bc = bc + (b"+" if bc else b"") + p
                else:
                    bc = fastq_1[e.bc_off:e.bc_off + self.L]
                a, b = e.best_sample, e.equally_fine_sample
                err.append(b"WARNING: Sequenced barcode %s was an equally good match (%d mismatches) for samples %s (%s) "
                           b"and %s (%s), and was therefore not assigned to any sample.\n"
                           % (bc, e.mismatches, names[a], bcs[a], names[b], bcs[b]))
        if first.status != L.DATA_OK:
            off = res.consumed[L.IN_R1] if res is not None else 0
            nl = fastq_1.find(b"\n", off)
            hdr = fastq_1[off:] if nl < 0 else fastq_1[off:nl + 1]
            if first.status == L.DATA_BAD_HEADER:  # :118-120
                err.append(b"ERROR: Invalid FASTQ header line:\n" + hdr + b"\n")
                R["exit_code"] = 255
            elif first.status == L.DATA_NO_BC:  # :141
                err.append(b"ERROR: No BC:xxxx field found.\n")
                R["exit_code"] = 255
            elif first.status == L.DATA_BC_LEN:  # :148-150
                if use_index:
                    parts = []
                    for which, src in ((L.IN_AUX1, index1), (L.IN_AUX2, index2)):
                        if src is None:
                            continue
                        o = res.consumed[which] if res is not None else 0
                        lines = src[o:].split(b"\n", 2)
                        parts.append((lines[1] if len(lines) > 1 else b"").rstrip(_WS))
                    bc = parts[0]
                    for p in parts[1:]:
                        bc = bc + (b"+" if bc else b"") + p
                else:
                    m = _BC_RE.search(hdr)
                    bc = hdr[m.start() + 4:m.end()]
                err.append(b"ERROR: Sequenced barcode %s is of different length (%d nt) than barcodes in the sample "
                           b"sheet (%d nt).\n" % (bc, len(bc), self.L))
                R["exit_code"] = 255
            elif first.status == L.DATA_INDEX_ASSERT:  # :130,:134
                err.append(b"thread 'main' panicked: assertion failed (fasta_demultiplex.rs:130/134)\n")
                R["exit_code"] = 101
            elif first.status == L.DATA_SEQ_SHORT:
                raise Unsupported("fused trim: sequence shorter than kept quality prefix (record %d)" % first.err_record)
            else:
                raise L.SkError("unexpected status %d" % first.status)
            R["stderr"] = b"".join(err)
            return R
        if first.flags & L.FLAG_MATE_COUNT:
            raise Unsupported("mate / index files hold fewer records than <fastq_1>")
        if dry_run:
            raise Unsupported("--dry-run report is produced by the host binary")
        total, ident = R["total"], R["identified"]
        pct = b"NaN" if total == 0 else (b"%.1f" % (ident / total * 100.0))
        err.append(b"%d / %d (%s%%) clusters carried a barcode matching one of the provided samples.\n"
                   % (ident, total, pct))  # :263-264
        R["stderr"] = b"".join(err)
        return R

    # ------------------------------------------------------------------ the line engine
    def line_op(self, op: int, a: bytes, b: bytes | None = None, x: int = 0, y: int = 0):
        """Raw call of one line operator (sk_line_op): returns (sk_result, output 0, output 1).  The reference's
        messages for the failing outcomes are composed by the `fasta` binary (seqkit_b200/host/fasta_main.cpp);
        tests/test_gpu_lineops.py compares that binary with the oracle's CLI."""
        self.upload(L.IN_R1, a)
        self.upload(L.IN_R2, b)
        _check(self.ctx, self.lib.sk_line_op(self.ctx, 0, op, x, y, 0), "sk_line_op")
        res = self.wait()
        self._refuse(res)
        out0 = self.fetch_out(0, res.out_bytes[0])
        out1 = self.fetch_out(1, res.out_bytes[1]) if op == L.LOP_DEINTERLEAVE else b""
        return res, out0, out1

    def statistics(self, data: bytes):
        """fasta statistics (fasta_statistics.rs:13-52): (total records, {barcode: count}) of a well-formed file."""
        res, _, _ = self.line_op(L.LOP_STATS, data)
        n = C.c_uint32()
        self.lib.sk_download_stats(self.ctx, 0, None, 0, C.byref(n))
        ents = (L.StatEntry * max(n.value, 1))()
        _check(self.ctx, self.lib.sk_download_stats(self.ctx, 0, ents, n.value, C.byref(n)), "sk_download_stats")
        return res, {data[e.off:e.off + e.len]: e.count for e in ents[:n.value]}

    # ------------------------------------------------------------------ synthetic workloads (bench / tests)
    def synth(self, which: int, n_pairs: int, seed: int = 1, first_pair: int = 0, read_len: int = 150, mate: int = 1,
              with_bc: bool = False, qual_profile: int = 0, p_sub_ppm: int = 10000, p_n_ppm: int = 5000,
              p_random_ppm: int = 20000, slot: int = 0) -> int:
        """Generates FASTQ text on the device into input stream `which`; returns its length in bytes."""
        spec = L.SynthSpec(seed, first_pair, n_pairs, read_len, mate, 1 if with_bc else 0, qual_profile, p_sub_ppm,
                           p_n_ppm, p_random_ppm, 0)
        n = C.c_uint64()
        _check(self.ctx, self.lib.sk_synth_fastq(self.ctx, slot, which, C.byref(spec), C.byref(n)), "sk_synth_fastq")
        return n.value

    def download_in(self, which: int, n: int, slot: int = 0) -> bytes:
        if n == 0:
            return b""
        buf = C.create_string_buffer(n)
        _check(self.ctx, self.lib.sk_download_in(self.ctx, slot, which, buf, n), "sk_download_in")
        return buf.raw[:n]
