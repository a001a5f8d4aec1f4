"""ctypes binding of oracle/fasta_oracle.c (the CPU parity oracle).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never import this from seqkit_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fasta_oracle.c")
    src2 = os.path.join(_HERE, "synth_host.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(src2)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _SO


class _Result(C.Structure):
    _fields_ = [("exit_code", C.c_int), ("out", C.c_void_p), ("out_n", C.c_size_t), ("err", C.c_void_p),
                ("err_n", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        for name in ("orc_trim_by_quality", "orc_mask_by_quality"):
            f = getattr(L, name)
            f.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.POINTER(_Result)]
            f.restype = C.c_int
        L.orc_add_barcode.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(_Result)]
        L.orc_add_barcode.restype = C.c_int
        L.orc_result_free.argtypes = [C.POINTER(_Result)]
        L.orc_demux.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int,
                                C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.c_uint64]
        L.orc_demux.restype = C.c_void_p
        for name, rt in (("orc_demux_exit_code", C.c_int), ("orc_demux_n_samples", C.c_size_t),
                         ("orc_demux_paired", C.c_int), ("orc_demux_outputs_created", C.c_int),
                         ("orc_demux_total", C.c_uint64), ("orc_demux_identified", C.c_uint64)):
            f = getattr(L, name)
            f.argtypes = [C.c_void_p]
            f.restype = rt
        L.orc_demux_sample_count.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_demux_sample_count.restype = C.c_uint64
        L.orc_demux_sample_name.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_demux_sample_name.restype = C.c_void_p
        L.orc_demux_sample_out.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_demux_sample_out.restype = C.c_void_p
        for name in ("orc_demux_stdout", "orc_demux_stderr"):
            f = getattr(L, name)
            f.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
            f.restype = C.c_void_p
        L.orc_demux_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _take(R: _Result):
    out = C.string_at(R.out, R.out_n) if R.out_n else b""
    err = C.string_at(R.err, R.err_n) if R.err_n else b""
    code = R.exit_code
    lib().orc_result_free(C.byref(R))
    return code, out, err


def trim_by_quality(data: bytes, min_baseq: int):
    R = _Result()
    lib().orc_trim_by_quality(data, len(data), min_baseq, C.byref(R))
    return _take(R)


def mask_by_quality(data: bytes, min_baseq: int):
    R = _Result()
    lib().orc_mask_by_quality(data, len(data), min_baseq, C.byref(R))
    return _take(R)


def add_barcode(fastq: bytes, barcodes: bytes):
    R = _Result()
    lib().orc_add_barcode(fastq, len(fastq), barcodes, len(barcodes), C.byref(R))
    return _take(R)


def _bytes_at(ptr, n):
    return C.string_at(ptr, n) if n else b""


def demultiplex(sheet: bytes, fastq_1: bytes, fastq_2: bytes | None = None, index1: bytes | None = None,
                index2: bytes | None = None, dry_run: int = 0):
    """Same return shape as oracle.restatement.demultiplex."""
    L = lib()

    def arg(b):
        return (b if b is not None else b""), (len(b) if b is not None else 0)

    r2, n2 = arg(fastq_2)
    i1, ni1 = arg(index1)
    i2, ni2 = arg(index2)
    D = L.orc_demux(sheet, len(sheet), fastq_1, len(fastq_1), r2, n2, int(fastq_2 is not None), i1, ni1,
                    int(index1 is not None), i2, ni2, int(index2 is not None), dry_run)
    try:
        n = C.c_size_t()
        S = L.orc_demux_n_samples(D)
        paired = bool(L.orc_demux_paired(D))
        created = bool(L.orc_demux_outputs_created(D))
        names, counts, files = [], [], {}
        for s in range(S):
            p = L.orc_demux_sample_name(D, s, C.byref(n))
            name = _bytes_at(p, n.value).decode("utf-8")
            names.append(name)
            counts.append(L.orc_demux_sample_count(D, s))
            if created:
                for mate in range(2 if paired else 1):
                    p = L.orc_demux_sample_out(D, s, mate, C.byref(n))
                    key = (name + "_%d.fq.gz" % (mate + 1)) if paired else (name + ".fq.gz")
                    files[key] = _bytes_at(p, n.value)
        p = L.orc_demux_stdout(D, C.byref(n))
        out = _bytes_at(p, n.value)
        p = L.orc_demux_stderr(D, C.byref(n))
        err = _bytes_at(p, n.value)
        return {"exit_code": L.orc_demux_exit_code(D), "stdout": out, "stderr": err, "files": files,
                "counts": counts, "names": names, "total": L.orc_demux_total(D),
                "identified": L.orc_demux_identified(D)}
    finally:
        L.orc_demux_free(D)


class _SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("first_pair", C.c_uint64), ("n_pairs", C.c_uint64), ("read_len", C.c_uint32),
                ("mate", C.c_uint32), ("with_bc", C.c_uint32), ("qual_profile", C.c_uint32), ("p_sub_ppm", C.c_uint32),
                ("p_n_ppm", C.c_uint32), ("p_random_ppm", C.c_uint32), ("reserved", C.c_uint32)]


def synth_fastq(n_pairs: int, seed: int = 1, first_pair: int = 0, read_len: int = 150, mate: int = 1,
                barcodes: list | None = None, qual_profile: int = 0, p_sub_ppm: int = 10000, p_n_ppm: int = 5000,
                p_random_ppm: int = 20000) -> bytes:
    """Host twin of the device generator (synth_host.c == seqkit_b200/csrc/sk_synth.cu): the same bytes for the
    same (seed, pair range, mate).  `barcodes` = the sheet's barcodes (observed barcodes are drawn from them)."""
    L = lib()
    L.orc_synth_fastq.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(_SynthSpec), C.c_char_p, C.c_uint32, C.c_uint32]
    L.orc_synth_fastq.restype = C.c_uint64
    S = len(barcodes) if barcodes else 0
    Lb = len(barcodes[0]) if S else 0
    spec = _SynthSpec(seed, first_pair, n_pairs, read_len, mate, 1 if S else 0, qual_profile, p_sub_ppm, p_n_ppm,
                      p_random_ppm, 0)
    cap = n_pairs * (64 + Lb + 2 * read_len) + 64
    buf = C.create_string_buffer(cap)
    n = L.orc_synth_fastq(buf, cap, C.byref(spec), b"".join(barcodes) if S else b"", S, Lb)
    assert n or n_pairs == 0, "synthetic buffer too small"
    return buf.raw[:n]


def next_op(op: int, a: bytes, b: bytes | None = None, x: int = 0, y: int = 0):
    """SURVEY.md section 8(f) operators of the oracle: 0 trim --first=x --last=y, 1 check, 2 statistics, 3 interleave(a, b),
    4 deinterleave, 5 extract dual umi --first-bases=x.  Returns (exit_code, stdout, stderr, second_output)."""
    L = lib()
    L.orc_next.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint64, C.POINTER(_Result),
                           C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.orc_next.restype = C.c_int
    R = _Result()
    o2, n2 = C.c_void_p(), C.c_size_t()
    L.orc_next(op, a, len(a), b or b"", len(b or b""), x, y, C.byref(R), C.byref(o2), C.byref(n2))
    out2 = C.string_at(o2, n2.value) if n2.value else b""
    code, out, err = _take(R)
    return code, out, err, out2
