"""ctypes binding of libseqkit_b200.so (include/seqkit_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SK_LIB") or os.path.join(HERE, "libseqkit_b200.so")

SK_N_INPUTS = 4
IN_R1, IN_R2, IN_AUX1, IN_AUX2 = 0, 1, 2, 3

# sk_result.status values
DATA_OK, DATA_BAD_HEADER, DATA_LEN_MISMATCH, DATA_SEQ_SHORT, DATA_NO_BC, DATA_BC_LEN, DATA_INDEX_ASSERT, \
    DATA_BAD_FASTX_LINE, DATA_NO_PLUS, DATA_INCONSISTENT, DATA_QUAL_SHORT = range(11)
LOP_TRIM, LOP_CHECK, LOP_STATS, LOP_INTERLEAVE, LOP_DEINTERLEAVE, LOP_DUAL_UMI = range(6)
DATA_NON_ASCII, DATA_RECORD_TOO_LONG, DATA_CHUNK_TOO_DENSE, DATA_MIXED_FORMAT, DATA_OUT_OVERFLOW, \
    DATA_TRUNCATED_FUSED = range(32, 38)
DATA_TOO_MANY_RECORDS = 39
FLAG_MATE_COUNT, FLAG_EVENTS_OVERFLOW = 1, 2


class SkError(RuntimeError):
    pass


class Limits(C.Structure):
    _fields_ = [("max_stream_bytes", C.c_uint64), ("max_records", C.c_uint64), ("n_slots", C.c_uint32),
                ("max_samples", C.c_uint32), ("aux_streams", C.c_uint32), ("reserved", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("flags", C.c_uint32), ("err_record", C.c_uint64), ("n_records", C.c_uint64),
                ("n_lines", C.c_uint64 * SK_N_INPUTS), ("consumed", C.c_uint64 * SK_N_INPUTS),
                ("out_bytes", C.c_uint64 * 2), ("out_extent", C.c_uint64 * 2), ("total_reads", C.c_uint64),
                ("identified_reads", C.c_uint64), ("n_chunks", C.c_uint32 * 2), ("n_events", C.c_uint32),
                ("gpu_launches", C.c_uint32), ("reserved", C.c_uint32), ("pass_ms", C.c_float * SK_N_INPUTS)]


class Event(C.Structure):
    _fields_ = [("record", C.c_uint32), ("bc_off", C.c_uint32), ("bc_off2", C.c_uint32), ("best_sample", C.c_int16),
                ("equally_fine_sample", C.c_int16), ("mismatches", C.c_uint32)]


class Group(C.Structure):
    _fields_ = [("sample", C.c_uint16), ("len", C.c_uint16)]


class ChunkRow(C.Structure):
    _fields_ = [("base", C.c_uint64), ("first_group", C.c_uint32), ("n_groups", C.c_uint32)]


class Slice(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("len", C.c_uint64)]


class StatEntry(C.Structure):
    _fields_ = [("off", C.c_uint32), ("len", C.c_uint32), ("count", C.c_uint64)]


class DemuxOpts(C.Structure):
    _fields_ = [("fused_trim_min_baseq", C.c_int32), ("use_index", C.c_uint32), ("rec_limit", C.c_uint64),
                ("no_output", C.c_uint32), ("reserved", C.c_uint32)]


class SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("first_pair", C.c_uint64), ("n_pairs", C.c_uint64), ("read_len", C.c_uint32),
                ("mate", C.c_uint32), ("with_bc", C.c_uint32), ("qual_profile", C.c_uint32), ("p_sub_ppm", C.c_uint32),
                ("p_n_ppm", C.c_uint32), ("p_random_ppm", C.c_uint32), ("reserved", C.c_uint32)]


# every symbol include/seqkit_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "sk_abi_version": (C.c_int, []),
    "sk_ctx_create": (C.c_int, [C.c_int, C.POINTER(Limits), C.POINTER(_P)]),
    "sk_ctx_destroy": (None, [_P]),
    "sk_last_error": (C.c_char_p, [_P]),
    "sk_slot_stream": (_P, [_P, C.c_uint32]),
    "sk_max_chunks": (C.c_uint32, [_P]),
    "sk_debug_phase_cycles": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "sk_set_profiling": (C.c_int, [_P, C.c_int]),
    "sk_device_count": (C.c_int, []),
    "sk_bind_thread_to_device": (C.c_int, [C.c_int]),
    "sk_pinned_alloc": (_P, [_P, C.c_uint64]),
    "sk_pinned_free": (None, [_P, _P]),
    "sk_out_capacity": (C.c_uint64, [_P]),
    "sk_slot_in": (_P, [_P, C.c_uint32, C.c_uint32]),
    "sk_slot_in_capacity": (C.c_uint64, [_P, C.c_uint32, C.c_uint32]),
    "sk_upload": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint64]),
    "sk_set_input_len": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint64]),
    "sk_set_sheet": (C.c_int, [_P, C.c_char_p, C.c_uint32, C.c_uint32]),
    "sk_trim_by_quality": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint64]),
    "sk_mask_by_quality": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint64]),
    "sk_add_barcode": (C.c_int, [_P, C.c_uint32, C.c_uint64]),
    "sk_demultiplex": (C.c_int, [_P, C.c_uint32, C.POINTER(DemuxOpts)]),
    "sk_line_op": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]),
    "sk_download_stats": (C.c_int, [_P, C.c_uint32, C.POINTER(StatEntry), C.c_uint32, C.POINTER(C.c_uint32)]),
    "sk_wait": (C.c_int, [_P, C.c_uint32, C.POINTER(Result)]),
    "sk_out_dev": (_P, [_P, C.c_uint32, C.c_uint32]),
    "sk_download_out": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint64]),
    "sk_download_demux_tables": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, C.c_uint64]),
    "sk_download_counts": (C.c_int, [_P, C.c_uint32, _P]),
    "sk_counts_dev": (_P, [_P, C.c_uint32]),
    "sk_download_events": (C.c_int, [_P, C.c_uint32, C.POINTER(Event), C.c_uint32]),
    "sk_download_assign": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64]),
    "sk_demux_gather": (C.c_uint64, [_P, _P, _P, C.c_uint32, C.c_uint32, _P, C.c_uint64]),
    "sk_demux_compact": (C.c_int, [_P, C.c_uint32]),
    "sk_compact_dev": (_P, [_P, C.c_uint32, C.c_uint32]),
    "sk_download_compact": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint64]),
    "sk_download_slices": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(Slice)]),
    "sk_allreduce_counts": (C.c_int, [_P, C.c_uint32, _P]),
    "sk_counts_accumulate": (C.c_int, [_P, C.c_uint32]),
    "sk_totals_reset": (C.c_int, [_P]),
    "sk_allreduce_totals": (C.c_int, [C.POINTER(_P), C.c_int]),
    "sk_download_totals": (C.c_int, [_P, _P]),
    "sk_nccl_unique_id": (C.c_int, [_P, _P]),
    "sk_nccl_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(_P)]),
    "sk_nccl_comm_destroy": (C.c_int, [_P, _P]),
    "sk_synth_fastq": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(SynthSpec), C.POINTER(C.c_uint64)]),
    "sk_download_in": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint64]),
}

_lib = None


def lib():
    """Loads libseqkit_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SkError("%s is missing: run `make` (or __graft_entry__.build()) first; there is no CPU fallback"
                          % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (rt, at) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = rt
            f.argtypes = at
        _lib = L
    return _lib
