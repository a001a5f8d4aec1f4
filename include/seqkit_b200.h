/*
 * seqkit_b200.h -- C ABI of libseqkit_b200.so: the B200 (sm_100a) implementation of
 * annalam/seqkit's per-read FASTQ batch path.
 *
 * The reference has no plugin/FFI interface: every operator is the body of a `main()` that
 * calls `FileReader::read_line` four times per record and prints.  This header
 * is the boundary a Rust `-sys` crate (or the C++ host binary in seqkit_b200/host/) binds: the
 * host batcher fills pinned multi-MB buffers with raw FASTQ bytes, hands them to a *slot*, and
 * gets back output bytes plus small tables.  Plain C types only; no exceptions cross it.
 *
 * Each entry point cites the reference code it replaces (paths relative to
 * /root/reference/src/).  There is NO CPU fallback behind any of them: without a CUDA device
 * sk_ctx_create fails with SK_E_CUDA.
 */
#ifndef SEQKIT_B200_H
#define SEQKIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SK_ABI_VERSION 2

/* ---- API return codes (every function returning int) ------------------------------------- */
#define SK_OK 0
#define SK_E_INVALID (-1)     /* bad argument / call order */
#define SK_E_CUDA (-2)        /* CUDA error or no device; text in sk_last_error */
#define SK_E_NOMEM (-3)
#define SK_E_TOO_LARGE (-4)   /* input larger than the slot capacity given to sk_ctx_create */
#define SK_E_NO_SHEET (-5)    /* sk_demux before sk_set_sheet */
#define SK_E_UNSUPPORTED (-6) /* e.g. sheet alphabet/length outside what the matcher packs */

/* ---- data outcomes: sk_result.status -------------------------------------------------------
 * Conditions the reference treats as fatal while streaming.  The kernels process every record
 * with index < err_record exactly as the reference would have before stopping, so the host
 * prints the reference's message AFTER flushing that output (same observable order). */
#define SK_DATA_OK 0
#define SK_DATA_BAD_HEADER 1      /* line 0 lacks '@'  (fasta_trim_by_quality.rs:20, fasta_mask_by_quality.rs:21, fasta_demultiplex.rs:118) */
#define SK_DATA_LEN_MISMATCH 2    /* mask: seq/qual length differ (fasta_mask_by_quality.rs:35-37) */
#define SK_DATA_SEQ_SHORT 3       /* trim: &seq[..k] out of range -> Rust panic, status 101 (fasta_trim_by_quality.rs:47) */
#define SK_DATA_NO_BC 4           /* demux: no " BC:x" field (fasta_demultiplex.rs:141) */
#define SK_DATA_BC_LEN 5          /* demux: barcode length != sheet (fasta_demultiplex.rs:148-150) */
#define SK_DATA_INDEX_ASSERT 6    /* demux --index: '@' / '+' assertion -> panic 101 (fasta_demultiplex.rs:130,134) */
#define SK_DATA_BAD_FASTX_LINE 7  /* add barcode: header is neither '@' nor '>' (fasta_add_barcode.rs:41-43) */
#define SK_DATA_NO_PLUS 8         /* check: line 2 of a FASTQ record lacks '+' (fasta_check.rs:58-61) */
#define SK_DATA_INCONSISTENT 9    /* interleave / deinterleave / extract dual umi: the second record of a pair does not start
                                     like the first (fasta_interleave.rs:26-29, fasta_deinterleave.rs:30-33, fasta_extract_dual_umi.rs:42-52) */
#define SK_DATA_QUAL_SHORT 10     /* trim / extract dual umi: &qual[..] out of range -> Rust panic, status 101 (fasta_trim.rs:41) */
/* Inputs this implementation refuses instead of guessing: */
#define SK_DATA_NON_ASCII 32      /* a byte >= 0x80 in the batch */
#define SK_DATA_RECORD_TOO_LONG 33
#define SK_DATA_CHUNK_TOO_DENSE 34
#define SK_DATA_MIXED_FORMAT 35   /* '@' and '>' records mixed in one add-barcode input */
#define SK_DATA_OUT_OVERFLOW 36   /* output capacity of the slot exceeded */
#define SK_DATA_TRUNCATED_FUSED 37 /* fused trim+demux on a header line without '\n' */
#define SK_DATA_HASH_COLLISION 38  /* statistics: two different barcodes with one 64-bit hash (reported, never merged) */
#define SK_DATA_TOO_MANY_RECORDS 39 /* the batch holds more records than sk_limits.max_records (err_record = max_records): nothing of it
                                      is valid; re-issue the call with rec_limit = err_record (sk_result.consumed[] then tells where
                                      the next batch starts) or create the context with a larger limit */

/* sk_result.flags */
#define SK_FLAG_MATE_COUNT 1u      /* mate/index streams hold fewer records than stream 0 */
#define SK_FLAG_EVENTS_OVERFLOW 2u /* more ambiguity events than the slot can hold */

/* stream indices inside a slot */
#define SK_IN_R1 0    /* reads / mate 1            (<fastq_file>, <fastq_1>) */
#define SK_IN_R2 1    /* mate 2                    (<fastq_2>) */
#define SK_IN_AUX1 2  /* --index1 FASTQ, or the <barcode_file> of `add barcode` */
#define SK_IN_AUX2 3  /* --index2 FASTQ */
#define SK_N_INPUTS 4

typedef struct sk_ctx sk_ctx;

typedef struct sk_limits {
    uint64_t max_stream_bytes; /* capacity of each input stream of a slot (< 4 GiB) */
    uint64_t max_records;      /* records (pairs) per batch */
    uint32_t n_slots;          /* independent stream slots for H2D / kernel / D2H overlap (>= 1) */
    uint32_t max_samples;      /* largest sample sheet (0 = no demultiplexing) */
    uint32_t aux_streams;      /* 0: allocate only R1/R2; 1: also AUX1/AUX2 (index reads, barcode file) */
    uint32_t reserved;         /* low byte, tuning: 0 default geometry, 1 = 16 KiB chunks, 2 = 32 KiB chunks; bit 8 (0x100): no
                                  buffers for sk_demux_compact; bit 9 (0x200): buffers for the line operators (sk_line_op) */
} sk_limits;

typedef struct sk_result {
    int32_t status;       /* SK_DATA_* of the first failing record, else SK_DATA_OK */
    uint32_t flags;       /* SK_FLAG_* */
    uint64_t err_record;  /* record index at which `status` was raised */
    uint64_t n_records;   /* records processed (min of records present and rec_limit) */
    uint64_t n_lines[SK_N_INPUTS];
    uint64_t consumed[SK_N_INPUTS]; /* bytes of each input covered by the processed records */
    uint64_t out_bytes[2];          /* payload bytes written per output stream */
    uint64_t out_extent[2];         /* extent of the output buffer in use (demux chunks are 16 B aligned) */
    uint64_t total_reads;           /* fasta_demultiplex.rs:108,169 */
    uint64_t identified_reads;      /* fasta_demultiplex.rs:109,177 */
    uint32_t n_chunks[2];           /* rows of the demux slice table per output stream */
    uint32_t n_events;              /* ambiguity events (fasta_demultiplex.rs:184-188) */
    uint32_t gpu_launches;          /* kernels this call enqueued */
    uint32_t reserved;              /* diagnostic: bit0 = the warp engine (sk_warp.cu) ran, bit1 = the
                                       operator met something outside its limits and was re-run on the general engine, bit2 = mask by
                                       quality met a record that changes its length (or fails), or add barcode a record that does not
                                       grow like the first, and the pass was re-run in its ordered form, bit3 =
                                       the demultiplex output was compacted per sample (sk_demux_compact), bit4 = trim / mask by
                                       quality, add barcode or demultiplex ended up on the line engine (long or dense records,
                                       UTF-8 header lines) */
    float pass_ms[SK_N_INPUTS];     /* device time of the chunk-engine kernel over each input stream
                                       (CUDA events on the slot's stream; only with sk_set_profiling) */
} sk_result;

/* One "equally good match" occurrence; the host prints the WARNING (fasta_demultiplex.rs:184-188)
 * in record order.  Header route: bc_off = byte offset of the observed barcode in SK_IN_R1 and
 * bc_off2 = 0xFFFFFFFF.  --index route: bc_off / bc_off2 = byte offsets of the sequence lines of the
 * first / second index read in their streams (0xFFFFFFFF when absent). */
typedef struct sk_event {
    uint32_t record;
    uint32_t bc_off;
    uint32_t bc_off2;
    int16_t best_sample;
    int16_t equally_fine_sample;
    uint32_t mismatches;
} sk_event;

typedef struct sk_demux_opts {
    int32_t fused_trim_min_baseq; /* -1: plain demultiplex; 0..255: trim by quality first (north-star config 5) */
    uint32_t use_index;           /* bit0: AUX1 holds --index1, bit1: AUX2 holds --index2 */
    uint64_t rec_limit;           /* process only records < rec_limit (0 = all; --dry-run=N / error replay) */
    uint32_t no_output;           /* 1: count only (dry run, fasta_demultiplex.rs:77-78,179) */
    uint32_t reserved;
} sk_demux_opts;

/* ---- context ------------------------------------------------------------------------------ */
int sk_abi_version(void);
/* Allocates device buffers, streams and tables for `lim` on CUDA device `device`. */
int sk_ctx_create(int device, const sk_limits *lim, sk_ctx **out);
void sk_ctx_destroy(sk_ctx *ctx);
const char *sk_last_error(const sk_ctx *ctx); /* ctx may be NULL: last create failure */
/* cudaStream_t of a slot (so callers can record their own events on it). */
void *sk_slot_stream(sk_ctx *ctx, uint32_t slot);
/* Upper bound of sk_result.n_chunks for this context (rows of the demux slice tables). */
uint32_t sk_max_chunks(sk_ctx *ctx);
/* Diagnostic: per-phase SM cycles (summed over chunks, one timing thread per CTA) of the last kernel
 * over input `which`; all zero unless the library was built with -DSK_PHASE_TIMING. */
int sk_debug_phase_cycles(sk_ctx *ctx, uint32_t slot, uint32_t which, uint64_t out[16]);
/* on != 0: bracket every kernel with CUDA events so that sk_wait can fill sk_result.pass_ms. */
int sk_set_profiling(sk_ctx *ctx, int on);

/* Number of CUDA devices this process sees (0 when there is none). */
int sk_device_count(void);
/* Restricts the calling thread to the CPUs of the NUMA node `device` hangs off (sysfs); returns the node, or
 * -1 when the topology is not exposed.  Host threads that feed a GPU (and the pinned buffers they touch first)
 * belong next to that GPU's PCIe root: on a two-socket 8-GPU box the cross-socket hop halves H2D/D2H. */
int sk_bind_thread_to_device(int device);
/* Pinned (page-locked, portable) host memory for the batcher's multi-MB buffers, so that a host without its own
 * CUDA binding (the Rust crate, the C++ `fasta` binary) gets asynchronous H2D/D2H copies; allocated on the
 * NUMA node of the context's device. */
void *sk_pinned_alloc(sk_ctx *ctx, uint64_t bytes);
void sk_pinned_free(sk_ctx *ctx, void *p);
/* Capacity in bytes of each output stream of a slot (upper bound of sk_result.out_extent). */
uint64_t sk_out_capacity(sk_ctx *ctx);

/* ---- inputs ------------------------------------------------------------------------------- */
/* Device address / capacity of an input stream buffer (for producers that write on the device). */
void *sk_slot_in(sk_ctx *ctx, uint32_t slot, uint32_t which);
uint64_t sk_slot_in_capacity(sk_ctx *ctx, uint32_t slot, uint32_t which);
/* Async H2D of `n` raw FASTQ bytes (pin `host` for overlap); replaces FileReader (common.rs:88-112). */
int sk_upload(sk_ctx *ctx, uint32_t slot, uint32_t which, const void *host, uint64_t n);
/* Declare the length of an input already resident in the slot buffer (n = 0 clears the stream). */
int sk_set_input_len(sk_ctx *ctx, uint32_t slot, uint32_t which, uint64_t n);

/* ---- sample sheet ------------------------------------------------------------------------- */
/* `barcodes` = S rows of L raw bytes in sheet order (Sample.barcode, fasta_demultiplex.rs:23-28,63-95).
 * Packs them into bit-planes for the matcher that replaces barcode_diff (fasta_demultiplex.rs:269-277). */
int sk_set_sheet(sk_ctx *ctx, const uint8_t *barcodes, uint32_t S, uint32_t L);

/* ---- operators (async on the slot's stream; results via sk_wait) --------------------------- */
/* fasta_trim_by_quality.rs:10-50 on SK_IN_R1 -> output stream 0. */
int sk_trim_by_quality(sk_ctx *ctx, uint32_t slot, uint32_t min_baseq, uint64_t rec_limit);
/* fasta_mask_by_quality.rs:11-47 on SK_IN_R1 -> output stream 0. */
int sk_mask_by_quality(sk_ctx *ctx, uint32_t slot, uint32_t min_baseq, uint64_t rec_limit);
/* fasta_add_barcode.rs:11-45: SK_IN_R1 = <fastq_file>, SK_IN_AUX1 = <barcode_file> -> output stream 0. */
int sk_add_barcode(sk_ctx *ctx, uint32_t slot, uint64_t rec_limit);
/* fasta_demultiplex.rs:117-249: SK_IN_R1 (+SK_IN_R2, +AUX index reads) -> output streams 0/1,
 * slice tables, counters, events. */
int sk_demultiplex(sk_ctx *ctx, uint32_t slot, const sk_demux_opts *opts);

/* The line engine: a global line table (newline scan with block prefix sums) and the
 * record-shuffling operators on top of it.  SK_IN_R1 (+ SK_IN_R2 for interleave) -> output stream 0 (+ stream 1 for
 * deinterleave: sk_result.out_bytes[1]).  Framing follows the stream's first byte ('@': 4 lines per record, '>': 2);
 * a record that starts with the other character is SK_DATA_MIXED_FORMAT.  On a data outcome the records before
 * err_record are in the output (sk_result.n_records = their number), as the reference has printed them by then.
 *   SK_LOP_TRIM          fasta trim --first=x --last=y             (fasta_trim.rs:24-47)
 *   SK_LOP_CHECK         fasta check                               (fasta_check.rs:49-70); no output
 *   SK_LOP_STATS         fasta statistics                          (fasta_statistics.rs:13-52); sk_download_stats
 *   SK_LOP_INTERLEAVE    fasta interleave <R1> <R2>                (fasta_interleave.rs:14-35)
 *   SK_LOP_DEINTERLEAVE  fasta deinterleave                        (fasta_deinterleave.rs:14-39); n_records = pairs
 *   SK_LOP_DUAL_UMI      fasta extract dual umi --first-bases=x    (fasta_extract_dual_umi.rs:14-72); n_records = pairs
 * Needs a context created with sk_limits.reserved bit 9. */
#define SK_LOP_TRIM 0
#define SK_LOP_CHECK 1
#define SK_LOP_STATS 2
#define SK_LOP_INTERLEAVE 3
#define SK_LOP_DEINTERLEAVE 4
#define SK_LOP_DUAL_UMI 5
int sk_line_op(sk_ctx *ctx, uint32_t slot, uint32_t op, uint32_t x, uint32_t y, uint64_t rec_limit);
typedef struct sk_stat_entry {
    uint32_t off; /* one occurrence (the earliest) of the barcode in SK_IN_R1 */
    uint32_t len;
    uint64_t count;
} sk_stat_entry;
/* Distinct " BC:[ACGTNacgtn]+" barcodes of the last SK_LOP_STATS call, in no particular order; *n = their number
 * (SK_E_TOO_LARGE when cap is smaller: call again with room for *n). */
int sk_download_stats(sk_ctx *ctx, uint32_t slot, sk_stat_entry *entries, uint32_t cap, uint32_t *n);

/* Blocks until the slot's stream is idle and returns the outcome of the last operator. */
int sk_wait(sk_ctx *ctx, uint32_t slot, sk_result *res);

/* ---- outputs ------------------------------------------------------------------------------ */
/* Device address of output stream `which` (0/1) of the last operator. */
const void *sk_out_dev(sk_ctx *ctx, uint32_t slot, uint32_t which);
/* D2H of the first `n` bytes of an output stream (call after sk_wait, or with n = capacity bound
 * before it; the copy is ordered on the slot's stream). */
int sk_download_out(sk_ctx *ctx, uint32_t slot, uint32_t which, void *host, uint64_t n);

/* Demultiplex side tables (valid after sk_wait).  The kernels write the emitted records of a chunk back
 * to back as a sequence of *groups* -- runs of bytes that belong to one sample, in input order inside a
 * sample (the warp engine emits one group per record in input order, one row per round of 32 records, four
 * rows per tile, unused ones empty; the line engine one row per 32 records; the general engine groups a chunk's
 * records by sample).  For output stream m, row c describes chunk (or round) c: its groups are
 * groups[first_group .. first_group+n_groups), laid out back to back from byte `base` of output
 * stream m.  Appending, for every sample, its groups over c = 0..n_chunks-1 gives that sample's file
 * content in input order (fasta_demultiplex.rs:196-238). */
typedef struct sk_group {
    uint16_t sample;
    uint16_t len; /* bytes */
} sk_group;
typedef struct sk_chunk_row {
    uint64_t base;
    uint32_t first_group;
    uint32_t n_groups;
} sk_chunk_row;
/* rows[n_chunks], groups[n_records] (n_records = sk_result.n_records of the operator). */
int sk_download_demux_tables(sk_ctx *ctx, uint32_t slot, uint32_t which, sk_chunk_row *rows, sk_group *groups,
                             uint64_t n_records);
/* counts[0..S) = Sample.total_reads (:27,178), counts[S] = total_reads, counts[S+1] = identified_reads. */
int sk_download_counts(sk_ctx *ctx, uint32_t slot, uint64_t *counts /*[S+2]*/);
const void *sk_counts_dev(sk_ctx *ctx, uint32_t slot); /* device u64[S+2], for an in-place all-reduce */
int sk_download_events(sk_ctx *ctx, uint32_t slot, sk_event *events, uint32_t cap); /* sorted by record */
/* Per-record sample assignment of the last sk_demultiplex: >=0 sample, -1 no match, -2 ambiguous. */
int sk_download_assign(sk_ctx *ctx, uint32_t slot, int16_t *assign, uint64_t n_records);

/* Host helper: appends sample `s`'s groups (host copies of one output stream and its tables) to
 * `dst`; returns the number of bytes written, or the required size when dst_cap is too small. */
uint64_t sk_demux_gather(const uint8_t *out_host, const sk_chunk_row *rows, const sk_group *groups, uint32_t n_chunks,
                         uint32_t s, uint8_t *dst, uint64_t dst_cap);

/* Device-side stable per-sample compaction (fasta_demultiplex.rs:196-238: a sample's file is the sample's records
 * in input order).  Enqueued after sk_demultiplex on the same slot, it regroups the emitted records of both
 * mates so that every sample's records are one contiguous run of bytes in input order; the host then appends S
 * slices per batch and mate instead of one piece per record.  After sk_wait, sk_result.out_extent[m] is the
 * extent of the compacted buffer of mate m, sk_result.reserved has bit 3 set, and slices[s] = {offset, len} of
 * sample s inside it (every run starts on a 128-byte line); slices[S] = {extent, payload bytes}.
 * Returns SK_E_UNSUPPORTED for sheets of more than 4096 samples (use the slice tables above). */
typedef struct sk_slice {
    uint64_t offset;
    uint64_t len;
} sk_slice;
int sk_demux_compact(sk_ctx *ctx, uint32_t slot);
const void *sk_compact_dev(sk_ctx *ctx, uint32_t slot, uint32_t which);
int sk_download_compact(sk_ctx *ctx, uint32_t slot, uint32_t which, void *host, uint64_t n);
int sk_download_slices(sk_ctx *ctx, uint32_t slot, uint32_t which, sk_slice *slices /*[S+1]*/);

/* ---- multi-GPU ---------------------------------------------------------------------------- */
/* Sums the S+2 counters of this slot across ranks in place with one ncclAllReduce(sum, u64) on the
 * slot's stream.  `nccl_comm` is an ncclComm_t created by the caller; the only collective on the
 * path. */
int sk_allreduce_counts(sk_ctx *ctx, uint32_t slot, void *nccl_comm);
/* Thin wrappers over ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy (libnccl.so.2 is resolved at
 * run time) so that a host without its own NCCL binding can build the communicator: rank 0 fills
 * `id128` (128 bytes), every rank receives it out of band and calls sk_nccl_comm_init. */
/* One process driving several GPUs (the `fasta` binary: one context per device, batches dealt round-robin):
 * sk_counts_accumulate adds the counters of a finished batch (call it after sk_wait, once the batch's outcome
 * is final) to the context's device-side run totals; sk_allreduce_totals merges the totals of `n` contexts with
 * one grouped ncclAllReduce(sum, u64, S+2) over NVLink (ncclCommInitAll; n == 1: nothing to merge), after which
 * every context holds the run's counters (fasta_demultiplex.rs:108-109,177-178,263-264). */
int sk_counts_accumulate(sk_ctx *ctx, uint32_t slot);
int sk_totals_reset(sk_ctx *ctx);
int sk_allreduce_totals(sk_ctx **ctxs, int n);
int sk_download_totals(sk_ctx *ctx, uint64_t *totals /*[S+2]*/);
int sk_nccl_unique_id(sk_ctx *ctx, void *id128);
int sk_nccl_comm_init(sk_ctx *ctx, const void *id128, int nranks, int rank, void **comm);
int sk_nccl_comm_destroy(sk_ctx *ctx, void *comm);

/* ---- synthetic workloads (bench / tests) ------------------------------------ */
typedef struct sk_synth_spec {
    uint64_t seed;
    uint64_t first_pair;     /* global index of the first pair (shards regenerate any range) */
    uint64_t n_pairs;
    uint32_t read_len;       /* bases per read (150) */
    uint32_t mate;           /* 1 or 2 */
    uint32_t with_bc;        /* append " BC:<observed barcode>" to the header */
    uint32_t qual_profile;   /* 0: 3'-decaying + 5% crash, 1: RTA3 4-bin variant */
    uint32_t p_sub_ppm;      /* per-base substitution rate of the observed barcode */
    uint32_t p_n_ppm;        /* per-base N rate of the observed barcode */
    uint32_t p_random_ppm;   /* fraction of fully random barcodes */
    uint32_t reserved;
} sk_synth_spec;
/* Writes FASTQ text for `spec` into input stream `which` of the slot (on the device) and sets its
 * length; observed barcodes are drawn from the sheet given to sk_set_sheet.  Returns bytes via *n. */
int sk_synth_fastq(sk_ctx *ctx, uint32_t slot, uint32_t which, const sk_synth_spec *spec, uint64_t *n);
/* D2H copy of an input stream (to hand the same bytes to the CPU oracle). */
int sk_download_in(sk_ctx *ctx, uint32_t slot, uint32_t which, void *host, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* SEQKIT_B200_H */
