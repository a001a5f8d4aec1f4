// sk_internal.h -- structures shared by the chunk-engine kernels and the host side of the C ABI.
#pragma once
#include <stdint.h>

namespace sk {

// Operators of the chunk engine.
enum Op : int {
    OP_SCAN = 0,    // per-record (seq_off, seq_len, flags) table of an index / barcode file
    OP_TRIM = 1,    // fasta_trim_by_quality.rs:10-50
    OP_MASK = 2,    // fasta_mask_by_quality.rs:11-47
    OP_ADDBC = 3,   // fasta_add_barcode.rs:11-45
    OP_DEMUX1 = 4,  // fasta_demultiplex.rs:117-212 (mate 1: extract, match, decide, emit)
    OP_DEMUX2 = 5,  // fasta_demultiplex.rs:215-238 (mate 2: emit)
    OP_LINE = 6,    // the line engine's operators (sk_lineops.cu): trim --first/--last, check, statistics, ...
};

// Geometry of one chunk-engine configuration.  WIN_MAX = NT*PPL*16 bytes of window per CTA:
// PRE bytes before the chunk (byte c0-1 decides whether a line starts at c0), the chunk itself and
// the overhang a record may extend into.  PPL is odd so that the per-thread LDS.128 of the newline
// scan are bank-conflict free (thread stride PPL*16 bytes).
struct CfgA {  // 16 KiB chunks, 4 warps: every warp has work in every phase of a 150 bp workload
    static constexpr int ID = 0;
    static constexpr int NT = 128;
    static constexpr int PPL = 11;
    static constexpr int WIN_MAX = NT * PPL * 16;       // 22528
    static constexpr int CHUNK = 16384;
    static constexpr int PRE = 16;
    static constexpr int OVERHANG = WIN_MAX - CHUNK - PRE;  // 6128
    static constexpr int MAXREC = 256;
    static constexpr int MAXLINES = 4 * MAXREC + 16;
    static constexpr int STAGE = 18432;                 // chunk outputs beyond this take the unstaged path
    static constexpr int MIN_CTAS = 3;
};
struct CfgB {  // 32 KiB chunks, 8 warps
    static constexpr int ID = 1;
    static constexpr int NT = 256;
    static constexpr int PPL = 9;
    static constexpr int WIN_MAX = NT * PPL * 16;       // 36864
    static constexpr int CHUNK = 32768;
    static constexpr int PRE = 16;
    static constexpr int OVERHANG = WIN_MAX - CHUNK - PRE;  // 4080
    static constexpr int MAXREC = 512;
    static constexpr int MAXLINES = 4 * MAXREC + 16;
    static constexpr int STAGE = 35840;
    static constexpr int MIN_CTAS = 2;
};
// Geometry of the warp engine (sk_warp.cu): a warp owns a tile, a lane owns a record.  Every lane scans
// UPL 16-byte units (odd: conflict-free LDS.128); the tile is the first `tile_lanes` lanes' bytes
// (KParams::tile_lanes, 29 unless the records are short), the remaining lanes' bytes are the overhang.
struct GeoW {
    static constexpr int ID = 2;
#ifdef SKW_WARPS
    static constexpr int WARPS = SKW_WARPS;  // experiment: one CTA of 16 warps per SM
#else
    static constexpr int WARPS = 8;
#endif
    static constexpr int NT = WARPS * 32;
    static constexpr int UPL = 25;
    static constexpr int LANE_BYTES = UPL * 16;           // 400
    static constexpr int WIN = 32 * LANE_BYTES;           // 12800
    static constexpr int TILE_LANES = 29;                 // default tile: 11600 B, overhang 1200 B
    static constexpr int ROUNDS = 4;                      // rounds of 32 records; one slice-table row each
    static constexpr int MAXREC = 32 * ROUNDS;
    static constexpr int MAXLINES = 4 * MAXREC + 8;
    static constexpr int MIN_CTAS = 16 / WARPS;
};
constexpr int FAST_CCOUNT_MAX = 1024;  // per-sample counters live in shared memory up to this many samples
inline int cfg_chunk_bytes(int cfg) { return cfg == CfgB::ID ? CfgB::CHUNK : CfgA::CHUNK; }

constexpr int LAYOUT_FAST_MAXREC = 128;  // the bit-mask layout handles chunks of up to this many records

// Entry of an OP_SCAN table: where the sequence line of record i is.
struct RecRef {
    uint32_t seq_off;  // byte offset of line 1 in its stream
    uint16_t seq_len;  // after trim_end (ASCII white space)
    uint16_t flags;    // RR_*
};
enum : uint16_t {
    RR_L0_AT = 1,    // line 0 starts with '@'
    RR_L0_GT = 2,    // line 0 starts with '>'
    RR_L2_PLUS = 4,  // line 2 starts with '+'
    RR_LONG = 8,     // seq_len did not fit
};

struct Event {  // == sk_event
    uint32_t record, bc_off, bc_off2;
    int16_t best, last;
    uint32_t mismatches;
};

// Device-side outcome block of one kernel launch (one per stream pass).
struct DevStats {
    unsigned long long n_lines;
    unsigned long long n_records;     // records processed
    unsigned long long out_bytes;     // payload bytes
    unsigned long long out_extent;    // bytes of the out buffer in use
    unsigned long long err_key;       // max over failing records of ~(record << 8 | kind); 0 = none
    unsigned long long consumed;      // byte offset just past the last processed record
    unsigned int flags;
    unsigned int n_events;
    unsigned int n_slow;              // records that took the brute-force match (diagnostic)
    unsigned int pad0;
    unsigned long long phase_cycles[16];  // -DSK_PHASE_TIMING: per-phase SM cycles summed over chunks (thread 0)
    // the two words every chunk hits with an atomic, each on a 128-byte line of its own
    unsigned long long compact_extent;  // per-sample compaction (sk_compact.cu): bytes of the compacted buffer in use
    unsigned long long pad1[7];
    unsigned long long out_cursor;    // demux: bump allocator
    unsigned long long pad2[15];
    unsigned int ticket;              // dynamic chunk counter
    unsigned int pad3[31];
};

// Pigeonhole index over the sample sheet (built on the host, sk_api.cu).  Samples are grouped in
// classes of identical care mask.  The cared positions of a class are split into two halves; a read
// within one mismatch of a sample equals that sample exactly on at least one half, so two hash
// lookups per class yield every sample at distance <= 1 (plus harmless extras).  Each candidate's
// true distance is then computed on the cared bytes, which makes the result identical to the loop
// over all samples in fasta_demultiplex.rs:157-166 whenever the minimum is <= 1 -- and when it is
// larger the read is unassigned either way (:172).
constexpr int HIDX_NWMAX = 16;  // key words per barcode (L <= 64)
struct HalfIdx {
    uint32_t n_classes;  // 0 = index unusable, every read takes the brute-force path
    uint32_t nw;         // words per key = ceil(L/4)
    uint32_t nwp;        // row stride of skeys / class rows, nw rounded up to 4
    uint32_t tsize;      // slots per table (power of two)
    const uint32_t *cls;     // per class 7 rows of nwp words: care | half0 {mask, mulA, mulB} | half1 {mask, mulA, mulB}
    const uint2 *table;      // [n_classes][2][tsize]: {tag, start | count << 16}; count == 0: empty slot
    const uint16_t *cand;    // candidate lists (sample indices, ascending)
    const uint32_t *skeys;   // [S][nwp]: cared bytes of every sample, 0 elsewhere
};
constexpr int HIDX_CLS_ROWS = 7;
// The same index in the form the warp engine probes: one 8-byte entry per slot that already names
// the first sample of the key, further samples of a key chained through `next`.
struct FastIdx {
    const uint2 *table;     // [n_classes][2][tsize]: {tag, (first sample + 1) | more << 16}; y == 0: empty slot
    const uint16_t *next;   // [n_classes][2][S]: next sample with the same half key, 0xFFFF = none
};

// Sample sheet in device memory (packed by the host, sk_api.cu).
struct SheetDev {
    const uint32_t *planes;  // brute-force fallback: S entries of {p0,p1,p2,care} (u32 x4, or u64 x4 when wide)
    const uint32_t *umask;   // S entries (u32) or 2*S (u64 as lo,hi): positions where the sheet has 'U'
    const uint8_t *lut;      // 256: bits 0-2 = 3-bit code (0 = matches no literal), bit 3 = [ACGTNacgtn+]
    uint32_t S, L, Umax, wide;
    uint32_t u_uniform;            // 1: every sample has its 'U' at the same positions (the usual sheet) ...
    unsigned long long u_mask;     // ... namely these: no per-record look-up of umask[]
    HalfIdx hidx;
    FastIdx fidx;
};

// Shared-memory carve-up (bytes), computed on the host and passed in KParams.
struct SmemLayout {
    uint32_t win, stage, ls, rec, umask, lut, ccount, masks, gtot, hcls, slow, misc, total;
};

// One run of same-sample records inside a chunk's output (demux slice table, sparse form).
struct Group {
    uint16_t sample;
    uint16_t len;  // bytes; a chunk's output is < 64 KiB
};
struct ChunkRow {
    unsigned long long base;  // byte offset of the chunk's output in its output stream
    uint32_t first_group;     // index of the chunk's first Group (== global index of its first record)
    uint32_t n_groups;
};

struct KParams {
    SmemLayout sl;
    // input stream
    const uint8_t *in;
    uint64_t n;
    uint32_t n_chunks;
    uint32_t lpr;         // lines per record: 4 (FASTQ) or 2 (FASTA)
    uint64_t max_records; // sk_limits.max_records: size of the per-record tables (assign, umi, groups, record tables); a batch
                          // with more records is refused (K_TOO_MANY) before anything is written past them
    uint64_t rec_limit;   // process records with index < rec_limit
    uint32_t final_batch; // 1: end of buffer is end of file (EOF semantics); 0: trailing partial record is left
    uint32_t min_baseq;
    int32_t fused_trim;   // demux: >=0 -> trim by quality with this threshold
    uint32_t head_char;   // add barcode: '@' or '>' (uniform over the file)
    uint32_t tile_lanes;  // warp engine: lanes whose bytes form the tile (GeoW)
    uint32_t unordered;   // warp engine, trim: tiles take output space from the cursor (in `out`, a scratch buffer) and
                          // note (base, length) in tile_out; a scan and a gather pass then write `final_out` in input order
    uint8_t *final_out;
    uint32_t inplace;     // warp engine, mask: a tile writes its records at their input offsets (a regular file keeps every
                          // length), no look-back on output bytes; any other record shape raises F_NEED_ORDERED
    // look-back state (zeroed before launch)
    uint64_t *tile_lines;
    uint64_t *tile_out;
    DevStats *stats;
    // output
    uint8_t *out;
    uint64_t out_cap;
    // demux
    SheetDev sheet;
    int16_t *assign;        // [max_records]
    uint8_t *umi;           // [max_records * Umax]
    Group *groups;          // [max_records]
    ChunkRow *rows;         // [n_chunks]
    unsigned long long *counts;  // [S + 2]
    Event *events;
    uint32_t events_cap;
    uint32_t n_index;       // demux: number of index streams (0 = header route)
    const DevStats *r1_stats;  // demux mate 2: outcome block of the mate-1 pass (n_records read on device)
    // external per-record tables (OP_SCAN output of other streams)
    RecRef *scan_out;       // OP_SCAN destination
    uint64_t scan_cap;
    const RecRef *ext_tab[2];
    const uint8_t *ext_data[2];
    const DevStats *ext_stats[2];  // n_records of the OP_SCAN pass that produced ext_tab[q]
    const uint4 *bc_inline;        // add barcode: the first 32 bytes of every record's barcode, zero-padded (2 x uint4 per
                                   // record, next to ext_tab[0]; nullptr = read the barcode where it lies)
};

// Data outcome kinds (== SK_DATA_* in include/seqkit_b200.h); the low byte of DevStats::err_key.
enum : unsigned {
    K_BAD_HEADER = 1, K_LEN_MISMATCH = 2, K_SEQ_SHORT = 3, K_NO_BC = 4, K_BC_LEN = 5, K_INDEX_ASSERT = 6,
    K_BAD_FASTX_LINE = 7, K_NON_ASCII = 32, K_TOO_LONG = 33, K_TOO_DENSE = 34, K_MIXED = 35, K_OUT_OVERFLOW = 36,
    K_TRUNC_FUSED = 37, K_TOO_MANY = 39,
    // line operators (sk_lineops.cu): K_NO_PLUS = 8, K_INCONSISTENT = 9, K_QUAL_SHORT = 10, K_HASH_COLLISION = 38
};
enum : unsigned { F_MATE_COUNT = 1u, F_EVENTS_OVERFLOW = 2u, F_NON_ASCII = 0x100u, F_NEED_GENERAL = 0x200u, F_NEED_ORDERED = 0x400u };

constexpr int REC_BYTES = 26;  // per-record plan fields, see sk_kernels.cu
template <class Cfg>
inline __host__ __device__ SmemLayout smem_layout(uint32_t S, uint32_t wide, uint32_t n_classes, uint32_t nwp) {
    SmemLayout L;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 15u) & ~15u; return r; };
    L.win = take(Cfg::WIN_MAX);
    L.stage = take(Cfg::STAGE + 32);
    L.ls = take((Cfg::MAXLINES + 8) * 2);
    L.rec = take(Cfg::MAXREC * REC_BYTES);
    L.umask = take(S * (wide ? 8u : 4u));
    L.lut = take(S ? 256 : 0);
    L.ccount = take(S * 4);
    L.masks = take(S * 16);
    L.gtot = take(S ? LAYOUT_FAST_MAXREC * 4 : 0);
    L.hcls = take(n_classes * HIDX_CLS_ROWS * nwp * 4);
    L.slow = take(S ? Cfg::MAXREC * 2 : 0);
    L.misc = take(512);
    L.total = o;
    return L;
}

// Launchers (sk_kernels.cu)
int launch_chunk_kernel(int cfg, int op, const KParams &p, int sm_count, void *stream, const char **err);
int chunk_kernel_smem_bytes(int cfg, uint32_t S, uint32_t wide, uint32_t n_classes, uint32_t nwp);
// Warp engine (sk_warp.cu)
bool warp_supported(int op, const KParams &p);
int launch_tile_gather(const KParams &p, int sm_count, void *stream, const char **err);  // after an `unordered` launch
int launch_warp_kernel(int op, const KParams &p, int sm_count, void *stream, const char **err);
// Line engine (sk_lineops.cu): trim --first/--last, check, statistics, interleave, deinterleave, extract dual umi
uint64_t lineops_work_bytes(uint64_t max_stream_bytes, uint64_t max_records);
int launch_lineop(int op, const uint8_t *in_a, uint64_t n_a, const uint8_t *in_b, uint64_t n_b, uint32_t lpr, uint32_t head, uint32_t x,
                  uint32_t y, uint64_t rec_limit, uint8_t *out0, uint8_t *out1, uint64_t out_cap, void *work, uint64_t max_stream_bytes,
                  uint64_t max_records, void *stats_tab, uint32_t h_cap, DevStats *st, int sm_count, void *stream, const char **err,
                  const RecRef *bc_tab = nullptr, const DevStats *bc_stats = nullptr);
// OP_SCAN's record table from the global line table (sk_lineops.cu): any record length and density
int launch_scan_table(const uint8_t *in, uint64_t n, uint32_t lpr, uint32_t head_char, uint64_t rec_limit, uint32_t final_batch,
                      RecRef *out, uint4 *inline32, uint64_t cap, void *work, int k, uint64_t max_stream_bytes, uint64_t max_records,
                      DevStats *st, int sm_count, void *stream, const char **err);
// Header-route demultiplex of one mate on the line engine (sk_lineops.cu): records of any length and density, UTF-8 headers
int launch_line_demux(int mate, const uint8_t *in, uint64_t n, uint64_t rec_limit, int fused_trim, const SheetDev &sheet, int16_t *assign,
                      uint8_t *umi, Group *groups, ChunkRow *rows, uint32_t max_rows, unsigned long long *counts, Event *events,
                      uint32_t events_cap, const DevStats *r1_stats, uint8_t *out, uint64_t out_cap, void *work,
                      uint64_t max_stream_bytes, uint64_t max_records, DevStats *st, int sm_count, void *stream, uint32_t *n_rows,
                      const char **err, uint32_t n_index = 0, const RecRef *const *ext_tab = nullptr,
                      const uint8_t *const *ext_data = nullptr, const DevStats *const *ext_stats = nullptr);
// Per-sample compaction of a demultiplex result (sk_compact.cu)
uint64_t compact_work_bytes(uint32_t max_rows, uint32_t S);
int launch_compact(const ChunkRow *rows, const Group *groups, uint32_t n_rows, uint32_t S, const uint8_t *src, uint8_t *dst,
                   uint64_t dst_cap, void *work, unsigned long long *slices, unsigned long long *piece_dst, DevStats *st,
                   int sm_count, void *stream, const char **err);

}  // namespace sk
