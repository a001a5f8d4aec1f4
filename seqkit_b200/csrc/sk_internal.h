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
};

// Geometry of one chunk-engine configuration.  WIN_MAX = NT*PPL*16 bytes of window per CTA:
// PRE bytes before the chunk (byte c0-1 decides whether a line starts at c0), the chunk itself and
// the overhang a record may extend into.
struct CfgBig {  // 2 CTAs/SM
    static constexpr int ID = 0;
    static constexpr int NT = 512;
    static constexpr int PPL = 5;                       // 16-byte pieces scanned per thread (odd => LDS.128 conflict-free)
    static constexpr int WIN_MAX = NT * PPL * 16;       // 40960
    static constexpr int CHUNK = 32768;
    static constexpr int PRE = 16;
    static constexpr int OVERHANG = WIN_MAX - CHUNK - PRE;  // 8176
    static constexpr int MAXLINES = 1792;
    static constexpr int MAXREC = 512;
    static constexpr int STAGE = 37888;                 // chunk outputs beyond this take the unstaged path
    static constexpr int MIN_CTAS = 2;
};
struct CfgSmall {  // 3-4 CTAs/SM: more chunks in different phases per SM hide the serial plan phase
    static constexpr int ID = 1;
    static constexpr int NT = 256;
    static constexpr int PPL = 5;
    static constexpr int WIN_MAX = NT * PPL * 16;       // 20480
    static constexpr int CHUNK = 16384;
    static constexpr int PRE = 16;
    static constexpr int OVERHANG = WIN_MAX - CHUNK - PRE;  // 4080
    static constexpr int MAXLINES = 1024;
    static constexpr int MAXREC = 256;
    static constexpr int STAGE = WIN_MAX;
    static constexpr int MIN_CTAS = 3;
};
inline int cfg_chunk_bytes(int cfg) { return cfg == CfgSmall::ID ? CfgSmall::CHUNK : CfgBig::CHUNK; }

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
    unsigned long long out_cursor;    // demux: bump allocator
    unsigned int flags;
    unsigned int n_events;
    unsigned int ticket;              // dynamic chunk counter
    unsigned int n_slow;              // records that took the brute-force match (diagnostic)
    unsigned long long phase_cycles[16];  // -DSK_PHASE_TIMING: per-phase SM cycles summed over chunks (thread 0)
};

// Exact-match index over the sample sheet (built on the host, sk_api.cu).  Samples are grouped in
// classes of identical care mask; per class a hash table maps the cared bytes of a barcode to the
// (first, last) sample holding exactly those bytes.  A read whose barcode is found here is at
// distance 0 from those samples and from no others, which settles fasta_demultiplex.rs:157-173
// without visiting the other samples; every other read takes the brute-force bit-plane path.
constexpr int FAST_NWMAX = 16;                  // key words per barcode (L <= 64)
constexpr int FAST_CLS_WORDS = 3 * FAST_NWMAX;  // care[16] | mulA[16] | mulB[16]
struct FastIdx {
    uint32_t n_classes;  // 0 = index unusable, every read takes the brute-force path
    uint32_t nw;         // words per key = ceil(L/4)
    uint32_t tsize;      // slots per class table (power of two)
    uint32_t pad;
    const uint32_t *cls;                 // n_classes * FAST_CLS_WORDS
    const unsigned long long *table;     // n_classes * tsize: tag(32) | first(16) | last(16); first == 0xFFFF: empty
    const uint32_t *skeys;               // S * nw: cared bytes of every sample
};

// Sample sheet in device memory (packed by the host, sk_api.cu).
struct SheetDev {
    const uint32_t *planes;  // S entries of {p0,p1,p2,care} (u32 x4) or, when wide, {p0,p1,p2,care} (u64 x4)
    const uint32_t *umask;   // S entries (u32) or 2*S (u64 as lo,hi): positions where the sheet has 'U'
    const uint8_t *lut;      // 256: bits 0-2 = 3-bit code (0 = matches no literal), bit 3 = [ACGTNacgtn+]
    uint32_t S, L, Umax, wide;
    FastIdx fast;
};

// Shared-memory carve-up (bytes), computed on the host and passed in KParams (the device code only
// reads it: recomputing it in the kernel costs ~5 % of all instructions).
struct SmemLayout {
    uint32_t win, stage, ls, rec, sheet, umask, lut, hist, sbase, ccount, fcls, ftab, slow, misc, total;
};

struct KParams {
    SmemLayout sl;
    // input stream
    const uint8_t *in;
    uint64_t n;
    uint32_t n_chunks;
    uint32_t lpr;         // lines per record: 4 (FASTQ) or 2 (FASTA)
    uint64_t rec_limit;   // process records with index < rec_limit
    uint32_t final_batch; // 1: end of buffer is end of file (EOF semantics); 0: trailing partial record is left
    uint32_t min_baseq;
    int32_t fused_trim;   // demux: >=0 -> trim by quality with this threshold
    uint32_t head_char;   // add barcode: '@' or '>' (uniform over the file)
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
    uint16_t *lens;         // [n_chunks * S]
    uint64_t *chunk_base;   // [n_chunks]
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
};

// Data outcome kinds (== SK_DATA_* in include/seqkit_b200.h); the low byte of DevStats::err_key.
enum : unsigned {
    K_BAD_HEADER = 1, K_LEN_MISMATCH = 2, K_SEQ_SHORT = 3, K_NO_BC = 4, K_BC_LEN = 5, K_INDEX_ASSERT = 6,
    K_BAD_FASTX_LINE = 7, K_NON_ASCII = 32, K_TOO_LONG = 33, K_TOO_DENSE = 34, K_MIXED = 35, K_OUT_OVERFLOW = 36,
    K_TRUNC_FUSED = 37,
};
enum : unsigned { F_MATE_COUNT = 1u, F_EVENTS_OVERFLOW = 2u, F_NON_ASCII = 0x100u };

constexpr int REC_BYTES = 26;  // per-record plan fields, see sk_kernels.cu
template <class Cfg>
inline __host__ __device__ SmemLayout smem_layout(uint32_t S, uint32_t wide, uint32_t n_classes, uint32_t tsize) {
    SmemLayout L;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 15u) & ~15u; return r; };
    L.win = take(Cfg::WIN_MAX);
    L.stage = take(Cfg::STAGE + 16);
    L.ls = take(Cfg::MAXLINES * 2);
    L.rec = take(Cfg::MAXREC * REC_BYTES);
    L.sheet = take(S * (wide ? 32u : 16u));
    L.umask = take(S * (wide ? 8u : 4u));
    L.lut = take(S ? 256 : 0);
    L.hist = take(S * 4);
    L.sbase = take(S * 4);
    L.ccount = take(S * 4);
    L.fcls = take(n_classes * FAST_CLS_WORDS * 4);
    L.ftab = take(n_classes * tsize * 8);
    L.slow = take(S ? Cfg::MAXREC * 2 : 0);
    L.misc = take(512);
    L.total = o;
    return L;
}

// Launchers (sk_kernels.cu)
int launch_chunk_kernel(int cfg, int op, const KParams &p, int sm_count, void *stream, const char **err);
int chunk_kernel_smem_bytes(int cfg, uint32_t S, uint32_t wide, uint32_t n_classes, uint32_t tsize);

}  // namespace sk
