// sk_lineops.cu -- the line engine: record-boundary scan by global line table, and the record-shuffling
// operators of SURVEY.md section 8(f) on top of it:
//
//   fasta trim --first/--last   fasta_trim.rs:24-47          LOP_TRIMFIX
//   fasta check                 fasta_check.rs:49-70         LOP_CHECK
//   fasta statistics            fasta_statistics.rs:13-52    LOP_STATS
//   fasta interleave            fasta_interleave.rs:14-35    LOP_INTERLEAVE
//   fasta deinterleave          fasta_deinterleave.rs:14-39  LOP_DEINTERLEAVE
//   fasta extract dual umi      fasta_extract_dual_umi.rs:14-72  LOP_DUALUMI
//
// Every operator is a handful of bandwidth-bound kernels over the whole stream:
//   nl_count / nl_bases / nl_fill   the north star's kernel (a): newline flags by SWAR, warp shuffles and a
//                                   block scan give every 16 KiB block its newline count; one CTA turns the
//                                   counts into block bases; the blocks write the line-start table.  Record
//                                   i is lines [lpr*i, lpr*i + lpr) exactly as the reference's read_line
//                                   calls see them (common.rs:106-112), whatever the record length.
//   plan      one thread per record: validity, output length, failure kind (the first failing record wins)
//   scan      exclusive prefix of the output lengths (len_sum / len_bases / len_apply)
//   emit      one warp per record: the record's pieces go to their place, 16 destination-aligned bytes per
//             lane and step (warp_copy_piece), literals by lane 0
// Framing is uniform per stream -- '@' files are 4 lines per record, '>' files 2 -- as decided by the stream's
// first byte; a record that starts with the other character is reported as SK_DATA_MIXED_FORMAT (the
// reference decides per record).  Streams are ASCII (F_NON_ASCII otherwise).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "sk_internal.h"

namespace sk {
extern __shared__ __align__(128) unsigned char sk_smem[];
}
#include "sk_device.cuh"
#include "sk_record.cuh"

namespace sk {

constexpr uint32_t NLB = 16384;  // bytes per block of the newline kernels (256 threads x 64 bytes)

// ---- line table ------------------------------------------------------------------------------------
// newline map of a 16-byte piece in natural order, exact for every byte value (eq_flags; nl_map_nat is the 7-bit form)
__device__ __forceinline__ uint32_t nl_map_exact(const uint4 v) {
    const uint32_t zx = eq_flags(v.x, 0x0A0A0A0Au), zy = eq_flags(v.y, 0x0A0A0A0Au);
    const uint32_t zz = eq_flags(v.z, 0x0A0A0A0Au), zw = eq_flags(v.w, 0x0A0A0A0Au);
    const uint32_t lo = __dp4a(zx, 0x08040201u, __dp4a(zy, 0x80402010u, 0u));
    const uint32_t hi = __dp4a(zz, 0x08040201u, __dp4a(zw, 0x80402010u, 0u));
    return (lo >> 7) + hi * 2u;
}
// Newline flags of the 64 bytes of a thread: bit k <=> byte k is '\n'.  Bytes at or past n read as 0.
__device__ __forceinline__ unsigned long long nl_bits64(const uint8_t *in, uint64_t pos, uint64_t n, uint32_t &hib) {
    unsigned long long m = 0;
    if (pos + 64 <= n) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 v = *(const uint4 *)(in + pos + 16 * q);
            hib |= v.x | v.y | v.z | v.w;
            m |= (unsigned long long)nl_map_exact(v) << (16 * q);
        }
    } else {
        for (uint32_t k = 0; k < 64 && pos + k < n; k++) {
            const uint8_t c = in[pos + k];
            hib |= c;
            if (c == '\n') m |= 1ull << k;
        }
    }
    return m;
}
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t *scratch, uint32_t &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) scratch[w] = x;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < NT / 32; k++) {
        const uint32_t t = scratch[k];
        if (k < w) before += t;
        all += t;
    }
    total = all;
    return before + x - v;
}
// `high`: set when the stream holds a byte >= 0x80.  strict: that refuses the batch (F_NON_ASCII); otherwise the
// operator's plan looks at the records one by one (UTF-8 in header and '+' lines is data like any other).
__global__ void __launch_bounds__(256) sk_nl_count_kernel(const uint8_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ blk, DevStats *st,
                                                          uint32_t *high, int strict) {
    __shared__ uint32_t scratch[8];
    const uint64_t pos = (uint64_t)blockIdx.x * NLB + 64ull * threadIdx.x;
    uint32_t hib = 0;
    const unsigned long long m = pos < n ? nl_bits64(in, pos, n, hib) : 0ull;
    if (hib & 0x80808080u) {
        if (strict) atomicOr(&st->flags, F_NON_ASCII);
        else *high = 1u;
    }
    uint32_t total;
    block_excl_scan_u32<256>((uint32_t)__popcll(m), scratch, total);
    if (threadIdx.x == 0) blk[blockIdx.x] = total;
}
// One CTA: blk[b] -> exclusive prefix; info = {n_lines, n_rec[lpr], ...}.  n_lines = newlines + 1 when the
// stream does not end with '\n' (the reference's last read_line returns the unterminated rest).
struct LineInfo {
    uint32_t n_lines, n_rec, overflow, high;  // high: a byte >= 0x80 somewhere in the stream (non-strict operators)
};
__global__ void __launch_bounds__(1024) sk_nl_bases_kernel(uint32_t *blk, uint32_t nb, const uint8_t *__restrict__ in, uint64_t n, uint32_t lpr,
                                                           uint32_t *starts, uint32_t cap, LineInfo *info) {
    __shared__ uint32_t scratch[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < nb; b0 += 1024u) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < nb ? blk[b] : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan_u32<1024>(v, scratch, total);
        if (b < nb) blk[b] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t newlines = carry;
        const uint32_t n_lines = newlines + ((n > 0 && in[n - 1] != '\n') ? 1u : 0u);
        info->n_lines = n_lines;
        info->n_rec = (n_lines + lpr - 1) / lpr;
        info->overflow = n_lines + 1u > cap ? 1u : 0u;
        if (cap) starts[0] = 0;
        if (n_lines < cap) starts[n_lines] = (uint32_t)n;  // sentinel: one past the last line
    }
}
__global__ void __launch_bounds__(256) sk_nl_fill_kernel(const uint8_t *__restrict__ in, uint64_t n, const uint32_t *__restrict__ blk,
                                                         uint32_t *__restrict__ starts, uint32_t cap) {
    __shared__ uint32_t scratch[8];
    const uint64_t pos = (uint64_t)blockIdx.x * NLB + 64ull * threadIdx.x;
    uint32_t hib = 0;
    unsigned long long m = pos < n ? nl_bits64(in, pos, n, hib) : 0ull;
    uint32_t total;
    uint32_t idx = blk[blockIdx.x] + block_excl_scan_u32<256>((uint32_t)__popcll(m), scratch, total);
    while (m) {  // newline number idx (0-based) at pos + k starts line idx + 1 at pos + k + 1
        const uint32_t k = (uint32_t)__ffsll((long long)m) - 1u;
        m &= m - 1;
        if (idx + 1u < cap) starts[idx + 1u] = (uint32_t)(pos + k + 1u);
        idx++;
    }
}

// ---- per-record views --------------------------------------------------------------------------------
struct LStream {
    const uint8_t *in;
    uint64_t n;
    const uint32_t *starts;
    const LineInfo *info;
};
struct LineRef {
    uint32_t s, len;
};
__device__ __forceinline__ LineRef line_of(const LStream &S, uint32_t k) {
    LineRef r;
    const uint32_t nl = S.info->n_lines;
    if (k >= nl) {  // past the end: read_line left the string empty
        r.s = (uint32_t)S.n;
        r.len = 0;
        return r;
    }
    r.s = S.starts[k];
    r.len = S.starts[k + 1] - r.s;
    return r;
}
__device__ __forceinline__ uint32_t trim_end_len_dev(const uint8_t *p, uint32_t len) {
    while (len && is_ws(p[len - 1])) len--;
    return len;
}

// str::trim_end (Unicode White_Space) of valid UTF-8: the new length
__device__ __forceinline__ uint32_t trim_end_unicode(const uint8_t *s, uint32_t n) {
    for (;;) {
        if (!n) return 0;
        const uint8_t c = s[n - 1];
        if (c < 0x80) {
            if (!is_ws(c)) return n;
            n--;
            continue;
        }
        if (n >= 2 && s[n - 2] == 0xC2 && (c == 0x85 || c == 0xA0)) {  // U+0085, U+00A0
            n -= 2;
            continue;
        }
        if (n >= 3) {
            const uint8_t a = s[n - 3], b = s[n - 2];
            const bool ws = (a == 0xE1 && b == 0x9A && c == 0x80) ||                                                   // U+1680
                            (a == 0xE2 && b == 0x80 && ((c >= 0x80 && c <= 0x8A) || c == 0xA8 || c == 0xA9 || c == 0xAF)) ||  // U+2000-200A, 2028, 2029, 202F
                            (a == 0xE2 && b == 0x81 && c == 0x9F) ||                                                   // U+205F
                            (a == 0xE3 && b == 0x80 && c == 0x80);                                                     // U+3000
            if (ws) {
                n -= 3;
                continue;
            }
        }
        return n;
    }
}

struct LParams {
    int op;
    LStream a, b;          // b: second input of interleave
    uint32_t lpr;          // lines per record of stream a: 4 ('@') or 2 ('>')
    uint32_t head;         // '@' or '>'
    uint32_t x, y;         // trim: first, last; dual umi: first_bases; deinterleave: parity of the pass
    uint64_t rec_limit;
    uint64_t max_records;  // entries of the per-record arrays below (sk_limits.max_records)
    uint32_t *out_len;     // [records]
    uint64_t *dst;         // [records] exclusive prefix of out_len
    uint8_t *out;
    uint64_t out_cap;
    DevStats *st;
    // statistics
    unsigned long long *h_keys, *h_rep;
    uint32_t *h_cnt;
    uint32_t h_mask;
    uint32_t *bc_ref;      // [records] (off << 8 | len) of the record's barcode, 0xFFFFFFFF = none
    // add barcode: record table of the barcode stream (b.in holds its bytes)
    const RecRef *bc_tab;
    const DevStats *bc_stats;
};

enum : int { LOP_TRIMFIX = 0, LOP_CHECK = 1, LOP_STATS = 2, LOP_INTERLEAVE = 3, LOP_DEINTERLEAVE = 4, LOP_DUALUMI = 5,
             // trim / mask by quality on the line engine: the last resort behind the warp and the general engine, for
             // records of any length and density and for UTF-8 in header and '+' lines (sk_api.cu: sk_wait)
             LOP_TRIMQ = 6, LOP_MASKQ = 7,
             // add barcode on the line engine (fasta_add_barcode.rs:19-44): the same last resort, reads of any length
             LOP_ADDBC = 8 };
// further data outcome kinds of the line operators (== SK_DATA_* in the header)
enum : unsigned { K_NO_PLUS = 8, K_INCONSISTENT = 9, K_QUAL_SHORT = 10, K_HASH_COLLISION = 38 };

// number of output units of the operator: records, or record pairs
__device__ __forceinline__ uint32_t units_of(const LParams &p) {
    const uint32_t nr = p.a.info->n_rec;
    uint32_t u = (p.op == LOP_DEINTERLEAVE || p.op == LOP_DUALUMI) ? (nr + 1u) / 2u : nr;
    if ((uint64_t)u > p.rec_limit) u = (uint32_t)p.rec_limit;
    return u;
}
// more lines than the line table holds, or more records than the per-record arrays: the batch is refused (K_TOO_MANY)
__device__ __forceinline__ bool line_refused(const LParams &p) {
    return p.a.info->overflow || (p.op == LOP_INTERLEAVE && p.b.info->overflow) || (uint64_t)units_of(p) > p.max_records;
}
// header check of a record of stream a: 0 fine, else failure kind
__device__ __forceinline__ unsigned head_kind(const LStream &S, LineRef h, uint32_t head) {
    const uint32_t c = h.len ? S.in[h.s] : 0u;
    if (c == head) return 0;
    return (c == '@' || c == '>') ? K_MIXED : K_BAD_HEADER;
}
// bytes of record r of a stream (its lpr lines are contiguous)
__device__ __forceinline__ LineRef record_of(const LStream &S, uint32_t r, uint32_t lpr) {
    const LineRef a = line_of(S, r * lpr), e = line_of(S, r * lpr + lpr);
    LineRef o;
    o.s = a.s;
    o.len = e.s - a.s;
    return o;
}

// FNV-1a over the barcode bytes; bit 0 forced so that 0 marks an empty slot
__device__ __forceinline__ unsigned long long fnv64(const uint8_t *p, uint32_t len) {
    unsigned long long h = 1469598103934665603ull;
    for (uint32_t i = 0; i < len; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h | 1ull;
}
__device__ __forceinline__ bool stat_class(uint8_t c) {  // [ACGTNacgtn] (fasta_statistics.rs:17: no '+')
    switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'N': case 'a': case 'c': case 'g': case 't': case 'n': return true;
    }
    return false;
}

// str::from_utf8 acceptance of a line (what BufRead::read_line enforces, common.rs:106-112)
__device__ __forceinline__ bool utf8_ok(const uint8_t *s, uint32_t n) {
    uint32_t i = 0;
    while (i < n) {
        const uint8_t c = s[i];
        if (c < 0x80) {
            i++;
        } else if (c >= 0xC2 && c <= 0xDF) {
            if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return false;
            i += 2;
        } else if (c >= 0xE0 && c <= 0xEF) {
            if (i + 2 >= n) return false;
            const uint8_t c1 = s[i + 1], c2 = s[i + 2];
            if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80) return false;
            if (c == 0xE0 && c1 < 0xA0) return false;  // overlong
            if (c == 0xED && c1 >= 0xA0) return false;  // surrogates
            i += 3;
        } else if (c >= 0xF0 && c <= 0xF4) {
            if (i + 3 >= n) return false;
            const uint8_t c1 = s[i + 1], c2 = s[i + 2], c3 = s[i + 3];
            if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80 || (c3 & 0xC0) != 0x80) return false;
            if (c == 0xF0 && c1 < 0x90) return false;
            if (c == 0xF4 && c1 >= 0x90) return false;
            i += 4;
        } else {
            return false;
        }
    }
    return true;
}

// ---- plan --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sk_line_plan_kernel(const LParams p) {
    if (line_refused(p)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) report_err(p.st, p.max_records, K_TOO_MANY);
        return;
    }
    const uint32_t nu = units_of(p);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nu; i += gridDim.x * blockDim.x) {
        uint32_t olen = 0;
        unsigned kind = 0;
        const uint8_t *in = p.a.in;
        if (p.op == LOP_TRIMFIX) {
            const LineRef h = line_of(p.a, i * p.lpr);
            kind = head_kind(p.a, h, p.head);
            if (!kind) {
                const LineRef sq = line_of(p.a, i * p.lpr + 1);
                const uint32_t sl = trim_end_len_dev(in + sq.s, sq.len);              // seq.trim_end().len()  (:31)
                const bool cut = (uint64_t)p.x + p.y < sl;                            // :32
                const uint32_t L = cut ? sl - p.x - p.y : 0u;
                olen = h.len + L + 1u;                                                // :33 / :35
                if (p.lpr == 4) {
                    const LineRef ql = line_of(p.a, i * p.lpr + 3);
                    if (cut && ql.len < sl - p.y) kind = K_QUAL_SHORT;                // &qual[..] out of range (:41): panic
                    olen += 2u + L + 1u;                                              // "+\n" qual "\n"  (:41 / :43)
                }
            }
        } else if (p.op == LOP_CHECK) {
            const LineRef h = line_of(p.a, i * p.lpr);
            kind = head_kind(p.a, h, p.head);
            if (!kind && p.lpr == 4) {
                const LineRef pl = line_of(p.a, i * p.lpr + 2);
                if (!(pl.len && in[pl.s] == '+')) kind = K_NO_PLUS;                   // fasta_check.rs:58-61
            }
        } else if (p.op == LOP_STATS) {
            const LineRef h = line_of(p.a, i * p.lpr);
            uint32_t ref = 0xFFFFFFFFu;
            // leftmost " BC:" followed by at least one class byte; the run extends while in class (:17,:25-28)
            for (uint32_t k = 0; k + 5 <= h.len; k++) {
                const uint8_t *q = in + h.s + k;
                if (q[0] == ' ' && q[1] == 'B' && q[2] == 'C' && q[3] == ':' && stat_class(q[4])) {
                    uint32_t e = k + 5;
                    while (e < h.len && stat_class(in[h.s + e])) e++;
                    const uint32_t off = h.s + k + 4, len = e - k - 4;
                    // one slot per distinct hash; count and the earliest occurrence (the representative) by atomics
                    const unsigned long long key = fnv64(in + off, len);
                    uint32_t slot = (uint32_t)(key >> 17) & p.h_mask;
                    for (;;) {
                        const unsigned long long old = atomicCAS(&p.h_keys[slot], 0ull, key);
                        if (old == 0ull || old == key) break;
                        slot = (slot + 1u) & p.h_mask;
                    }
                    atomicAdd(&p.h_cnt[slot], 1u);
                    atomicMin(&p.h_rep[slot], ((unsigned long long)off << 32) | len);
                    ref = slot;
                    break;
                }
            }
            p.bc_ref[i] = ref;
            kind = head_kind(p.a, h, p.head);  // the header is validated after the search (:31-37)
        } else if (p.op == LOP_INTERLEAVE) {
            const LineRef h = line_of(p.a, i * p.lpr);
            kind = head_kind(p.a, h, p.head);
            if (!kind) {
                const LineRef h2 = line_of(p.b, i * p.lpr);
                const uint32_t c2 = h2.len ? p.b.in[h2.s] : 0u;
                if (c2 != p.head) kind = K_INCONSISTENT;                              // fasta_interleave.rs:26-29
                olen = record_of(p.a, i, p.lpr).len + record_of(p.b, i, p.lpr).len;
            }
        } else if (p.op == LOP_DEINTERLEAVE) {
            const LineRef h = line_of(p.a, 2u * i * p.lpr);
            kind = head_kind(p.a, h, p.head);
            if (!kind) {
                const LineRef h2 = line_of(p.a, (2u * i + 1u) * p.lpr);
                const uint32_t c2 = h2.len ? in[h2.s] : 0u;
                if (c2 != p.head) kind = K_INCONSISTENT;                              // fasta_deinterleave.rs:30-33
                olen = record_of(p.a, 2u * i + p.x, p.lpr).len;                        // pass x = 0: mate 1, 1: mate 2
            }
        } else if (p.op == LOP_TRIMQ || p.op == LOP_MASKQ) {
            const LineRef h = line_of(p.a, i * 4u), sq = line_of(p.a, i * 4u + 1u), pl = line_of(p.a, i * 4u + 2u), ql = line_of(p.a, i * 4u + 3u);
            if (!(h.len && in[h.s] == '@')) kind = K_BAD_HEADER;  // fasta_trim_by_quality.rs:20-22, fasta_mask_by_quality.rs:21-23
            if (!kind && p.a.info->high) {
                // bytes >= 0x80: data like any other in the header and the '+' line when they are valid UTF-8 (read_line,
                // common.rs:106-112); in the bases or qualities (char-wise zip, Unicode trim_end) the batch is refused
                bool bad = !utf8_ok(in + h.s, h.len) || !utf8_ok(in + pl.s, pl.len);
                for (uint32_t t = 0; t < sq.len; t++) bad = bad || in[sq.s + t] >= 0x80;
                for (uint32_t t = 0; t < ql.len; t++) bad = bad || in[ql.s + t] >= 0x80;
                if (bad) kind = K_NON_ASCII;
            }
            if (!kind && p.op == LOP_TRIMQ) {  // fasta_trim_by_quality.rs:28-48
                const int minq = (int)p.x;
                uint32_t k = trim_end_len_dev(in + ql.s, ql.len);  // :31
                int total = -50, lowest = -50;                     // :28-29
                uint32_t lowest_k = k;
                while (k > 0) {                                    // :33-42
                    k--;
                    total += (int)(uint8_t)(in[ql.s + k] - 33u) - minq;  // wrapping u8 subtraction (:35)
                    if (total > 0) break;
                    if (total < lowest) {
                        lowest = total;
                        lowest_k = k;
                    }
                }
                p.bc_ref[i] = lowest_k;
                if (lowest_k == 0) olen = h.len + 6u;              // "N\n+\n!\n"  (:44-45)
                else if (lowest_k > sq.len) kind = K_SEQ_SHORT;    // &seq[..k] panics (:47)
                else olen = h.len + 2u * lowest_k + 4u;            // :47
            } else if (!kind) {  // fasta_mask_by_quality.rs:32-45
                const uint32_t sl = sq.len - ((sq.len && in[sq.s + sq.len - 1] == '\n') ? 1u : 0u);  // :32
                const uint32_t qn = ql.len - ((ql.len && in[ql.s + ql.len - 1] == '\n') ? 1u : 0u);  // :33
                if (sl != qn) kind = K_LEN_MISMATCH;                                                  // :35-37
                else olen = h.len + 2u * sl + 4u;                                                     // :26,:44
            }
        } else if (p.op == LOP_ADDBC) {  // fasta_add_barcode.rs:29-43
            const LineRef h = line_of(p.a, i * p.lpr);
            const LineRef whole = record_of(p.a, i, p.lpr);
            const uint32_t c = h.len ? in[h.s] : 0u;
            if (c != p.head) kind = (c == '@' || c == '>') ? K_MIXED : K_BAD_FASTX_LINE;
            // every line is data like any other as long as it is valid UTF-8 (read_line, common.rs:106-112)
            if (!kind && p.a.info->high && !utf8_ok(in + whole.s, whole.len)) kind = K_NON_ASCII;
            if (!kind) {
                uint32_t bl = 0;
                const unsigned long long nb = p.bc_stats->n_records;
                if (nb) {  // barcode of iteration i; the last one is reused once the barcode file is exhausted (:20-27)
                    const RecRef rr = p.bc_tab[(unsigned long long)i < nb ? i : nb - 1ull];  // (i < units <= max_records = the table's size)
                    bl = rr.seq_len;
                    if (rr.flags & RR_LONG) kind = K_TOO_LONG;
                }
                olen = trim_end_unicode(in + h.s, h.len) + 4u + bl + 1u + (whole.len - h.len);  // :33 + the other lines
            }
        } else {  // LOP_DUALUMI (fasta_extract_dual_umi.rs:27-70)
            const uint32_t r1 = 2u * i, r2 = 2u * i + 1u, N = p.x;
            const LineRef h1 = line_of(p.a, r1 * p.lpr);
            kind = head_kind(p.a, h1, p.head);
            if (!kind) {
                const LineRef h2 = line_of(p.a, r2 * p.lpr);
                const uint32_t c2 = h2.len ? in[h2.s] : 0u;
                const LineRef s1 = line_of(p.a, r1 * p.lpr + 1), s2 = line_of(p.a, r2 * p.lpr + 1);
                if (c2 != p.head) {
                    kind = K_INCONSISTENT;                                            // :42-44 / :50-52
                } else if (N > s1.len || N > s2.len) {
                    kind = K_SEQ_SHORT;                                               // &seq_1[0..N] out of range: panic (:55,:57)
                } else {
                    const uint32_t t1 = trim_end_len_dev(in + h1.s, h1.len), t2 = trim_end_len_dev(in + h2.s, h2.len);
                    const uint32_t umi = 2u * N + 1u;
                    olen = t1 + 4u + umi + 1u + (s1.len - N) + t2 + 4u + umi + 1u + (s2.len - N);
                    if (p.lpr == 4) {
                        const LineRef q1 = line_of(p.a, r1 * p.lpr + 3), q2 = line_of(p.a, r2 * p.lpr + 3);
                        if (N > q1.len || N > q2.len) kind = K_QUAL_SHORT;            // &qual_1[N..] out of range: panic (:62,:64)
                        olen += 2u + (q1.len - N) + 2u + (q2.len - N);
                    }
                }
            }
        }
        if (kind) {
            report_err(p.st, i, kind);
            olen = 0;
        }
        if (p.out_len) p.out_len[i] = olen;
    }
}

// statistics, second pass: every record's bytes against the representative of its slot (a 64-bit collision
// between different barcodes is reported, never merged)
__global__ void __launch_bounds__(256) sk_stats_verify_kernel(const LParams p) {
    if (line_refused(p)) return;
    const uint32_t nu = units_of(p);
    const uint8_t *in = p.a.in;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nu; i += gridDim.x * blockDim.x) {
        const uint32_t slot = p.bc_ref[i];
        if (slot == 0xFFFFFFFFu) continue;
        const LineRef h = line_of(p.a, i * p.lpr);
        // locate the barcode again (cheap) and compare with the representative
        for (uint32_t k = 0; k + 5 <= h.len; k++) {
            const uint8_t *q = in + h.s + k;
            if (q[0] == ' ' && q[1] == 'B' && q[2] == 'C' && q[3] == ':' && stat_class(q[4])) {
                uint32_t e = k + 5;
                while (e < h.len && stat_class(in[h.s + e])) e++;
                const uint32_t off = h.s + k + 4, len = e - k - 4;
                const unsigned long long rep = p.h_rep[slot];
                const uint32_t roff = (uint32_t)(rep >> 32), rlen = (uint32_t)rep;
                bool same = rlen == len;
                for (uint32_t t = 0; same && t < len; t++) same = in[off + t] == in[roff + t];
                if (!same) report_err(p.st, i, K_HASH_COLLISION);
                break;
            }
        }
    }
}
// statistics, third pass: the table's entries as a dense list {rep, count}
__global__ void __launch_bounds__(256) sk_stats_list_kernel(const unsigned long long *keys, const unsigned long long *rep, const uint32_t *cnt,
                                                            uint32_t cap, unsigned long long *list, uint32_t *n_list) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x)
        if (keys[s]) {
            const uint32_t k = atomicAdd(n_list, 1u);
            list[2 * k] = rep[s];
            list[2 * k + 1] = cnt[s];
        }
}

// ---- finish: records before the first failure are the output -------------------------------------------
__global__ void sk_line_finish_kernel(const LParams p, int which_out) {
    const uint32_t nu = line_refused(p) ? 0u : units_of(p);
    uint32_t lim = nu;
    if (p.st->err_key) {
        const unsigned long long k = ~p.st->err_key;
        const unsigned long long rec = k >> 8;
        if (rec < lim) lim = (uint32_t)rec;
    }
    unsigned long long bytes = 0;
    if (lim && p.out_len) bytes = p.dst[lim - 1] + p.out_len[lim - 1];
    p.st->n_records = lim;
    p.st->n_lines = p.a.info->n_lines;
    if (which_out == 0) {
        p.st->out_bytes = bytes;
        p.st->out_extent = bytes;
    } else {
        p.st->compact_extent = bytes;  // second output stream (deinterleave): length parked here
    }
    // bytes of stream a covered by the processed records
    const uint32_t per = (p.op == LOP_DEINTERLEAVE || p.op == LOP_DUALUMI) ? 2u * p.lpr : p.lpr;
    p.st->consumed = line_of(p.a, lim * per).s;
    if (bytes > p.out_cap) report_err(p.st, 0, K_OUT_OVERFLOW);
}

// ---- emit --------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_lit(uint8_t *dst, unsigned long long at, const char *lit, uint32_t n, int lane) {
    if ((uint32_t)lane < n) dst[at + lane] = (uint8_t)lit[lane];
}
__global__ void __launch_bounds__(256) sk_line_emit_kernel(const LParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t wpb = blockDim.x >> 5, gw = blockIdx.x * wpb + (threadIdx.x >> 5), nw = gridDim.x * wpb;
    if (line_refused(p)) return;
    uint32_t lim = units_of(p);
    if (p.st->err_key) {
        const unsigned long long rec = (~p.st->err_key) >> 8;
        if (rec < lim) lim = (uint32_t)rec;
    }
    if (lim && p.dst[lim - 1] + p.out_len[lim - 1] > p.out_cap) return;  // K_OUT_OVERFLOW (finish kernel)
    const uint8_t *in = p.a.in;
    uint8_t *out = p.out;
    for (uint32_t i = gw; i < lim; i += nw) {
        unsigned long long d = p.dst[i];
        if (p.op == LOP_TRIMFIX) {
            const LineRef h = line_of(p.a, i * p.lpr), sq = line_of(p.a, i * p.lpr + 1);
            const uint32_t sl = trim_end_len_dev(in + sq.s, sq.len);
            const bool cut = (uint64_t)p.x + p.y < sl;
            const uint32_t L = cut ? sl - p.x - p.y : 0u;
            warp_copy_piece(in, out, h.s, d, h.len, lane);
            d += h.len;
            if (L) warp_copy_piece(in, out, sq.s + p.x, d, L, lane);
            d += L;
            if (p.lpr == 4) {
                const LineRef ql = line_of(p.a, i * p.lpr + 3);
                put_lit(out, d, "\n+\n", 3, lane);
                d += 3;
                if (L) warp_copy_piece(in, out, ql.s + p.x, d, L, lane);
                d += L;
            }
            put_lit(out, d, "\n", 1, lane);
        } else if (p.op == LOP_TRIMQ) {
            const LineRef h = line_of(p.a, i * 4u), sq = line_of(p.a, i * 4u + 1u), ql = line_of(p.a, i * 4u + 3u);
            const uint32_t kk = p.bc_ref[i];
            warp_copy_piece(in, out, h.s, d, h.len, lane);  // header verbatim (:23)
            d += h.len;
            if (kk == 0) {
                put_lit(out, d, "N\n+\n!\n", 6, lane);    // :44-45
            } else {                                         // :47
                warp_copy_piece(in, out, sq.s, d, kk, lane);
                put_lit(out, d + kk, "\n+\n", 3, lane);
                warp_copy_piece(in, out, ql.s, d + kk + 3u, kk, lane);
                put_lit(out, d + 2ull * kk + 3u, "\n", 1, lane);
            }
        } else if (p.op == LOP_MASKQ) {
            const LineRef h = line_of(p.a, i * 4u), sq = line_of(p.a, i * 4u + 1u), ql = line_of(p.a, i * 4u + 3u);
            const uint32_t sl = sq.len - ((sq.len && in[sq.s + sq.len - 1] == '\n') ? 1u : 0u);
            warp_copy_piece(in, out, h.s, d, h.len, lane);  // :26
            d += h.len;
            for (uint32_t t = (uint32_t)lane; t < sl; t += 32u)  // :40-43 (wrapping u8: bytes below '!' are never masked)
                out[d + t] = (uint8_t)(in[ql.s + t] - 33u) < (uint8_t)p.x ? (uint8_t)'N' : in[sq.s + t];
            put_lit(out, d + sl, "\n+\n", 3, lane);       // :44
            if (sl) warp_copy_piece(in, out, ql.s, d + sl + 3u, sl, lane);
            put_lit(out, d + 2ull * sl + 3u, "\n", 1, lane);
        } else if (p.op == LOP_ADDBC) {
            const LineRef h = line_of(p.a, i * p.lpr), whole = record_of(p.a, i, p.lpr);
            const uint32_t alen = trim_end_unicode(in + h.s, h.len);
            if (alen) warp_copy_piece(in, out, h.s, d, alen, lane);
            d += alen;
            put_lit(out, d, " BC:", 4, lane);
            d += 4;
            const unsigned long long nb = p.bc_stats->n_records;
            if (nb) {
                const RecRef rr = p.bc_tab[(unsigned long long)i < nb ? i : nb - 1ull];
                if (rr.seq_len) warp_copy_piece(p.b.in, out, rr.seq_off, d, rr.seq_len, lane);
                d += rr.seq_len;
            }
            put_lit(out, d, "\n", 1, lane);
            d += 1;
            if (whole.len > h.len) warp_copy_piece(in, out, (unsigned long long)h.s + h.len, d, whole.len - h.len, lane);
        } else if (p.op == LOP_INTERLEAVE) {
            const LineRef r1 = record_of(p.a, i, p.lpr), r2 = record_of(p.b, i, p.lpr);
            if (r1.len) warp_copy_piece(in, out, r1.s, d, r1.len, lane);
            if (r2.len) warp_copy_piece(p.b.in, out, r2.s, d + r1.len, r2.len, lane);
        } else if (p.op == LOP_DEINTERLEAVE) {
            const LineRef r = record_of(p.a, 2u * i + p.x, p.lpr);
            if (r.len) warp_copy_piece(in, out, r.s, d, r.len, lane);
        } else if (p.op == LOP_DUALUMI) {
            const uint32_t N = p.x;
            for (uint32_t m = 0; m < 2; m++) {
                const uint32_t r = 2u * i + m;
                const LineRef h = line_of(p.a, r * p.lpr), sq = line_of(p.a, r * p.lpr + 1);
                const LineRef s1 = line_of(p.a, 2u * i * p.lpr + 1), s2 = line_of(p.a, (2u * i + 1u) * p.lpr + 1);
                const uint32_t t = trim_end_len_dev(in + h.s, h.len);
                if (t) warp_copy_piece(in, out, h.s, d, t, lane);               // header.trim_end()
                d += t;
                put_lit(out, d, " RX:", 4, lane);
                d += 4;
                if (N) warp_copy_piece(in, out, s1.s, d, N, lane);              // umi = seq_1[0..N] "+" seq_2[0..N]
                d += N;
                put_lit(out, d, "+", 1, lane);
                d += 1;
                if (N) warp_copy_piece(in, out, s2.s, d, N, lane);
                d += N;
                put_lit(out, d, "\n", 1, lane);
                d += 1;
                if (sq.len > N) warp_copy_piece(in, out, sq.s + N, d, sq.len - N, lane);   // &seq[N..] with its newline
                d += sq.len - N;
                if (p.lpr == 4) {
                    const LineRef ql = line_of(p.a, r * p.lpr + 3);
                    put_lit(out, d, "+\n", 2, lane);
                    d += 2;
                    if (ql.len > N) warp_copy_piece(in, out, ql.s + N, d, ql.len - N, lane);
                    d += ql.len - N;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side of the launch sequences (called from sk_api.cu)
// ------------------------------------------------------------------------------------------------
// dst[i] = sum of len[0..i): sums of blocks of 1024, one CTA over the block sums, then every block adds its base
__global__ void __launch_bounds__(1024) sk_len_sum_kernel(const uint32_t *__restrict__ len, unsigned long long *__restrict__ bsum, uint32_t n) {
    __shared__ uint32_t scratch[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    uint32_t total;
    block_excl_scan_u32<1024>(i < n ? len[i] : 0u, scratch, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) sk_len_bases_kernel(unsigned long long *bsum, uint32_t nb) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < nb; b0 += 1024u) {
        const uint32_t b = b0 + threadIdx.x;
        const unsigned long long v = b < nb ? bsum[b] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += y;
        }
        if (lane == 31) ws[w] = x;
        __syncthreads();
        unsigned long long before = carry;
        for (uint32_t k = 0; k < w; k++) before += ws[k];
        if (b < nb) bsum[b] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + x;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) sk_len_apply_kernel(const uint32_t *__restrict__ len, const unsigned long long *__restrict__ bsum,
                                                            uint64_t *__restrict__ dst, uint32_t n) {
    __shared__ uint32_t scratch[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    uint32_t total;
    const uint32_t ex = block_excl_scan_u32<1024>(i < n ? len[i] : 0u, scratch, total);
    if (i < n) dst[i] = bsum[blockIdx.x] + ex;
}
static int launch_len_scan(const uint32_t *len, uint64_t *dst, uint32_t n, unsigned long long *bsum, cudaStream_t stream) {
    if (!n) return 0;
    const uint32_t nb = (n + 1023u) / 1024u;
    sk_len_sum_kernel<<<nb, 1024, 0, stream>>>(len, bsum, n);
    sk_len_bases_kernel<<<1, 1024, 0, stream>>>(bsum, nb);
    sk_len_apply_kernel<<<nb, 1024, 0, stream>>>(len, bsum, dst, n);
    return 3;
}

struct LineWork {
    uint32_t *blk[2], *starts[2];
    LineInfo *info[2];
    uint32_t *out_len, *bc_ref;
    uint64_t *dst;
    unsigned long long *bsum;
    uint32_t cap_lines;
};
static LineWork carve(void *work, uint64_t max_stream_bytes, uint64_t max_records) {
    LineWork w;
    uint8_t *q = (uint8_t *)work;
    auto take = [&](uint64_t bytes) {
        uint8_t *r = q;
        q += (bytes + 255) & ~255ull;
        return r;
    };
    const uint64_t nb = (max_stream_bytes + NLB - 1) / NLB + 2;
    w.cap_lines = (uint32_t)std::min<uint64_t>(max_records * 4 + 64, 0xFFFFFFF0ull);
    for (int k = 0; k < 2; k++) {
        w.blk[k] = (uint32_t *)take(nb * 4);
        w.starts[k] = (uint32_t *)take((uint64_t)w.cap_lines * 4);
        w.info[k] = (LineInfo *)take(64);
    }
    w.out_len = (uint32_t *)take(max_records * 4);
    w.dst = (uint64_t *)take(max_records * 8);
    w.bc_ref = (uint32_t *)take(max_records * 4);
    w.bsum = (unsigned long long *)take((max_records / 1024 + 2) * 8);
    return w;
}
uint64_t lineops_work_bytes(uint64_t max_stream_bytes, uint64_t max_records) {
    const uint64_t nb = (max_stream_bytes + NLB - 1) / NLB + 2;
    const uint64_t cap_lines = std::min<uint64_t>(max_records * 4 + 64, 0xFFFFFFF0ull);
    auto r = [](uint64_t b) { return (b + 255) & ~255ull; };
    return 2 * (r(nb * 4) + r(cap_lines * 4) + r(64)) + r(max_records * 4) + r(max_records * 8) + r(max_records * 4) +
           r((max_records / 1024 + 2) * 8) + 256;
}

static int index_lines(const uint8_t *in, uint64_t n, uint32_t lpr, const LineWork &w, int k, DevStats *st, int strict, cudaStream_t stream) {
    const uint32_t nb = (uint32_t)((n + NLB - 1) / NLB);
    cudaMemsetAsync(&w.info[k]->high, 0, 4, stream);
    if (nb) sk_nl_count_kernel<<<nb, 256, 0, stream>>>(in, n, w.blk[k], st, &w.info[k]->high, strict);
    sk_nl_bases_kernel<<<1, 1024, 0, stream>>>(w.blk[k], nb, in, n, lpr, w.starts[k], w.cap_lines, w.info[k]);
    if (nb) sk_nl_fill_kernel<<<nb, 256, 0, stream>>>(in, n, w.blk[k], w.starts[k], w.cap_lines);
    return nb ? 3 : 1;
}

// ---- demultiplex on the line engine ------------------------------------------------------------------
// Header-route demultiplex (fasta_demultiplex.rs:117-249) for the batches neither chunk engine takes: records longer
// than the general engine's overhang, denser than its record slots, UTF-8 in header and '+' lines.  One thread per record
// plans (validate, leftmost " BC:x", length, distance to every sample over the sheet's bit planes, decide, UMI, header
// pieces, fused quality trim, output length), an exclusive scan places the records, one warp per record writes its pieces.
// The tables the other engines leave -- assign[], umi[], counts[], events, one (sample, len) group per record and one
// slice-table row per 32 records -- are filled in the same form, so the compaction and the host code do not know the
// difference.  A failing record is reported (lowest record first) and the host replays the batch up to it, as with the
// other engines.  A record's output is limited to 64 KiB - 1 by the group table (SK_DATA_RECORD_TOO_LONG beyond).
struct DParams {
    LStream a;             // the mate's stream
    uint32_t mate;         // 0: mate 1 (extract, match, decide, emit), 1: mate 2 (emit)
    uint64_t rec_limit;
    uint64_t max_records;  // entries of assign[], groups[], out_len[] ... (sk_limits.max_records)
    int32_t fused_trim;    // >= 0: trim by quality with this threshold first
    SheetDev sheet;
    int16_t *assign;
    uint8_t *umi;
    Group *groups;
    ChunkRow *rows;
    uint32_t max_rows;
    unsigned long long *counts;
    Event *events;
    uint32_t events_cap;
    const DevStats *r1_stats;
    uint32_t n_index;        // --index1 / --index2: the barcode is the index reads' sequence lines (:126-136), the header stays whole
    const RecRef *ext_tab[2];
    const uint8_t *ext_data[2];
    const DevStats *ext_stats[2];
    uint32_t *out_len, *kk;  // [records]; kk: bases kept by the fused trim (0 = the "N" record), 0xFFFFFFFF = three lines verbatim
    uint64_t *dst;
    uint8_t *out;            // nullptr: dry run
    uint64_t out_cap;
    DevStats *st;
};
__device__ __forceinline__ uint32_t dm_units(const DParams &p) {
    uint32_t u = p.a.info->n_rec;
    if ((uint64_t)u > p.rec_limit) u = (uint32_t)p.rec_limit;
    return u;
}
// leftmost " BC:" followed by at least one class byte, greedy class run (fasta_demultiplex.rs:38,138-141)
__device__ __forceinline__ bool dm_find_bc(const uint8_t *h, uint32_t n, const uint8_t *lut, uint32_t &st, uint32_t &en) {
    for (uint32_t k = 0; k + 5 <= n; k++) {
        if (h[k] == ' ' && h[k + 1] == 'B' && h[k + 2] == 'C' && h[k + 3] == ':' && (lut[h[k + 4]] & 8u)) {
            uint32_t e = k + 5;
            while (e < n && (lut[h[e]] & 8u)) e++;
            st = k;
            en = e;
            return true;
        }
    }
    return false;
}
// header.drain(st..en) then trim_end (:145,:206 / :219-229): the header is its first alen bytes and blen bytes from en
__device__ __forceinline__ void dm_pieces(const uint8_t *h, uint32_t n, bool cut, uint32_t st, uint32_t en, uint32_t &alen, uint32_t &blen) {
    if (!cut) {
        alen = trim_end_unicode(h, n);
        blen = 0;
        return;
    }
    blen = trim_end_unicode(h + en, n - en);
    alen = blen ? st : trim_end_unicode(h, st);
}
// fasta_trim_by_quality.rs:28-48 on one record: bases kept (0: the "N" record); false: &seq[..k] would panic (:47)
__device__ __forceinline__ bool dm_trimq(const uint8_t *in, LineRef sq, LineRef ql, int minq, uint32_t &kept) {
    uint32_t k = trim_end_len_dev(in + ql.s, ql.len);
    int total = -50, lowest = -50;
    uint32_t lowest_k = k;
    while (k > 0) {
        k--;
        total += (int)(uint8_t)(in[ql.s + k] - 33u) - minq;
        if (total > 0) break;
        if (total < lowest) {
            lowest = total;
            lowest_k = k;
        }
    }
    kept = lowest_k;
    return lowest_k <= sq.len;
}
__device__ __forceinline__ bool dm_utf8_fine(const LStream &S, LineRef h, LineRef sq, LineRef pl, LineRef ql) {
    const uint8_t *in = S.in;
    bool bad = !utf8_ok(in + h.s, h.len) || !utf8_ok(in + pl.s, pl.len);
    for (uint32_t t = 0; t < sq.len; t++) bad = bad || in[sq.s + t] >= 0x80;
    for (uint32_t t = 0; t < ql.len; t++) bad = bad || in[ql.s + t] >= 0x80;
    return !bad;
}

__global__ void __launch_bounds__(256) sk_dm_plan_kernel(const DParams p) {
    if (p.a.info->overflow || (uint64_t)dm_units(p) > p.max_records) {
        if (blockIdx.x == 0 && threadIdx.x == 0) report_err(p.st, p.max_records, K_TOO_MANY);
        return;
    }
    const uint32_t nu = dm_units(p);
    if ((nu + 31u) / 32u > p.max_rows) {
        if (blockIdx.x == 0 && threadIdx.x == 0) report_err(p.st, 0, K_TOO_DENSE);
        return;
    }
    const uint8_t *in = p.a.in;
    const uint8_t *lut = p.sheet.lut;
    const uint32_t S = p.sheet.S, Lb = p.sheet.L;
    const bool fused = p.fused_trim >= 0;
    const unsigned long long n1 = p.mate ? p.r1_stats->n_records : 0ull;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nu; i += gridDim.x * blockDim.x) {
        const LineRef h = line_of(p.a, i * 4u), sq = line_of(p.a, i * 4u + 1u), pl = line_of(p.a, i * 4u + 2u), ql = line_of(p.a, i * 4u + 3u);
        unsigned kind = 0;
        uint32_t olen = 0, kept = 0xFFFFFFFFu;
        int sample = -1;
        const bool nl_ok = h.len && in[h.s + h.len - 1] == '\n';
        if (p.mate == 0) {
            if (!(h.len && in[h.s] == '@')) kind = K_BAD_HEADER;  // :118-120
            // bytes >= 0x80: plain demultiplex copies the three lines as they are, so valid UTF-8 is data like any other; the
            // fused quality trim works char by char on bases and qualities, which therefore stay ASCII
            if (!kind && p.a.info->high &&
                !(fused ? dm_utf8_fine(p.a, h, sq, pl, ql) : utf8_ok(in + h.s, h.len + sq.len + pl.len + ql.len)))
                kind = K_NON_ASCII;
            if (!kind && fused && !nl_ok) kind = K_TRUNC_FUSED;
            uint32_t st = 0, en = 0;
            RecRef ir0{0, 0, 0}, ir1{0, 0, 0};
            uint32_t sep = 0;
            if (!kind && p.n_index) {  // :126-136
                for (uint32_t q = 0; q < p.n_index && !kind; q++) {
                    if ((unsigned long long)i >= p.ext_stats[q]->n_records) {
                        kind = K_INDEX_ASSERT;
                        break;
                    }
                    const RecRef t = p.ext_tab[q][i];
                    if (!(t.flags & RR_L0_AT) || !(t.flags & RR_L2_PLUS)) kind = K_INDEX_ASSERT;  // :130,:134
                    if (t.flags & RR_LONG) kind = K_TOO_LONG;
                    if (q == 0) ir0 = t;
                    else ir1 = t;
                }
                if (!kind) {
                    uint32_t bclen = ir0.seq_len;
                    if (p.n_index == 2) {
                        sep = bclen ? 1u : 0u;  // '+' only behind a non-empty first part (:128)
                        bclen += sep + ir1.seq_len;
                    }
                    if (bclen != Lb) kind = K_BC_LEN;  // :148-150
                }
            } else {
                if (!kind && !dm_find_bc(in + h.s, h.len, lut, st, en)) kind = K_NO_BC;  // :138-141
                if (!kind && en - st - 4u != Lb) kind = K_BC_LEN;                         // :148-150
            }
            if (!kind) {
                const uint8_t *hb = in + h.s + st + 4u;
                auto obs = [&](uint32_t q) -> uint8_t {
                    if (!p.n_index) return hb[q];
                    if (q < ir0.seq_len) return p.ext_data[0][(uint64_t)ir0.seq_off + q];
                    if (q < ir0.seq_len + sep) return (uint8_t)'+';
                    return p.ext_data[1][(uint64_t)ir1.seq_off + (q - ir0.seq_len - sep)];
                };
                unsigned long long o0 = 0, o1 = 0, o2 = 0;
                for (uint32_t q = 0; q < Lb; q++) {
                    const unsigned long long code = lut[obs(q)] & 7u;
                    o0 |= (code & 1ull) << q;
                    o1 |= ((code >> 1) & 1ull) << q;
                    o2 |= ((code >> 2) & 1ull) << q;
                }
                uint32_t lowest = 0xFFFFFFFFu, best = 0, last = 0;  // :154-166
                for (uint32_t s2 = 0; s2 < S; s2++) {
                    unsigned long long p0, p1, p2, care;
                    if (p.sheet.wide) {
                        const unsigned long long *pw = (const unsigned long long *)p.sheet.planes + 4ull * s2;
                        p0 = pw[0], p1 = pw[1], p2 = pw[2], care = pw[3];
                    } else {
                        const uint32_t *pw = p.sheet.planes + 4ull * s2;
                        p0 = pw[0], p1 = pw[1], p2 = pw[2], care = pw[3];
                    }
                    const uint32_t d = (uint32_t)__popcll(((o0 ^ p0) | (o1 ^ p1) | (o2 ^ p2)) & care);
                    if (d < lowest) {
                        lowest = d;
                        best = s2;
                        last = s2;
                    } else if (d == lowest) {
                        last = s2;
                    }
                }
                atomicAdd(&p.counts[S], 1ull);  // :169
                if (S && lowest <= 1u) {         // :172
                    if (best == last) {          // :173-178
                        sample = (int)best;
                        atomicAdd(&p.counts[S + 1], 1ull);
                        atomicAdd(&p.counts[best], 1ull);
                    } else {                     // :184-188
                        const uint32_t ei = atomicAdd(&p.st->n_events, 1u);
                        if (ei < p.events_cap) {
                            Event ev;
                            ev.record = i;
                            ev.bc_off = p.n_index ? ir0.seq_off : h.s + st + 4u;
                            ev.bc_off2 = p.n_index == 2 ? ir1.seq_off : 0xFFFFFFFFu;
                            ev.best = (int16_t)best;
                            ev.last = (int16_t)last;
                            ev.mismatches = lowest;
                            p.events[ei] = ev;
                        } else {
                            atomicOr(&p.st->flags, F_EVENTS_OVERFLOW);
                        }
                    }
                }
                if (sample >= 0) {
                    // UMI = observed characters where the sheet barcode has 'U' (:200-203), parked for mate 2
                    unsigned long long um = p.sheet.wide ? ((const unsigned long long *)p.sheet.umask)[sample]
                                                         : (unsigned long long)p.sheet.umask[sample];
                    uint32_t t = 0;
                    while (um) {
                        const uint32_t q = (uint32_t)__ffsll((long long)um) - 1u;
                        um &= um - 1ull;
                        p.umi[(uint64_t)i * p.sheet.Umax + t++] = obs(q);
                    }
                    uint32_t body = sq.len + pl.len + ql.len;  // three lines verbatim (:209-212)
                    if (fused) {
                        if (!dm_trimq(in, sq, ql, p.fused_trim, kept)) {
                            kind = K_SEQ_SHORT;
                            sample = -1;
                        }
                        body = kept ? 2u * kept + 4u : 6u;
                    }
                    if (!kind && p.out) {
                        uint32_t alen, blen;
                        dm_pieces(in + h.s, h.len, p.n_index == 0, st, en, alen, blen);
                        olen = alen + blen + (t ? 5u + t : 0u) + 1u + body;  // :206-212
                    }
                }
            }
            p.assign[i] = (int16_t)sample;
        } else {
            // mate 2 of an assigned pair (:215-237)
            sample = (unsigned long long)i < n1 ? (int)p.assign[i] : -1;
            if (sample >= 0 && p.out) {
                if (p.a.info->high &&
                    !(fused ? dm_utf8_fine(p.a, h, sq, pl, ql) : utf8_ok(in + h.s, h.len + sq.len + pl.len + ql.len)))
                    kind = K_NON_ASCII;
                if (!kind && fused && !nl_ok) kind = K_TRUNC_FUSED;
                if (!kind) {
                    uint32_t st = 0, en = 0, alen, blen;
                    const bool cut = p.n_index == 0 && dm_find_bc(in + h.s, h.len, lut, st, en);  // :219-227
                    dm_pieces(in + h.s, h.len, cut, st, en, alen, blen);                           // :229
                    const uint32_t ul = p.sheet.wide ? (uint32_t)__popcll(((const unsigned long long *)p.sheet.umask)[sample])
                                                     : (uint32_t)__popc(p.sheet.umask[sample]);
                    uint32_t body = sq.len + pl.len + ql.len;
                    if (fused) {
                        if (!dm_trimq(in, sq, ql, p.fused_trim, kept)) kind = K_SEQ_SHORT;
                        body = kept ? 2u * kept + 4u : 6u;
                    }
                    if (!kind) olen = alen + blen + (ul ? 5u + ul : 0u) + 1u + body;
                }
            }
        }
        if (!kind && olen > 0xFFFFu) kind = K_TOO_LONG;  // a group's length is 16 bits
        if (kind) {
            report_err(p.st, i, kind);
            olen = 0;
        }
        p.out_len[i] = olen;
        p.kk[i] = kept;
    }
}
__global__ void sk_dm_finish_kernel(const DParams p) {
    const uint32_t nu = (p.a.info->overflow || (uint64_t)dm_units(p) > p.max_records) ? 0u : dm_units(p);
    unsigned long long bytes = 0;
    if (nu) bytes = p.dst[nu - 1] + p.out_len[nu - 1];
    p.st->n_records = nu;
    p.st->n_lines = p.a.info->n_lines;
    p.st->out_bytes = bytes;
    p.st->out_cursor = bytes;
    p.st->consumed = line_of(p.a, nu * 4u).s;
    if (bytes > p.out_cap) report_err(p.st, 0, K_OUT_OVERFLOW);
}
__global__ void __launch_bounds__(256) sk_dm_emit_kernel(const DParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t wpb = blockDim.x >> 5, gw = blockIdx.x * wpb + (threadIdx.x >> 5), nw = gridDim.x * wpb;
    if (p.st->err_key || p.a.info->overflow) return;  // the host replays the batch up to the failing record (K_TOO_MANY included)
    const uint32_t nu = dm_units(p);
    const uint8_t *in = p.a.in;
    const uint8_t *lut = p.sheet.lut;
    uint8_t *out = p.out;
    for (uint32_t i = gw; i < nu; i += nw) {
        const uint32_t olen = p.out_len[i];
        const int sample = (int)p.assign[i];
        if (lane == 0) {
            Group g;
            g.sample = (uint16_t)(olen ? sample : 0xFFFF);
            g.len = (uint16_t)olen;
            p.groups[i] = g;
            if ((i & 31u) == 0u) {
                ChunkRow row;
                row.base = p.dst[i];
                row.first_group = i;
                row.n_groups = min(32u, nu - i);
                p.rows[i >> 5] = row;
            }
        }
        if (!olen || !out) continue;
        const LineRef h = line_of(p.a, i * 4u), sq = line_of(p.a, i * 4u + 1u), pl = line_of(p.a, i * 4u + 2u), ql = line_of(p.a, i * 4u + 3u);
        uint32_t st = 0, en = 0, alen, blen;
        const bool cut = p.n_index == 0 && dm_find_bc(in + h.s, h.len, lut, st, en);
        dm_pieces(in + h.s, h.len, cut, st, en, alen, blen);
        const uint32_t ul = p.sheet.wide ? (uint32_t)__popcll(((const unsigned long long *)p.sheet.umask)[sample])
                                         : (uint32_t)__popc(p.sheet.umask[sample]);
        unsigned long long d = p.dst[i];
        if (alen) warp_copy_piece(in, out, h.s, d, alen, lane);
        d += alen;
        if (blen) warp_copy_piece(in, out, (unsigned long long)h.s + en, d, blen, lane);
        d += blen;
        if (ul) {  // " UMI:" + umi (:207 / :230)
            put_lit(out, d, " UMI:", 5, lane);
            for (uint32_t t = (uint32_t)lane; t < ul; t += 32u) out[d + 5u + t] = p.umi[(uint64_t)i * p.sheet.Umax + t];
            d += 5u + ul;
        }
        put_lit(out, d, "\n", 1, lane);
        d += 1;
        const uint32_t kept = p.kk[i];
        if (kept == 0xFFFFFFFFu) {  // three lines verbatim: they are contiguous in the stream
            const uint32_t body = sq.len + pl.len + ql.len;
            if (body) warp_copy_piece(in, out, sq.s, d, body, lane);
        } else if (kept == 0) {
            put_lit(out, d, "N\n+\n!\n", 6, lane);
        } else {
            warp_copy_piece(in, out, sq.s, d, kept, lane);
            put_lit(out, d + kept, "\n+\n", 3, lane);
            warp_copy_piece(in, out, ql.s, d + kept + 3u, kept, lane);
            put_lit(out, d + 2ull * kept + 3u, "\n", 1, lane);
        }
    }
}

// ---- record table of an index / barcode stream ------------------------------------------------------
// The (seq_off, seq_len after trim_end, flags) table of OP_SCAN (sk_kernels.cu) from the global line table: no limit
// on record length or density (a FASTQ of 8-base index reads has hundreds of records per 16 KiB).  Record counting
// follows the chunk engines: with final_batch every record whose first line exists, otherwise only records followed
// by the start of another line (the rest is the next batch's).
__global__ void __launch_bounds__(256) sk_recref_kernel(const LStream a, uint32_t lpr, uint32_t head_char, uint64_t rec_limit, uint32_t final_batch,
                                                        RecRef *out, uint4 *inline32, uint64_t cap, DevStats *st) {
    const uint32_t nl = a.info->n_lines;
    if (a.info->overflow) {  // more lines than the line table holds: more records than sk_limits.max_records
        if (blockIdx.x == 0 && threadIdx.x == 0) report_err(st, cap, K_TOO_MANY);
        return;
    }
    uint32_t nrec = final_batch ? (nl + lpr - 1u) / lpr : (nl ? (nl - 1u) / lpr : 0u);
    if ((uint64_t)nrec > rec_limit) nrec = (uint32_t)rec_limit;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_lines = nl;
        st->n_records = nrec;
        st->consumed = line_of(a, nrec * lpr).s;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += gridDim.x * blockDim.x) {
        const LineRef h = line_of(a, i * lpr), sq = line_of(a, i * lpr + 1u);
        uint32_t sl = trim_end_len_dev(a.in + sq.s, sq.len);
        const uint32_t c0 = h.len ? a.in[h.s] : 0x100u;
        uint16_t fl = 0;
        if (c0 == '@') fl |= RR_L0_AT;
        if (c0 == '>') fl |= RR_L0_GT;
        if (lpr == 4) {
            const LineRef pl = line_of(a, i * lpr + 2u);
            if (pl.len && a.in[pl.s] == '+') fl |= RR_L2_PLUS;
        }
        if (sl > 0xFFFFu) {
            fl |= RR_LONG;
            sl = 0xFFFFu;
        }
        if (head_char && c0 != head_char) report_err(st, i, K_MIXED);
        if (i < cap) {
            RecRef rr;
            rr.seq_off = sq.s;
            rr.seq_len = (uint16_t)sl;
            rr.flags = fl;
            out[i] = rr;
            if (inline32) {  // the barcode itself, for readers that want it at a predictable address (add barcode)
                uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                const uint32_t m = sl < 32u ? sl : 32u;
                for (uint32_t t = 0; t < m; t++) w[t >> 2] |= (uint32_t)a.in[sq.s + t] << (8u * (t & 3u));
                inline32[2ull * i] = make_uint4(w[0], w[1], w[2], w[3]);
                inline32[2ull * i + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
        }
    }
}
// Line table `k` (0 or 1) of the work area is used.  Returns the number of launches, < 0 on a launch error.
int launch_scan_table(const uint8_t *in, uint64_t n, uint32_t lpr, uint32_t head_char, uint64_t rec_limit, uint32_t final_batch,
                      RecRef *out, uint4 *inline32, uint64_t cap, void *work, int k, uint64_t max_stream_bytes, uint64_t max_records,
                      DevStats *st, int sm_count, void *stream_, const char **err) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const LineWork w = carve(work, max_stream_bytes, max_records);
    int launches = index_lines(in, n, lpr, w, k, st, 1, stream);
    LStream a;
    a.in = in, a.n = n, a.starts = w.starts[k], a.info = w.info[k];
    const uint64_t want = (n / 64 + 255) / 256 + 1;  // a thread per 64 bytes is plenty
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)sm_count * 8, want));
    sk_recref_kernel<<<grid, 256, 0, stream>>>(a, lpr, head_char, rec_limit, final_batch, out, inline32, cap, st);
    launches++;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    return launches;
}

// Header-route demultiplex of one mate on the line engine.  mate 0 must run before mate 1 (assign[], umi[], r1_stats).
// Line table k = mate.  Returns the number of launches, < 0 on a launch error; *n_rows = slice-table rows in use.
int launch_line_demux(int mate, const uint8_t *in, uint64_t n, uint64_t rec_limit, int fused_trim, const SheetDev &sheet, int16_t *assign,
                      uint8_t *umi, Group *groups, ChunkRow *rows, uint32_t max_rows, unsigned long long *counts, Event *events,
                      uint32_t events_cap, const DevStats *r1_stats, uint8_t *out, uint64_t out_cap, void *work,
                      uint64_t max_stream_bytes, uint64_t max_records, DevStats *st, int sm_count, void *stream_, uint32_t *n_rows,
                      const char **err, uint32_t n_index, const RecRef *const *ext_tab, const uint8_t *const *ext_data,
                      const DevStats *const *ext_stats) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const LineWork w = carve(work, max_stream_bytes, max_records);
    int launches = index_lines(in, n, 4, w, mate, st, 0, stream);
    DParams p;
    memset(&p, 0, sizeof p);
    p.a.in = in, p.a.n = n, p.a.starts = w.starts[mate], p.a.info = w.info[mate];
    p.mate = (uint32_t)mate;
    p.rec_limit = rec_limit ? rec_limit : ~0ull;
    p.max_records = max_records;
    p.fused_trim = fused_trim;
    p.sheet = sheet;
    p.assign = assign, p.umi = umi, p.groups = groups, p.rows = rows;
    p.max_rows = (uint32_t)std::min<uint64_t>(max_rows, (max_records + 31) / 32);
    p.counts = counts, p.events = events, p.events_cap = events_cap, p.r1_stats = r1_stats;
    p.n_index = n_index;
    for (uint32_t q = 0; q < n_index && q < 2; q++) p.ext_tab[q] = ext_tab[q], p.ext_data[q] = ext_data[q], p.ext_stats[q] = ext_stats[q];
    p.out_len = w.out_len, p.kk = w.bc_ref, p.dst = w.dst;
    p.out = out, p.out_cap = out_cap, p.st = st;
    *n_rows = p.max_rows;
    cudaMemsetAsync(rows, 0, (size_t)p.max_rows * sizeof(ChunkRow), stream);
    const unsigned grid = (unsigned)std::max(1, sm_count * 8);
    const uint32_t n_scan = (uint32_t)std::min<uint64_t>(max_records, 0xFFFFFFFFull);
    sk_dm_plan_kernel<<<grid, 256, 0, stream>>>(p);
    launches += 1 + launch_len_scan(w.out_len, w.dst, n_scan, w.bsum, stream);
    sk_dm_finish_kernel<<<1, 1, 0, stream>>>(p);
    sk_dm_emit_kernel<<<grid, 256, 0, stream>>>(p);
    launches += 2;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    return launches;
}

// One line operator over the slot's streams.  stats table: keys / rep / cnt of `h_cap` slots, list of 2*h_cap u64.
int launch_lineop(int op, const uint8_t *in_a, uint64_t n_a, const uint8_t *in_b, uint64_t n_b, uint32_t lpr, uint32_t head, uint32_t x,
                  uint32_t y, uint64_t rec_limit, uint8_t *out0, uint8_t *out1, uint64_t out_cap, void *work, uint64_t max_stream_bytes,
                  uint64_t max_records, void *stats_tab, uint32_t h_cap, DevStats *st, int sm_count, void *stream_, const char **err,
                  const RecRef *bc_tab, const DevStats *bc_stats) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const LineWork w = carve(work, max_stream_bytes, max_records);
    const int strict = (op == LOP_TRIMQ || op == LOP_MASKQ || op == LOP_ADDBC) ? 0 : 1;
    int launches = index_lines(in_a, n_a, lpr, w, 0, st, strict, stream);
    if (op == LOP_INTERLEAVE) launches += index_lines(in_b, n_b, lpr, w, 1, st, strict, stream);
    LParams p;
    memset(&p, 0, sizeof p);
    p.op = op;
    p.a.in = in_a, p.a.n = n_a, p.a.starts = w.starts[0], p.a.info = w.info[0];
    p.b.in = in_b, p.b.n = n_b, p.b.starts = w.starts[1], p.b.info = w.info[1];
    p.lpr = lpr, p.head = head, p.x = x, p.y = y;
    p.rec_limit = rec_limit ? rec_limit : ~0ull;
    p.max_records = max_records;
    p.out_len = w.out_len, p.dst = w.dst, p.out = out0, p.out_cap = out_cap, p.st = st;
    p.bc_ref = w.bc_ref;
    p.bc_tab = bc_tab, p.bc_stats = bc_stats;
    const unsigned grid = (unsigned)std::max(1, sm_count * 8);
    const uint32_t n_scan = (uint32_t)std::min<uint64_t>(max_records, 0xFFFFFFFFull);
    if (op == LOP_STATS) {
        unsigned long long *keys = (unsigned long long *)stats_tab, *rep = keys + h_cap;
        uint32_t *cnt = (uint32_t *)(rep + h_cap);
        unsigned long long *list = (unsigned long long *)(cnt + h_cap);
        uint32_t *n_list = (uint32_t *)(list + 2ull * h_cap);
        cudaMemsetAsync(keys, 0, (size_t)h_cap * 8, stream);
        cudaMemsetAsync(rep, 0xFF, (size_t)h_cap * 8, stream);
        cudaMemsetAsync(cnt, 0, (size_t)h_cap * 4, stream);
        cudaMemsetAsync(n_list, 0, 4, stream);
        p.h_keys = keys, p.h_rep = rep, p.h_cnt = cnt, p.h_mask = h_cap - 1u;
        p.out_len = nullptr;
        sk_line_plan_kernel<<<grid, 256, 0, stream>>>(p);
        sk_stats_verify_kernel<<<grid, 256, 0, stream>>>(p);
        sk_line_finish_kernel<<<1, 1, 0, stream>>>(p, 0);
        sk_stats_list_kernel<<<grid, 256, 0, stream>>>(keys, rep, cnt, h_cap, list, n_list);
        launches += 4;
    } else if (op == LOP_CHECK) {
        p.out_len = nullptr;
        sk_line_plan_kernel<<<grid, 256, 0, stream>>>(p);
        sk_line_finish_kernel<<<1, 1, 0, stream>>>(p, 0);
        launches += 2;
    } else {
        const int passes = op == LOP_DEINTERLEAVE ? 2 : 1;
        for (int pass = 0; pass < passes; pass++) {
            if (op == LOP_DEINTERLEAVE) {
                p.x = (uint32_t)pass;
                p.out = pass == 0 ? out0 : out1;
            }
            sk_line_plan_kernel<<<grid, 256, 0, stream>>>(p);
            launches += 1 + launch_len_scan(w.out_len, w.dst, n_scan, w.bsum, stream);
            sk_line_finish_kernel<<<1, 1, 0, stream>>>(p, pass);
            sk_line_emit_kernel<<<grid, 256, 0, stream>>>(p);
            launches += 2;
        }
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    return launches;
}

}  // namespace sk
