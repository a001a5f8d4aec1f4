// sk_kernels.cu -- the single-pass chunk engine, hand-written for sm_100a.
//
// One persistent kernel template implements every operator of the path:
//   OP_SCAN / OP_TRIM / OP_MASK / OP_ADDBC / OP_DEMUX1 / OP_DEMUX2.
// Input bytes are read from HBM once (TMA bulk copy into shared memory), output bytes are written
// once (16-byte vector stores from a shared staging image).  Record framing is by global line
// index (decoupled look-back over per-chunk line counts), exactly like the reference's four
// read_line calls per record (common.rs:106-112).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sk_internal.h"

namespace sk {

extern __shared__ __align__(128) unsigned char sk_smem[];

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// look-back words: 2-bit status | 62-bit value, one 8-byte relaxed gpu-scope access
constexpr uint64_t TS_INVALID = 0, TS_AGG = 1, TS_INC = 2;
constexpr uint64_t TS_VMASK = (1ull << 62) - 1;
__device__ __forceinline__ void ts_store(uint64_t *p, uint64_t status, uint64_t v) {
    uint64_t w = (status << 62) | (v & TS_VMASK);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint64_t ts_load(const uint64_t *p) {
    uint64_t w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}

__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Decoupled look-back (one warp).  Publishes this chunk's aggregate, returns the exclusive prefix
// over all earlier chunks and publishes the inclusive prefix.  Chunks are handed out by a ticket
// counter, so every predecessor is owned by a resident CTA that never waits on a later chunk.
__device__ uint64_t lookback(uint64_t *tiles, uint32_t c, uint64_t agg, int lane) {
    if (c == 0) {
        if (lane == 0) ts_store(&tiles[0], TS_INC, agg);
        return 0;
    }
    if (lane == 0) ts_store(&tiles[c], TS_AGG, agg);
    uint64_t excl = 0;
    int64_t base = (int64_t)c - 1;
    for (;;) {
        int64_t idx = base - lane;
        uint64_t w = TS_INC << 62;  // chunks before 0: inclusive prefix 0
        if (idx >= 0) {
            do {
                w = ts_load(&tiles[idx]);
            } while ((w >> 62) == TS_INVALID);
        }
        uint32_t inc = __ballot_sync(0xffffffffu, (w >> 62) == TS_INC);
        uint64_t v = w & TS_VMASK;
        if (inc) {
            int first = __ffs(inc) - 1;  // nearest predecessor holding an inclusive prefix
            excl += warp_sum64(lane <= first ? v : 0);
            break;
        }
        excl += warp_sum64(v);
        base -= 32;
    }
    if (lane == 0) ts_store(&tiles[c], TS_INC, excl + agg);
    return excl;
}

// Exclusive block scan of one u32 per thread; `scratch` holds NT/32+1 words.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *scratch, uint32_t &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) scratch[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t t = lane < NW ? scratch[lane] : 0, s = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < NW) scratch[lane] = s - t;
        if (lane == NW - 1) scratch[NW] = s;
    }
    __syncthreads();
    total = scratch[NW];
    uint32_t r = scratch[w] + x - v;
    __syncthreads();
    return r;
}

// 4-bit mask of the bytes of x equal to '\n' (exact SWAR zero-byte test, no cross-byte carries).
__device__ __forceinline__ uint32_t nl4(uint32_t x) {
    uint32_t t = x ^ 0x0A0A0A0Au;
    uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
    return (((z >> 7) * 0x00204081u) >> 21) & 0xFu;
}
// bits b of a 16-bit piece mask whose window position o+b lies in [lo, hi)
__device__ __forceinline__ uint32_t range16(uint32_t o, uint32_t lo, uint32_t hi) {
    uint32_t a = lo > o ? lo - o : 0u, b = hi > o ? hi - o : 0u;
    if (a > 16u) a = 16u;
    if (b > 16u) b = 16u;
    return b > a ? (((1u << b) - 1u) & ~((1u << a) - 1u)) : 0u;
}

__device__ __forceinline__ bool is_ws(uint8_t c) { return c == 32u || (c >= 9u && c <= 13u); }  // ASCII White_Space
__device__ __forceinline__ bool is_bc_class(uint8_t c) {  // [ACGTNacgtn+], fasta_demultiplex.rs:38
    switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'N':
        case 'a': case 'c': case 'g': case 't': case 'n': case '+':
            return true;
    }
    return false;
}

__device__ __forceinline__ void report_err(DevStats *st, uint64_t rec, unsigned kind) {
    atomicMax(&st->err_key, ~((rec << 8) | (unsigned long long)kind));
}

// body modes of a planned record
enum : uint8_t { B_VERBATIM = 0, B_TRIM = 1, B_GARBAGE = 2, B_MASK = 3, B_NONE = 4 };
// tag kinds (r_taglen high bit unused; kind is implied by OP)

struct Win {
    const uint8_t *b;
    const uint16_t *ls;
    uint32_t nls;   // line starts found in the window (may exceed MAXLINES; indices needed are checked)
    uint32_t wlen;  // valid bytes in the window
};
__device__ __forceinline__ uint32_t lb(const Win &W, uint32_t x) { return x < W.nls ? (uint32_t)W.ls[x] : W.wlen; }

// fasta_trim_by_quality.rs:28-48 on the quality line [L3,L4) / sequence line [L1,L2).
// Returns false when &seq[..k] would panic.
__device__ __forceinline__ bool plan_trim_body(const Win &W, uint32_t L1, uint32_t L2, uint32_t L3, uint32_t L4,
                                               int minq, uint8_t &mode, uint32_t &kk, uint32_t &body_len) {
    uint32_t k = L4 - L3;
    while (k > 0 && is_ws(W.b[L3 + k - 1])) k--;  // qual.trim_end().len()  (:31)
    int total = -50, lowest = -50;                // :28-29
    uint32_t lowest_k = k;
    while (k > 0) {  // :33-42
        k--;
        total += (int)(uint8_t)(W.b[L3 + k] - 33u) - minq;  // wrapping u8 subtraction (:35)
        if (total > 0) break;
        if (total < lowest) {
            lowest = total;
            lowest_k = k;
        }
    }
    if (lowest_k == 0) {  // :44-45
        mode = B_GARBAGE;
        kk = 0;
        body_len = 6;  // "N\n+\n!\n"
        return true;
    }
    mode = B_TRIM;
    kk = lowest_k;
    body_len = 2 * lowest_k + 4;  // seq[..k] "\n+\n" qual[..k] "\n"  (:47)
    return lowest_k <= L2 - L1;
}

// Leftmost match of " BC:[ACGTNacgtn+]+" in [h0,h1) (window offsets). Returns false if none.
__device__ __forceinline__ bool bc_find(const uint8_t *b, uint32_t h0, uint32_t h1, uint32_t &st, uint32_t &en) {
    for (uint32_t i = h0; i + 5 <= h1; i++) {
        if (b[i] == ' ' && b[i + 1] == 'B' && b[i + 2] == 'C' && b[i + 3] == ':' && is_bc_class(b[i + 4])) {
            uint32_t e = i + 5;
            while (e < h1 && is_bc_class(b[e])) e++;
            st = i;
            en = e;
            return true;
        }
    }
    return false;
}

// header.drain(cut) then trim_end(): the kept pieces are [h0, h0+alen) and [c1, c1+blen).
__device__ __forceinline__ void header_pieces(const uint8_t *b, uint32_t h0, uint32_t h1, uint32_t c0, uint32_t c1,
                                              uint32_t &alen, uint32_t &blen) {
    uint32_t e = h1;
    while (e > c1 && is_ws(b[e - 1])) e--;
    if (e > c1) {
        alen = c0 - h0;
        blen = e - c1;
        return;
    }
    blen = 0;
    e = c0;
    while (e > h0 && is_ws(b[e - 1])) e--;
    alen = e - h0;
}

template <typename WT>
__device__ __forceinline__ uint32_t popcw(WT x);
template <>
__device__ __forceinline__ uint32_t popcw<uint32_t>(uint32_t x) { return __popc(x); }
template <>
__device__ __forceinline__ uint32_t popcw<uint64_t>(uint64_t x) { return __popcll(x); }

// Replaces the S x barcode_diff loop of fasta_demultiplex.rs:157-166.
template <typename WT>
__device__ __forceinline__ void match_sheet(const WT *sh, uint32_t S, WT o0, WT o1, WT o2, uint32_t &lowest,
                                            uint32_t &best, uint32_t &last) {
    lowest = 0xFFFFFFFFu;
    best = 0;
    last = 0;
    for (uint32_t s = 0; s < S; s++) {
        const WT p0 = sh[4 * s], p1 = sh[4 * s + 1], p2 = sh[4 * s + 2], care = sh[4 * s + 3];
        const uint32_t d = popcw<WT>(((o0 ^ p0) | (o1 ^ p1) | (o2 ^ p2)) & care);
        if (d < lowest) {
            lowest = d;
            best = s;
            last = s;
        } else if (d == lowest) {
            last = s;
        }
    }
}

__device__ __forceinline__ void wcopy(uint8_t *dst, const uint8_t *src, uint32_t len, int lane) {
    for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// the chunk engine
// ------------------------------------------------------------------------------------------------
struct Misc {
    uint64_t mbar;
    uint64_t g0;        // global index of the first line that starts in this chunk
    uint64_t out_base;  // where this chunk's output goes
    uint32_t chunk;
    uint32_t chunk_out;
    uint32_t cta_total, cta_ident;
    uint32_t scratch[40];
};

template <class Cfg, int OP, typename WT>
__global__ void __launch_bounds__(Cfg::NT, 2) sk_chunk_kernel(const __grid_constant__ KParams p) {
    constexpr int NT = Cfg::NT, NW = NT / 32, MAXREC = Cfg::MAXREC, MAXLINES = Cfg::MAXLINES;
    constexpr bool WIDE = sizeof(WT) == 8;
    constexpr bool IS_DEMUX = (OP == OP_DEMUX1 || OP == OP_DEMUX2);
    constexpr bool ORDERED = (OP == OP_TRIM || OP == OP_MASK || OP == OP_ADDBC);
    constexpr bool HAS_OUT = ORDERED || IS_DEMUX;
    static_assert(MAXREC <= NT, "one planning thread per record");

    const uint32_t S = IS_DEMUX ? p.sheet.S : 0u;
    const SmemLayout SL = smem_layout<Cfg>(S, WIDE ? 1u : 0u);
    uint8_t *win = sk_smem + SL.win;
    uint8_t *stage = sk_smem + SL.stage;
    uint16_t *ls = (uint16_t *)(sk_smem + SL.ls);
    uint32_t *r_outoff = (uint32_t *)(sk_smem + SL.rec);
    uint32_t *r_outlen = r_outoff + MAXREC;
    uint32_t *r_ext = r_outlen + MAXREC;
    uint16_t *r_alen = (uint16_t *)(r_ext + MAXREC);
    uint16_t *r_cut1 = r_alen + MAXREC;
    uint16_t *r_blen = r_cut1 + MAXREC;
    uint16_t *r_k = r_blen + MAXREC;
    uint16_t *r_taglen = r_k + MAXREC;
    int16_t *r_sample = (int16_t *)(r_taglen + MAXREC);
    uint8_t *r_mode = (uint8_t *)(r_sample + MAXREC);
    WT *sh_planes = (WT *)(sk_smem + SL.sheet);
    WT *sh_umask = (WT *)(sk_smem + SL.umask);
    uint8_t *sh_lut = sk_smem + SL.lut;
    uint32_t *hist = (uint32_t *)(sk_smem + SL.hist);
    uint32_t *sbase = (uint32_t *)(sk_smem + SL.sbase);
    uint32_t *ccount = (uint32_t *)(sk_smem + SL.ccount);
    Misc *M = (Misc *)(sk_smem + SL.misc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevStats *st = p.stats;

    // ---- one-time per CTA: mbarrier, sheet -> shared memory
    if (tid == 0) {
        mbar_init(&M->mbar, 1);
        M->cta_total = 0;
        M->cta_ident = 0;
    }
    if (IS_DEMUX) {
        const WT *gp = (const WT *)p.sheet.planes;
        for (uint32_t i = tid; i < 4 * S; i += NT) sh_planes[i] = gp[i];
        const WT *gu = (const WT *)p.sheet.umask;
        for (uint32_t i = tid; i < S; i += NT) {
            sh_umask[i] = gu[i];
            ccount[i] = 0;
        }
        if (tid < 256) sh_lut[tid] = p.sheet.lut[tid];
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    uint32_t parity = 0;
    for (;;) {
        // ---- P0 ticket
        if (tid == 0) M->chunk = atomicAdd(&st->ticket, 1u);
        __syncthreads();
        const uint32_t c = M->chunk;
        if (c >= p.n_chunks) break;

        const uint64_t c0 = (uint64_t)c * Cfg::CHUNK;
        const uint64_t w0 = c0 ? c0 - Cfg::PRE : 0;
        const uint32_t cs = (uint32_t)(c0 - w0);  // window offset of the chunk's first byte
        uint64_t wend = c0 + Cfg::CHUNK + Cfg::OVERHANG;
        if (wend > p.n) wend = p.n;
        const uint32_t wlen = (uint32_t)(wend - w0);
        const bool at_end = (wend == p.n);
        const uint32_t ce = cs + Cfg::CHUNK;  // window offset one past the chunk

        // ---- P1 load the window: TMA bulk copy for the 16-byte multiple, plain loads for the tail
        const uint32_t bulk = wlen & ~15u;
        if (tid == 0 && bulk) {
            fence_proxy_async();  // order earlier generic-proxy reads of `win` before the async write
            mbar_expect_tx(&M->mbar, bulk);
            bulk_g2s(win, p.in + w0, bulk, &M->mbar);
        }
        if (tid < 16) {  // tail bytes; zero padding so the last partial piece reads defined data
            const uint32_t o = bulk + tid;
            if (o < (uint32_t)Cfg::WIN_MAX) win[o] = (o < wlen) ? p.in[w0 + o] : (uint8_t)0;
        }
        if (bulk) mbar_wait(&M->mbar, parity);
        if (bulk) parity ^= 1;
        __syncthreads();

        // ---- P2 newline scan: 80 contiguous bytes per thread
        const uint32_t ls_lo = cs ? cs - 1 : 0;              // newline at q starts a line at q+1 >= cs
        const uint32_t ls_hi = at_end ? (wlen ? wlen - 1 : 0) : wlen;  // a final '\n' starts no line in this buffer
        uint32_t m16[Cfg::PPL];
        uint32_t cnt_all = 0, cnt_chunk = 0, hib = 0;
#pragma unroll
        for (int q = 0; q < Cfg::PPL; q++) {
            const uint32_t o = (uint32_t)tid * (Cfg::PPL * 16) + q * 16;
            uint32_t m = 0;
            if (o < wlen) {
                const uint4 v = *(const uint4 *)(win + o);
                m = nl4(v.x) | (nl4(v.y) << 4) | (nl4(v.z) << 8) | (nl4(v.w) << 12);
                m &= range16(o, ls_lo, ls_hi);
                hib |= (v.x | v.y | v.z | v.w) & 0x80808080u;  // bytes past wlen in the last piece are zero
            }
            m16[q] = m;
            cnt_all += __popc(m);
            cnt_chunk += __popc(m & range16(o, 0, ce - 1));
        }
        if (__any_sync(0xffffffffu, hib != 0) && lane == 0) atomicOr(&st->flags, F_NON_ASCII);

        uint32_t tot;
        const uint32_t pre = block_excl_scan<NT>((cnt_chunk << 16) | cnt_all, M->scratch, tot);
        const uint32_t extra = (c0 == 0) ? 1u : 0u;  // the line that starts at byte 0
        const uint32_t nls = (tot & 0xFFFFu) + extra;       // line starts in [cs, wlen)
        const uint32_t nls_chunk = (tot >> 16) + extra;     // ... of which inside the chunk

        // ---- P3 line-start table + look-back for the global line index
        {
            uint32_t idx = (pre & 0xFFFFu) + extra;
            if (tid == 0 && extra) ls[0] = 0;
#pragma unroll
            for (int q = 0; q < Cfg::PPL; q++) {
                uint32_t m = m16[q];
                const uint32_t o = (uint32_t)tid * (Cfg::PPL * 16) + q * 16;
                while (m) {
                    const uint32_t b2 = __ffs(m) - 1;
                    m &= m - 1;
                    if (idx < (uint32_t)MAXLINES) ls[idx] = (uint16_t)(o + b2 + 1);
                    idx++;
                }
            }
        }
        if (warp == 0) {
            const uint64_t excl = lookback(p.tile_lines, c, nls_chunk, lane);
            if (lane == 0) {
                M->g0 = excl;
                if (c == p.n_chunks - 1) st->n_lines = excl + nls_chunk;
            }
        }
        __syncthreads();

        // ---- P4 which records does this chunk own?
        const uint64_t g0 = M->g0;
        const uint32_t lpr = p.lpr;
        const uint32_t j0 = (uint32_t)((lpr - (g0 % lpr)) % lpr);
        const uint64_t rec0 = (g0 + j0) / lpr;
        uint32_t nrec = j0 < nls_chunk ? (nls_chunk - 1 - j0) / lpr + 1 : 0;
        if (rec0 >= p.rec_limit) nrec = 0;
        else if ((uint64_t)nrec > p.rec_limit - rec0) nrec = (uint32_t)(p.rec_limit - rec0);
        unsigned chunk_err = 0;
        if (nrec) {
            uint32_t jend = j0 + nrec * lpr;  // line index one past the last owned record
            const bool eof_ok = at_end && p.final_batch;
            if (jend >= nls && !eof_ok) {
                if (at_end) {  // non-final batch: leave the trailing incomplete record(s) to the next batch
                    nrec = nls > j0 + lpr ? (nls - j0 - lpr - 1) / lpr + 1 : 0;
                    jend = j0 + nrec * lpr;
                } else {
                    chunk_err = K_TOO_LONG;
                }
            }
            if (!chunk_err && nrec) {
                const uint32_t need = jend < nls ? jend : nls - 1;
                if (need >= (uint32_t)MAXLINES || nrec > (uint32_t)MAXREC) chunk_err = K_TOO_DENSE;
            }
            if (chunk_err) {
                if (tid == 0) report_err(st, rec0, chunk_err);
                nrec = 0;
            }
        }
        const Win W{win, ls, nls, wlen};
        if (tid == 0 && nrec) {
            atomicAdd(&st->n_records, (unsigned long long)nrec);
            atomicMax(&st->consumed, (unsigned long long)(w0 + lb(W, j0 + nrec * lpr)));
        }

        // ---- P5 plan: one thread per record runs the reference's per-record logic
        uint32_t my_outlen = 0;
        if (tid < (int)nrec) {
            const uint32_t r = tid;
            const uint32_t j = j0 + r * lpr;
            const uint64_t rec = rec0 + r;
            const uint32_t L0 = lb(W, j), L1 = lb(W, j + 1), L2 = lb(W, j + 2);
            const uint32_t L3 = lpr == 4 ? lb(W, j + 3) : L2, L4 = lpr == 4 ? lb(W, j + 4) : L2;
            const uint32_t Lend = lpr == 4 ? L4 : L2;
            uint8_t mode = B_NONE;
            uint32_t kk = 0, outlen = 0, alen = 0, cut1 = 0, blen = 0, taglen = 0, ext = 0;
            int sample = -1;

            if (OP == OP_SCAN) {
                // (seq_off, seq_len after trim_end, flags) of an index read / barcode record
                uint32_t sl = L2 - L1;
                while (sl > 0 && is_ws(win[L1 + sl - 1])) sl--;
                uint16_t fl = 0;
                if (L1 > L0 && win[L0] == '@') fl |= RR_L0_AT;
                if (L1 > L0 && win[L0] == '>') fl |= RR_L0_GT;
                if (lpr == 4 && L3 > L2 && win[L2] == '+') fl |= RR_L2_PLUS;
                if (sl > 0xFFFFu) { fl |= RR_LONG; sl = 0xFFFFu; }
                if (p.head_char && !(L1 > L0 && win[L0] == (uint8_t)p.head_char)) report_err(st, rec, K_MIXED);
                if (rec < p.scan_cap) {
                    RecRef rr;
                    rr.seq_off = (uint32_t)(w0 + L1);
                    rr.seq_len = (uint16_t)sl;
                    rr.flags = fl;
                    p.scan_out[rec] = rr;
                }
            } else if (OP == OP_TRIM) {
                if (win[L0] != '@') {  // fasta_trim_by_quality.rs:20-22
                    report_err(st, rec, K_BAD_HEADER);
                } else {
                    uint32_t body;
                    if (!plan_trim_body(W, L1, L2, L3, L4, (int)p.min_baseq, mode, kk, body)) {
                        report_err(st, rec, K_SEQ_SHORT);
                        mode = B_NONE;
                    } else {
                        outlen = (L1 - L0) + body;  // header verbatim (:23) + body
                    }
                }
            } else if (OP == OP_MASK) {
                if (win[L0] != '@') {  // fasta_mask_by_quality.rs:21-23
                    report_err(st, rec, K_BAD_HEADER);
                } else {
                    uint32_t sl = L2 - L1, ql = L4 - L3;
                    if (sl && win[L2 - 1] == '\n') sl--;  // :32
                    if (ql && win[L4 - 1] == '\n') ql--;  // :33
                    if (sl != ql) {                       // :35-37
                        report_err(st, rec, K_LEN_MISMATCH);
                    } else {
                        mode = B_MASK;
                        kk = sl;
                        outlen = (L1 - L0) + 2 * sl + 4;  // header, masked, "\n+\n", qual, "\n"  (:26,:44)
                    }
                }
            } else if (OP == OP_ADDBC) {
                // fasta_add_barcode.rs:29-43
                const uint8_t h = win[L0];
                uint32_t e = L1;
                while (e > L0 && is_ws(win[e - 1])) e--;  // header.trim_end()
                alen = e - L0;
                // barcode of iteration i = sequence line of barcode record i; the last one is reused
                // once the barcode file is exhausted (:20-27)
                uint32_t bl = 0, bo = 0;
                const uint64_t nb = p.ext_stats[0] ? p.ext_stats[0]->n_records : 0;
                if (nb) {
                    const uint64_t bi = rec < nb ? rec : nb - 1;
                    const RecRef rr = p.ext_tab[0][bi];
                    bl = rr.seq_len;
                    bo = rr.seq_off;
                }
                taglen = 4 + bl;  // " BC:" + barcode
                ext = bo;
                if (h != (uint8_t)p.head_char) {
                    // the reference prints the BC'd header and then stops (:33 before :41-43); the host
                    // reproduces that line, the kernel only reports where.
                    report_err(st, rec, (h == '@' || h == '>') ? K_MIXED : K_BAD_FASTX_LINE);
                    taglen = 0;
                } else {
                    mode = B_VERBATIM;
                    outlen = alen + taglen + 1 + (Lend - L1);
                }
            } else if (OP == OP_DEMUX1) {
                bool ok = true;
                if (win[L0] != '@') {  // fasta_demultiplex.rs:118-120
                    report_err(st, rec, K_BAD_HEADER);
                    ok = false;
                }
                uint32_t c0h = L1, c1h = L1;  // cut [c0h, c1h) (window offsets); empty on the index route
                uint32_t bclen = 0;
                RecRef ir[2];
                uint32_t sep = 0;
                if (ok && p.n_index) {  // :126-136
                    for (uint32_t q = 0; q < p.n_index && ok; q++) {
                        if (rec >= p.ext_stats[q]->n_records) { ok = false; break; }
                        ir[q] = p.ext_tab[q][rec];
                        if (!(ir[q].flags & RR_L0_AT) || !(ir[q].flags & RR_L2_PLUS)) ok = false;
                    }
                    if (!ok) report_err(st, rec, K_INDEX_ASSERT);
                    else {
                        bclen = ir[0].seq_len;
                        if (p.n_index == 2) {
                            sep = bclen ? 1u : 0u;  // '+' only if the barcode so far is non-empty (:128)
                            bclen += sep + ir[1].seq_len;
                        }
                    }
                } else if (ok) {  // :138-146
                    if (!bc_find(win, L0, L1, c0h, c1h)) {
                        report_err(st, rec, K_NO_BC);
                        ok = false;
                    } else {
                        bclen = c1h - (c0h + 4);
                    }
                }
                if (ok && bclen != p.sheet.L) {  // :148-150
                    report_err(st, rec, K_BC_LEN);
                    ok = false;
                }
                if (ok && p.fused_trim >= 0 && !(L1 > L0 && win[L1 - 1] == '\n')) {
                    report_err(st, rec, K_TRUNC_FUSED);
                    ok = false;
                }
                if (ok) {
                    // observed barcode byte p
                    auto obs = [&](uint32_t q) -> uint8_t {
                        if (!p.n_index) return win[c0h + 4 + q];
                        if (q < ir[0].seq_len) return p.ext_data[0][(uint64_t)ir[0].seq_off + q];
                        if (q < ir[0].seq_len + sep) return (uint8_t)'+';
                        return p.ext_data[1][(uint64_t)ir[1].seq_off + (q - ir[0].seq_len - sep)];
                    };
                    WT o0 = 0, o1 = 0, o2 = 0;
                    for (uint32_t q = 0; q < bclen; q++) {
                        const uint32_t code = sh_lut[obs(q)];
                        o0 |= (WT)(code & 1u) << q;
                        o1 |= (WT)((code >> 1) & 1u) << q;
                        o2 |= (WT)((code >> 2) & 1u) << q;
                    }
                    uint32_t lowest, best, last;
                    match_sheet<WT>(sh_planes, S, o0, o1, o2, lowest, best, last);  // :154-166
                    atomicAdd(&M->cta_total, 1u);                                   // :169
                    if (lowest <= 1u) {                                             // :172
                        if (best == last) {
                            sample = (int)best;
                            atomicAdd(&M->cta_ident, 1u);  // :177
                            atomicAdd(&ccount[best], 1u);  // :178
                        } else {                           // :184-188
                            sample = -2;
                            const uint32_t ei = atomicAdd(&st->n_events, 1u);
                            if (ei < p.events_cap) {
                                Event ev;
                                ev.record = (uint32_t)rec;
                                ev.bc_off = p.n_index ? ir[0].seq_off : (uint32_t)(w0 + c0h + 4);
                                ev.bc_off2 = p.n_index == 2 ? ir[1].seq_off : 0xFFFFFFFFu;
                                ev.best = (int16_t)best;
                                ev.last = (int16_t)last;
                                ev.mismatches = lowest;
                                p.events[ei] = ev;
                            } else {
                                atomicOr(&st->flags, F_EVENTS_OVERFLOW);
                            }
                        }
                    }
                    if (sample >= 0) {
                        // UMI = observed chars where the sheet barcode has 'U' (:200-203)
                        uint32_t ul = 0;
                        WT um = sh_umask[sample];
                        while (um) {
                            const uint32_t q = WIDE ? (uint32_t)(__ffsll((long long)um) - 1) : (uint32_t)(__ffs((int)um) - 1);
                            um &= um - 1;
                            p.umi[rec * p.sheet.Umax + ul] = obs(q);
                            ul++;
                        }
                        taglen = ul ? 5 + ul : 0;  // " UMI:" + umi (:207)
                        header_pieces(win, L0, L1, c0h, c1h, alen, blen);  // drain (:145) + trim_end (:206)
                        cut1 = c1h - L0;
                        uint32_t body = Lend - L1;  // three lines verbatim (:209-212)
                        mode = B_VERBATIM;
                        bool fine = true;
                        if (p.fused_trim >= 0) fine = plan_trim_body(W, L1, L2, L3, L4, p.fused_trim, mode, kk, body);
                        if (!fine) {
                            report_err(st, rec, K_SEQ_SHORT);
                            sample = -1;
                            mode = B_NONE;
                        } else if (!p.out) {
                            mode = B_NONE;  // dry run: count only (:77-78,:179)
                        } else {
                            outlen = alen + blen + taglen + 1 + body;
                        }
                    }
                }
            } else if (OP == OP_DEMUX2) {
                // fasta_demultiplex.rs:215-237: mate 2 of an assigned pair
                sample = rec < p.r1_stats->n_records ? (int)p.assign[rec] : -1;
                if (sample >= 0 && p.out) {
                    uint32_t c0h = L1, c1h = L1;
                    if (!p.n_index) {  // :219-227
                        uint32_t a, b;
                        if (bc_find(win, L0, L1, a, b)) { c0h = a; c1h = b; }
                    }
                    header_pieces(win, L0, L1, c0h, c1h, alen, blen);  // :229
                    cut1 = c1h - L0;
                    const uint32_t ul = popcw<WT>(sh_umask[sample]);
                    taglen = ul ? 5 + ul : 0;
                    uint32_t body = Lend - L1;
                    mode = B_VERBATIM;
                    bool fine = true;
                    if (p.fused_trim >= 0) {
                        if (!(L1 > L0 && win[L1 - 1] == '\n')) {
                            report_err(st, rec, K_TRUNC_FUSED);
                            fine = false;
                        } else if (!plan_trim_body(W, L1, L2, L3, L4, p.fused_trim, mode, kk, body)) {
                            report_err(st, rec, K_SEQ_SHORT);
                            fine = false;
                        }
                    }
                    if (fine) outlen = alen + blen + taglen + 1 + body;
                    else mode = B_NONE;
                }
            }
            r_outlen[r] = outlen;
            r_ext[r] = ext;
            r_alen[r] = (uint16_t)alen;
            r_cut1[r] = (uint16_t)cut1;
            r_blen[r] = (uint16_t)blen;
            r_k[r] = (uint16_t)kk;
            r_taglen[r] = (uint16_t)taglen;
            r_sample[r] = (int16_t)sample;
            r_mode[r] = mode;
            my_outlen = outlen;
        }

        if (HAS_OUT) {
            // ---- P6 layout of the chunk's output
            uint32_t chunk_out = 0;
            if (ORDERED) {
                const uint32_t off = block_excl_scan<NT>(my_outlen, M->scratch, chunk_out);
                if (tid < (int)nrec) r_outoff[tid] = off;
            } else {
                for (uint32_t s = tid; s < S; s += NT) hist[s] = 0;
                __syncthreads();
                if (tid < (int)nrec && r_sample[tid] >= 0 && my_outlen) atomicAdd(&hist[r_sample[tid]], my_outlen);
                __syncthreads();
                // sample-major bases: exclusive scan of hist over S (tiles of NT with a carry)
                uint32_t carry = 0;
                for (uint32_t s0 = 0; s0 < S; s0 += NT) {
                    const uint32_t s = s0 + tid;
                    const uint32_t v = s < S ? hist[s] : 0;
                    uint32_t t2;
                    const uint32_t e = block_excl_scan<NT>(v, M->scratch, t2);
                    if (s < S) sbase[s] = carry + e;
                    carry += t2;
                }
                chunk_out = carry;
                __syncthreads();
                // stable order inside a sample: bytes of earlier records of this chunk with the same sample
                if (tid < (int)nrec) {
                    const int sm = r_sample[tid];
                    uint32_t off = 0;
                    if (sm >= 0 && my_outlen) {
                        for (int i = 0; i < tid; i++)
                            if (r_sample[i] == sm) off += r_outlen[i];
                        off += sbase[sm];
                    }
                    r_outoff[tid] = off;
                }
                // slice table row (u16 lengths)
                if (p.out) {
                    for (uint32_t s = tid; s < S; s += NT) {
                        const uint32_t h = hist[s];
                        if (h > 0xFFFFu) report_err(st, rec0, K_OUT_OVERFLOW);
                        p.lens[(uint64_t)c * S + s] = (uint16_t)h;
                    }
                }
            }

            // ---- P7 reserve output space
            if (ORDERED) {
                if (warp == 0) {
                    const uint64_t excl = lookback(p.tile_out, c, chunk_out, lane);
                    if (lane == 0) {
                        M->out_base = excl;
                        if (c == p.n_chunks - 1) {
                            st->out_bytes = excl + chunk_out;
                            st->out_extent = excl + chunk_out;
                        }
                    }
                }
            } else if (tid == 0) {
                unsigned long long base = 0;
                if (p.out) {
                    base = atomicAdd(&st->out_cursor, (unsigned long long)((chunk_out + 15u) & ~15u));
                    p.chunk_base[c] = base;
                    if (chunk_out) atomicAdd(&st->out_bytes, (unsigned long long)chunk_out);
                }
                M->out_base = base;
            }
            __syncthreads();
            const uint64_t out_base = M->out_base;
            bool writable = p.out != nullptr && chunk_out > 0;
            if (writable && out_base + chunk_out > p.out_cap) {
                if (tid == 0) report_err(st, rec0, K_OUT_OVERFLOW);
                writable = false;
            }

            if (writable) {
                // ---- P8 assemble: one warp per record.  Fast path builds the chunk's output image in
                // shared memory (aligned like its global destination); oversized outputs go straight to
                // global memory byte by byte.
                const uint32_t shift = (uint32_t)(out_base & 15u);
                const bool staged = chunk_out <= (uint32_t)Cfg::STAGE;
                uint8_t *dst0 = staged ? stage + shift : p.out + out_base;
                for (uint32_t r = warp; r < nrec; r += NW) {
                    const uint32_t outlen = r_outlen[r];
                    if (!outlen) continue;
                    const uint32_t j = j0 + r * lpr;
                    const uint32_t L0 = lb(W, j), L1 = lb(W, j + 1), L2 = lb(W, j + 2);
                    const uint32_t L3 = lpr == 4 ? lb(W, j + 3) : L2, L4 = lpr == 4 ? lb(W, j + 4) : L2;
                    const uint32_t Lend = lpr == 4 ? L4 : L2;
                    uint8_t *d = dst0 + r_outoff[r];
                    const uint32_t kk = r_k[r];
                    const uint8_t mode = r_mode[r];
                    if (OP == OP_TRIM || OP == OP_MASK) {
                        wcopy(d, win + L0, L1 - L0, lane);
                        d += L1 - L0;
                    } else {
                        const uint32_t alen = r_alen[r], blen = r_blen[r], taglen = r_taglen[r];
                        wcopy(d, win + L0, alen, lane);
                        d += alen;
                        wcopy(d, win + L0 + r_cut1[r], blen, lane);
                        d += blen;
                        if (taglen) {
                            if (OP == OP_ADDBC) {
                                if (lane < 4) d[lane] = (uint8_t)" BC:"[lane];
                                wcopy(d + 4, p.ext_data[0] + r_ext[r], taglen - 4, lane);
                            } else {
                                if (lane < 5) d[lane] = (uint8_t)" UMI:"[lane];
                                wcopy(d + 5, p.umi + (rec0 + r) * p.sheet.Umax, taglen - 5, lane);
                            }
                            d += taglen;
                        }
                        if (lane == 0) d[0] = '\n';
                        d += 1;
                    }
                    if (mode == B_VERBATIM) {
                        wcopy(d, win + L1, Lend - L1, lane);
                    } else if (mode == B_TRIM) {
                        wcopy(d, win + L1, kk, lane);
                        d += kk;
                        if (lane < 3) d[lane] = (lane == 1) ? '+' : '\n';
                        d += 3;
                        wcopy(d, win + L3, kk, lane);
                        if (lane == 0) d[kk] = '\n';
                    } else if (mode == B_GARBAGE) {
                        if (lane < 6) d[lane] = (uint8_t)"N\n+\n!\n"[lane];
                    } else if (mode == B_MASK) {
                        const uint32_t minq = p.min_baseq;
                        for (uint32_t i = lane; i < kk; i += 32) {
                            const uint8_t q = (uint8_t)(win[L3 + i] - 33u);  // fasta_mask_by_quality.rs:42
                            d[i] = q < minq ? (uint8_t)'N' : win[L1 + i];
                        }
                        d += kk;
                        if (lane < 3) d[lane] = (lane == 1) ? '+' : '\n';
                        d += 3;
                        wcopy(d, win + L3, kk, lane);
                        if (lane == 0) d[kk] = '\n';
                    }
                }
                __syncthreads();
                // ---- P9 store the staged image with 16-byte vectors (bytes at an unaligned head/tail)
                if (staged) {
                    const uint32_t span = shift + chunk_out;
                    uint8_t *g16 = p.out + (out_base - shift);
                    for (uint32_t v = tid; v * 16 < span; v += NT) {
                        const uint32_t b0 = v * 16;
                        if (b0 >= shift && b0 + 16 <= span) {
                            *(uint4 *)(g16 + b0) = *(const uint4 *)(stage + b0);
                        } else {
                            const uint32_t lo = b0 < shift ? shift : b0, hi = b0 + 16 < span ? b0 + 16 : span;
                            for (uint32_t b2 = lo; b2 < hi; b2++) g16[b2] = stage[b2];
                        }
                    }
                }
            }
        }

        // ---- P10 per-record side tables
        if (OP == OP_DEMUX1 && tid < (int)nrec) p.assign[rec0 + tid] = r_sample[tid];
        __syncthreads();  // window, staging and record arrays are reused by the next chunk
    }

    // ---- flush per-CTA counters (fasta_demultiplex.rs:108-109,169,177-178)
    if (OP == OP_DEMUX1) {
        __syncthreads();
        for (uint32_t s = tid; s < S; s += NT)
            if (ccount[s]) atomicAdd(&p.counts[s], (unsigned long long)ccount[s]);
        if (tid == 0) {
            if (M->cta_total) atomicAdd(&p.counts[S], (unsigned long long)M->cta_total);
            if (M->cta_ident) atomicAdd(&p.counts[S + 1], (unsigned long long)M->cta_ident);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
int chunk_kernel_smem_bytes(uint32_t S, uint32_t wide) { return (int)smem_layout<CfgStd>(S, wide).total; }

template <int OP, typename WT>
static int launch_one(const KParams &p, int sm_count, cudaStream_t stream, const char **err) {
    auto kfn = sk_chunk_kernel<CfgStd, OP, WT>;
    const bool demux = (OP == OP_DEMUX1 || OP == OP_DEMUX2);
    const int smem = (int)smem_layout<CfgStd>(demux ? p.sheet.S : 0u, sizeof(WT) == 8).total;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, CfgStd::NT, smem);
    if (e != cudaSuccess || per_sm < 1) { *err = e != cudaSuccess ? cudaGetErrorString(e) : "kernel does not fit on an SM"; return -1; }
    long long grid = (long long)sm_count * per_sm;  // persistent: every CTA is resident (look-back needs it)
    if (grid > (long long)p.n_chunks) grid = p.n_chunks;
    if (grid < 1) return 0;
    kfn<<<(unsigned)grid, CfgStd::NT, smem, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 1;
}

int launch_chunk_kernel(int op, const KParams &p, int sm_count, void *stream_, const char **err) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool wide = p.sheet.wide != 0;
    switch (op) {
        case OP_SCAN: return launch_one<OP_SCAN, uint32_t>(p, sm_count, stream, err);
        case OP_TRIM: return launch_one<OP_TRIM, uint32_t>(p, sm_count, stream, err);
        case OP_MASK: return launch_one<OP_MASK, uint32_t>(p, sm_count, stream, err);
        case OP_ADDBC: return launch_one<OP_ADDBC, uint32_t>(p, sm_count, stream, err);
        case OP_DEMUX1:
            return wide ? launch_one<OP_DEMUX1, uint64_t>(p, sm_count, stream, err)
                        : launch_one<OP_DEMUX1, uint32_t>(p, sm_count, stream, err);
        case OP_DEMUX2:
            return wide ? launch_one<OP_DEMUX2, uint64_t>(p, sm_count, stream, err)
                        : launch_one<OP_DEMUX2, uint32_t>(p, sm_count, stream, err);
    }
    *err = "unknown operator";
    return -1;
}

}  // namespace sk
