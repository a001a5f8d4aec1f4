// sk_device.cuh -- device-side building blocks shared by the engines (sk_kernels.cu, sk_warp.cu):
// PTX wrappers (mbarrier, TMA bulk copies, look-back words), SWAR byte tests and the per-record
// restatements of the reference logic (trim scan, " BC:" search, barcode match, header surgery).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sk_internal.h"

namespace sk {

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
// TMA 1-D bulk copy shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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
static __device__ __noinline__ uint64_t lookback(uint64_t *tiles, uint32_t c, uint64_t agg, int lane) {
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
            w = ts_load(&tiles[idx]);
            while ((w >> 62) == TS_INVALID) {
                __nanosleep(40);
                w = ts_load(&tiles[idx]);
            }
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

// Decoupled look-back in two halves, with a 256-entry window per round (every lane reads eight
// consecutive words with four 16-byte loads).  lookback_publish makes this chunk's aggregate visible as
// early as possible; lookback_consume, called by one whole warp whenever the prefix is needed, returns
// the exclusive prefix over all earlier chunks and publishes the inclusive one.  Chunks are handed out
// by a ticket counter, so every predecessor is owned by a resident CTA that never waits on a later chunk.
__device__ __forceinline__ void lookback_publish(uint64_t *tiles, uint32_t c, uint64_t agg) {
    ts_store(&tiles[c], c == 0 ? TS_INC : TS_AGG, agg);
}
static __device__ __noinline__ uint64_t lookback_consume(uint64_t *tiles, uint32_t c, uint64_t agg, int lane) {
    if (c == 0) return 0;
    constexpr int E = 8;  // entries per lane and round (256-entry window)
    uint64_t excl = 0;
    int64_t hi = (int64_t)c - 1;  // nearest predecessor not yet accounted for
    for (;;) {
        // lane L takes the aligned block of E entries B = hi / E - L; entries above hi are skipped
        const int64_t blk = (hi / E) - lane;
        uint64_t lsum = 0;
        bool lhas = false;
        if (blk < 0) {
            lhas = true;  // before chunk 0: inclusive prefix 0
        } else {
            const uint64_t *q = tiles + blk * E;
            uint64_t w[E];
            for (;;) {
#pragma unroll
                for (int k = 0; k < E; k += 2)
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[k]), "=l"(w[k + 1]) : "l"(q + k) : "memory");
                bool wait = false;
#pragma unroll
                for (int k = E - 1; k >= 0; k--) {
                    if (blk * E + k > hi) continue;
                    if ((w[k] >> 62) == TS_INVALID) wait = true;
                    if ((w[k] >> 62) == TS_INC) break;  // entries below an inclusive one are not needed
                }
                if (!wait) break;
                __nanosleep(40);
            }
#pragma unroll
            for (int k = E - 1; k >= 0; k--) {
                if (blk * E + k > hi || lhas) continue;
                lsum += w[k] & TS_VMASK;
                if ((w[k] >> 62) == TS_INC) lhas = true;
            }
        }
        const uint32_t inc = __ballot_sync(0xffffffffu, lhas);
        if (inc) {
            const int first = __ffs(inc) - 1;
            excl += warp_sum64(lane <= first ? lsum : 0);
            break;
        }
        excl += warp_sum64(lsum);
        hi = ((hi / E) - 32) * E + (E - 1);  // the block below the window, whole
    }
    if (lane == 0) ts_store(&tiles[c], TS_INC, excl + agg);
    return excl;
}
static __device__ __forceinline__ uint64_t lookback_wide(uint64_t *tiles, uint32_t c, uint64_t agg, int lane) {
    if (lane == 0) lookback_publish(tiles, c, agg);
    return lookback_consume(tiles, c, agg, lane);
}

// Exclusive block scan of one u32 per thread; `scratch` holds 2*(NT/32) words (double buffered by
// `flip`, so that one barrier per call suffices).
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *scratch, uint32_t &flip, uint32_t &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    uint32_t *s = scratch + flip * NW;
    flip ^= 1u;
    if (lane == 31) s[w] = x;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t t = s[k];
        if (k < w) before += t;
        all += t;
    }
    total = all;
    return before + x - v;
}

// 0x80 in every byte of x that equals the byte replicated in `pat` (exact SWAR test, no carries).
__device__ __forceinline__ uint32_t eq_flags(uint32_t x, uint32_t pat) {
    const uint32_t t = x ^ pat;
    return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
}
// number of non-zero bytes of x
__device__ __forceinline__ uint32_t nz_bytes(uint32_t x) {
    return __popc((((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u);
}
// Newline map of a 16-byte piece: bit (8*b + w) is set when byte b of word w is '\n'
// (window offset of that byte inside the piece = 4*w + b).
__device__ __forceinline__ uint32_t nl_map(const uint4 v) {
    const uint32_t zx = eq_flags(v.x, 0x0A0A0A0Au), zy = eq_flags(v.y, 0x0A0A0A0Au);
    const uint32_t zz = eq_flags(v.z, 0x0A0A0A0Au), zw = eq_flags(v.w, 0x0A0A0A0Au);
    return (zx >> 7) | (zy >> 6) | (zz >> 5) | (zw >> 4);
}
__device__ __forceinline__ uint32_t map_bit(uint32_t k) { return 8u * (k & 3u) + (k >> 2); }  // piece offset k -> map bit
// the map restricted to piece offsets k with lo <= o + k < hi (rare path: pieces at a range boundary)
__device__ __forceinline__ uint32_t map_clip(uint32_t y, uint32_t o, uint32_t lo, uint32_t hi) {
    uint32_t keep = 0;
    for (uint32_t k = 0; k < 16; k++)
        if (o + k >= lo && o + k < hi) keep |= 1u << map_bit(k);
    return y & keep;
}

__device__ __forceinline__ bool is_ws(uint8_t c) { return c == 32u || (c >= 9u && c <= 13u); }  // ASCII White_Space

__device__ __forceinline__ void report_err(DevStats *st, uint64_t rec, unsigned kind) {
    atomicMax(&st->err_key, ~((rec << 8) | (unsigned long long)kind));
}

// body modes of a planned record
enum : uint8_t { B_VERBATIM = 0, B_TRIM = 1, B_GARBAGE = 2, B_MASK = 3, B_NONE = 4, B_FAIL = 5 };
enum : uint8_t { RF_SLOW = 1, RF_DEAD = 2 };

// fasta_trim_by_quality.rs:28-48 on the quality line [L3,L4) / sequence line [L1,L2), one thread.
// Whole aligned words are examined with four dot products (the running totals after each of the
// word's bytes); bytes below '!' (wrapping u8 subtraction, :35) and the word in which the loop
// breaks (:37) fall back to the byte-wise form.  Returns false when &seq[..k] would panic (:47).
__device__ __forceinline__ bool plan_trim_body(const uint8_t *b, uint32_t L1, uint32_t L2, uint32_t L3, uint32_t L4,
                                               int minq, uint8_t &mode, uint32_t &kk, uint32_t &body_len) {
    uint32_t k = L4 - L3;
    while (k > 0 && is_ws(b[L3 + k - 1])) k--;  // qual.trim_end().len()  (:31)
    int total = -50, lowest = -50;              // :28-29
    uint32_t lowest_k = k;
    uint32_t pos = L3 + k;  // one past the byte examined next
    const int sub = 33 + minq;
#define SK_TRIM_BYTE(addr)                                                                               \
    {                                                                                                    \
        const uint32_t q = b[addr];                                                                      \
        total += q >= 33u ? (int)q - sub : (int)((q - 33u) & 0xFFu) - minq; /* wrapping u8 subtraction */ \
        if (total > 0) goto trim_done;                                                                   \
        if (total < lowest) {                                                                            \
            lowest = total;                                                                              \
            lowest_k = (addr) - L3;                                                                      \
        }                                                                                                \
    }
    while (pos > L3 && (pos & 3u)) {  // down to a word boundary
        pos--;
        SK_TRIM_BYTE(pos)
    }
    while (pos >= L3 + 4) {  // :33-42, four bytes per step
        const uint32_t w = *(const uint32_t *)(b + pos - 4);
        if ((((w | 0x80808080u) - 0x21212121u) & 0x80808080u) != 0x80808080u || (w & 0x80808080u)) {
            // a byte below '!' (or non-ASCII): byte-wise
            SK_TRIM_BYTE(pos - 1)
            SK_TRIM_BYTE(pos - 2)
            SK_TRIM_BYTE(pos - 3)
            SK_TRIM_BYTE(pos - 4)
            pos -= 4;
            continue;
        }
        const int T3 = (int)__dp4a(w, 0x01000000u, (uint32_t)(total - sub));
        const int T2 = (int)__dp4a(w, 0x01010000u, (uint32_t)(total - 2 * sub));
        const int T1 = (int)__dp4a(w, 0x01010100u, (uint32_t)(total - 3 * sub));
        const int T0 = (int)__dp4a(w, 0x01010101u, (uint32_t)(total - 4 * sub));
        if (max(max(T3, T2), max(T1, T0)) > 0) {  // the break (:37) is inside this word
            if (T3 > 0) goto trim_done;
            if (T3 < lowest) { lowest = T3; lowest_k = pos - 1 - L3; }
            if (T2 > 0) goto trim_done;
            if (T2 < lowest) { lowest = T2; lowest_k = pos - 2 - L3; }
            if (T1 > 0) goto trim_done;
            if (T1 < lowest) { lowest = T1; lowest_k = pos - 3 - L3; }
            goto trim_done;
        }
        const int mn = min(min(T3, T2), min(T1, T0));
        if (mn < lowest) {  // strict '<' (:38): the byte examined first (highest address) wins a tie
            lowest = mn;
            lowest_k = (T3 == mn ? pos - 1 : T2 == mn ? pos - 2 : T1 == mn ? pos - 3 : pos - 4) - L3;
        }
        total = T0;
        pos -= 4;
    }
    while (pos > L3) {
        pos--;
        SK_TRIM_BYTE(pos)
    }
#undef SK_TRIM_BYTE
trim_done:
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

// Leftmost match of " BC:[ACGTNacgtn+]" in [h0,h1): aligned words are tested for ' ' with SWAR and
// only the candidates are looked at byte-wise.  `lut` bit 3 = regex class.  Returns the offset of ' '.
__device__ __forceinline__ bool bc_find(const uint8_t *b, const uint8_t *lut, uint32_t h0, uint32_t h1, uint32_t &st) {
    if (h1 < h0 + 5) return false;
    const uint32_t last = h1 - 5;  // last admissible start
    for (uint32_t a = h0 & ~3u; a <= last; a += 4) {
        uint32_t z = eq_flags(*(const uint32_t *)(b + a), 0x20202020u);
        while (z) {
            const uint32_t i = a + ((__ffs(z) - 1) >> 3);
            z &= z - 1;
            if (i >= h0 && i <= last && b[i + 1] == 'B' && b[i + 2] == 'C' && b[i + 3] == ':' && (lut[b[i + 4]] & 8u)) {
                st = i;
                return true;
            }
        }
    }
    return false;
}
// end of the greedy class run that starts at `from` (from <= h1)
__device__ __forceinline__ uint32_t bc_run_end(const uint8_t *b, const uint8_t *lut, uint32_t from, uint32_t h1) {
    uint32_t e = from;
    while (e < h1 && (e & 3u)) {  // up to the next aligned word
        if (!(lut[b[e]] & 8u)) return e;
        e++;
    }
    while (e + 4 <= h1) {
        const uint32_t w = *(const uint32_t *)(b + e);
        const uint32_t c0 = lut[w & 0xFFu], c1 = lut[(w >> 8) & 0xFFu], c2 = lut[(w >> 16) & 0xFFu], c3 = lut[w >> 24];
        if (!(c0 & 8u)) return e;
        if (!(c1 & 8u)) return e + 1;
        if (!(c2 & 8u)) return e + 2;
        if (!(c3 & 8u)) return e + 3;
        e += 4;
    }
    while (e < h1 && (lut[b[e]] & 8u)) e++;
    return e;
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
template <typename WT>
__device__ __forceinline__ uint32_t ffsw(WT x);  // index of the lowest set bit
template <>
__device__ __forceinline__ uint32_t ffsw<uint32_t>(uint32_t x) { return (uint32_t)__ffs((int)x) - 1u; }
template <>
__device__ __forceinline__ uint32_t ffsw<uint64_t>(uint64_t x) { return (uint32_t)__ffsll((long long)x) - 1u; }

// Pigeonhole barcode match (HalfIdx, sk_internal.h): (lowest distance, first and last sample at that
// distance) over every sample within one mismatch of the barcode at window offset bs -- the outcome
// of fasta_demultiplex.rs:157-166 whenever it matters (:172).  lowest stays 0xFFFFFFFF when no
// sample is within one mismatch.
template <int NWMAX>
__device__ __forceinline__ void hidx_match(const uint8_t *b, uint32_t bs, const HalfIdx &H, const uint32_t *hcls,
                                           uint32_t &lowest, uint32_t &best, uint32_t &last) {
    const uint32_t a = bs & ~3u, sh = (bs & 3u) * 8u;
    const uint32_t nw = H.nw, nwp = H.nwp, tmask = H.tsize - 1u;
    uint32_t raw[NWMAX];
    {
        uint32_t lo = *(const uint32_t *)(b + a);
#pragma unroll
        for (int w = 0; w < NWMAX; w++) {
            raw[w] = 0;
            if (w < (int)nw) {
                const uint32_t hi = *(const uint32_t *)(b + a + 4 * w + 4);
                raw[w] = __funnelshift_r(lo, hi, sh);
                lo = hi;
            }
        }
    }
    lowest = 0xFFFFFFFFu;
    best = 0xFFFFFFFFu;
    last = 0;
    for (uint32_t c = 0; c < H.n_classes; c++) {
        const uint32_t *care = hcls + c * HIDX_CLS_ROWS * nwp;
        for (uint32_t h = 0; h < 2; h++) {
            const uint32_t *hm = care + (1 + 3 * h) * nwp;
            uint32_t h1 = 0, h2 = 0;
#pragma unroll
            for (int w = 0; w < NWMAX; w++)
                if (w < (int)nw) {
                    const uint32_t k = raw[w] & hm[w];
                    h1 += k * hm[nwp + w];
                    h2 += k * hm[2 * nwp + w];
                }
            h1 ^= h1 >> 15;
            uint32_t slot = h1 & tmask;
            const uint2 *tab = H.table + (size_t)(c * 2 + h) * H.tsize;
            for (;;) {
                const uint2 e = __ldg(&tab[slot]);
                const uint32_t cnt = e.y >> 16;
                if (!cnt) break;  // empty slot: end of the probe chain
                if (e.x == h2) {
                    const uint32_t st = e.y & 0xFFFFu;
                    for (uint32_t i = 0; i < cnt; i++) {
                        const uint32_t s = __ldg(&H.cand[st + i]);
                        const uint4 *sk = (const uint4 *)(H.skeys + (size_t)s * nwp);
                        uint32_t d = 0;
#pragma unroll
                        for (int q = 0; q < NWMAX / 4; q++)
                            if (4 * q < (int)nw) {
                                const uint4 kq = __ldg(&sk[q]);
                                d += nz_bytes((raw[4 * q] & care[4 * q]) ^ kq.x);
                                d += nz_bytes((raw[4 * q + 1] & care[4 * q + 1]) ^ kq.y);
                                d += nz_bytes((raw[4 * q + 2] & care[4 * q + 2]) ^ kq.z);
                                d += nz_bytes((raw[4 * q + 3] & care[4 * q + 3]) ^ kq.w);
                            }
                        if (d < lowest) {
                            lowest = d;
                            best = s;
                            last = s;
                        } else if (d == lowest) {
                            best = s < best ? s : best;
                            last = s > last ? s : last;
                        }
                    }
                }
                slot = (slot + 1) & tmask;
            }
        }
    }
    if (lowest > 1u) lowest = 0xFFFFFFFFu;  // farther samples were not enumerated completely
}

// Thread-serial copy of `len` bytes inside shared memory.  The destination is brought to 4- and then
// 16-byte alignment, the body moves 16 bytes per iteration (4 LDS.32 + 4 funnel shifts + 1 STS.128);
// the source is only ever read as aligned words.
__device__ __forceinline__ void tcopy(uint8_t *dst, const uint8_t *src, uint32_t len) {
    while (len && ((uint32_t)(uintptr_t)dst & 3u)) {
        *dst++ = *src++;
        len--;
    }
    if (len >= 4) {
        const uint32_t sh = ((uint32_t)(uintptr_t)src & 3u) * 8u;
        const uint32_t *sw = (const uint32_t *)((uintptr_t)src & ~(uintptr_t)3);
        uint32_t *dw = (uint32_t *)dst;
        uint32_t nwords = len >> 2;
        uint32_t lo = *sw++;
        while (nwords && ((uint32_t)(uintptr_t)dw & 15u)) {
            const uint32_t hi = *sw++;
            *dw++ = __funnelshift_r(lo, hi, sh);
            lo = hi;
            nwords--;
        }
        for (; nwords >= 4; nwords -= 4) {
            const uint32_t w1 = sw[0], w2 = sw[1], w3 = sw[2], w4 = sw[3];
            uint4 o;
            o.x = __funnelshift_r(lo, w1, sh);
            o.y = __funnelshift_r(w1, w2, sh);
            o.z = __funnelshift_r(w2, w3, sh);
            o.w = __funnelshift_r(w3, w4, sh);
            *(uint4 *)dw = o;
            lo = w4;
            sw += 4;
            dw += 4;
        }
        while (nwords) {
            const uint32_t hi = *sw++;
            *dw++ = __funnelshift_r(lo, hi, sh);
            lo = hi;
            nwords--;
        }
        const uint32_t done = len & ~3u;
        dst += done;
        src += done;
        len &= 3u;
    }
    while (len) {
        *dst++ = *src++;
        len--;
    }
}
// byte-wise variant for rare paths (global sources / unstaged global destination)
__device__ __forceinline__ void bcopy(uint8_t *dst, const uint8_t *src, uint32_t len) {
    for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
}

}  // namespace sk
