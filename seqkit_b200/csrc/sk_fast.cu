// sk_fast.cu -- the lean chunk engine for the hot operators of the path:
//   OP_TRIM / OP_MASK / OP_DEMUX1 / OP_DEMUX2 (header route, optional fused trim) on 4-line FASTQ.
//
// Same single-pass idea as sk_kernels.cu (input read once by TMA, output written once by TMA, record
// framing by global line index through a decoupled look-back), re-cut for latency hiding and
// instruction count:
//   * small chunks (8 KiB) and small CTAs (4 warps, ~26 KiB of shared memory) -> 8 CTAs per SM that
//     sit in different phases at any moment, so barrier and look-back latency of one CTA is covered
//     by the others;
//   * chunk and thread boundaries coincide (chunk = 102 threads x 80 bytes, no bytes before the
//     chunk): the newline scan has no range clipping except in the last window of a stream;
//   * newline maps are built in natural bit order with dp4a, so the line table is a plain ffs loop;
//   * per-record logic runs one thread per record with both table loads of the barcode index issued
//     before either is used; quality trim and header search/match run on different warps;
//   * a chunk's demultiplexed output keeps input order (one slice-table group per emitted record);
//   * anything outside this engine's limits (record longer than the overhang, > NT records per
//     chunk, output larger than the staging image) raises F_NEED_GENERAL and the operator is re-run
//     on the general engine (sk_kernels.cu) by sk_wait -- still CUDA, never a CPU path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sk_internal.h"

namespace sk {
extern __shared__ __align__(128) unsigned char sk_smem[];
}
#include "sk_device.cuh"

namespace sk {

constexpr int FREC_BYTES = 20;  // per-record plan fields (8 x u16, i16, 2 x u8)

template <class G>
struct FLayout {
    static constexpr uint32_t win = 0;
    static constexpr uint32_t stage = G::WIN_MAX;
    static constexpr uint32_t ls = stage + G::STAGE + 32;
    static constexpr uint32_t rec = ls + (((G::MAXLINES + 8) * 2 + 15) / 16) * 16;
    static constexpr uint32_t lut = rec + ((G::MAXREC * FREC_BYTES + 15) / 16) * 16;
    static constexpr uint32_t misc = lut + 256;
    static constexpr uint32_t dyn = misc + 256;  // hcls rows, then the per-sample counters
};

struct FMisc {
    uint64_t mbar;
    uint64_t g0;
    uint64_t out_base;
    uint32_t chunk;
    uint32_t pad;
    uint32_t scratch[2 * 8];
};

// 0x80 in every byte of x that is '\n'.  Exact for 7-bit input; a batch with a byte >= 0x80 is
// refused as a whole (F_NON_ASCII), whatever this returns for it.
__device__ __forceinline__ uint32_t nl_flags7(uint32_t x) {
    const uint32_t t = x ^ 0x0A0A0A0Au;
    return ~(t + 0x7F7F7F7Fu) & 0x80808080u;
}
// newline map of a 16-byte piece in natural order: bit k <=> byte k is '\n'
__device__ __forceinline__ uint32_t nl_map_nat(const uint4 v) {
    const uint32_t zx = nl_flags7(v.x), zy = nl_flags7(v.y), zz = nl_flags7(v.z), zw = nl_flags7(v.w);
    const uint32_t lo = __dp4a(zx, 0x08040201u, __dp4a(zy, 0x80402010u, 0u));  // (bits 0-7) << 7
    const uint32_t hi = __dp4a(zz, 0x08040201u, __dp4a(zw, 0x80402010u, 0u));  // (bits 8-15) << 7
    return (lo >> 7) + hi * 2u;
}
// bits i with 0 <= i < hi (hi may be <= 0 or >= 32)
__device__ __forceinline__ uint32_t bits_below(int hi) {
    return hi <= 0 ? 0u : (hi >= 32 ? 0xFFFFFFFFu : (1u << hi) - 1u);
}

// Exclusive block scan of one u32 per thread (4 or 8 warps): warp scans by shuffle, then every thread
// sums the warp totals out of two 16-byte shared-memory reads.  `scratch` holds 2*NW words (double
// buffered by `flip`, one barrier per call).
template <int NT>
__device__ __forceinline__ uint32_t block_scan_fast(uint32_t v, uint32_t *scratch, uint32_t &flip, uint32_t &total) {
    constexpr int NW = NT / 32;
    static_assert(NW == 4 || NW == 8, "block_scan_fast: 4 or 8 warps");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    uint32_t *s = scratch + flip * NW;
    flip ^= 1u;
    if (lane == 31) s[w] = x;
    __syncthreads();
    const uint4 a = *(const uint4 *)s;
    uint32_t before = (w > 0 ? a.x : 0u) + (w > 1 ? a.y : 0u) + (w > 2 ? a.z : 0u) + (w > 3 ? a.w : 0u);
    uint32_t all = a.x + a.y + a.z + a.w;
    if (NW == 8) {
        const uint4 b = *(const uint4 *)(s + 4);
        before += (w > 4 ? b.x : 0u) + (w > 5 ? b.y : 0u) + (w > 6 ? b.z : 0u);
        all += b.x + b.y + b.z + b.w;
    }
    total = all;
    return before + x - v;
}

// mbarrier wait that parks the thread in hardware for up to the hinted time per attempt
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
#ifdef SK_PHASE_TIMING
#define FK_T(i)                                              \
    do {                                                     \
        if (tid == 0) {                                      \
            const long long t_now = clock64();               \
            ph[i] += (unsigned long long)(t_now - t_prev);   \
            t_prev = t_now;                                  \
        }                                                    \
    } while (0)
#else
#define FK_T(i) do { } while (0)
#endif

// Leftmost match of " BC:[class]" in [h0,h1), 16 bytes per step (see bc_find).
__device__ __forceinline__ bool bc_find16(const uint8_t *b, const uint8_t *lut, uint32_t h0, uint32_t h1, uint32_t &st) {
    if (h1 < h0 + 5) return false;
    const uint32_t last = h1 - 5;
    for (uint32_t a = h0 & ~15u; a <= last; a += 16) {
        const uint4 v = *(const uint4 *)(b + a);
        const uint32_t z0 = eq_flags(v.x, 0x20202020u), z1 = eq_flags(v.y, 0x20202020u);
        const uint32_t z2 = eq_flags(v.z, 0x20202020u), z3 = eq_flags(v.w, 0x20202020u);
        if (!(z0 | z1 | z2 | z3)) continue;
        uint32_t m = __dp4a(z0, 0x08040201u, __dp4a(z1, 0x80402010u, 0u)) >> 7;
        m |= (__dp4a(z2, 0x08040201u, __dp4a(z3, 0x80402010u, 0u)) >> 7) << 8;
        while (m) {
            const uint32_t i = a + (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1;
            if (i >= h0 && i <= last && b[i + 1] == 'B' && b[i + 2] == 'C' && b[i + 3] == ':' && (lut[b[i + 4]] & 8u)) {
                st = i;
                return true;
            }
        }
    }
    return false;
}

// The observed barcode as words aligned to its first byte: raw[w] = bytes [bs+4w, bs+4w+4).
template <int NR>
__device__ __forceinline__ void load_raw(const uint8_t *b, uint32_t bs, uint32_t nwords, uint32_t (&raw)[NR]) {
    const uint32_t a = bs & ~3u, sh = (bs & 3u) * 8u;
    uint32_t lo = *(const uint32_t *)(b + a);
#pragma unroll
    for (int w = 0; w < NR; w++) {
        raw[w] = 0;
        if (w < (int)nwords) {
            const uint32_t hi = *(const uint32_t *)(b + a + 4 * w + 4);
            raw[w] = __funnelshift_r(lo, hi, sh);
            lo = hi;
        }
    }
}
// Is the greedy class run (fasta_demultiplex.rs:38) that starts at the barcode's first byte exactly L
// bytes long?  raw holds ceil((L+1)/4) words; `room` = header bytes from the barcode start to the end
// of the header line.  Every lane runs the same number of steps.
template <int NR>
__device__ __forceinline__ bool class_run_is(const uint32_t (&raw)[NR], const uint8_t *lut, uint32_t L, uint32_t room) {
    if (room < L) return false;
    uint32_t acc = 8u, term = 0;
#pragma unroll
    for (int w = 0; w < NR; w++) {
        const uint32_t x = raw[w];
        if (4u * w + 4u <= L) {
            acc &= lut[x & 0xFFu] & lut[(x >> 8) & 0xFFu] & lut[(x >> 16) & 0xFFu] & lut[x >> 24];
        } else if (4u * w < L) {
            acc &= lut[x & 0xFFu];
            if (4u * w + 1u < L) acc &= lut[(x >> 8) & 0xFFu];
            if (4u * w + 2u < L) acc &= lut[(x >> 16) & 0xFFu];
        }
        if ((uint32_t)w == (L >> 2)) term = (x >> (8u * (L & 3u))) & 0xFFu;
    }
    if (!(acc & 8u)) return false;          // the run ends early
    if (room == L) return true;             // the header line ends with the barcode
    return !(lut[term] & 8u);               // the byte after the barcode ends the run
}

// End of the greedy class run (fasta_demultiplex.rs:38) that starts at `from`, four bytes per step on
// words aligned to `from` (lanes of a warp step together when their barcodes are equally long).
__device__ __forceinline__ uint32_t class_run_end(const uint8_t *b, const uint8_t *lut, uint32_t from, uint32_t h1) {
    const uint32_t a = from & ~3u, sh = (from & 3u) * 8u;
    uint32_t lo = *(const uint32_t *)(b + a);
    uint32_t e = from;
    while (e < h1) {
        const uint32_t hi = *(const uint32_t *)(b + a + 4 + (e - from));
        const uint32_t x = __funnelshift_r(lo, hi, sh);
        lo = hi;
        const uint32_t c0 = lut[x & 0xFFu], c1 = lut[(x >> 8) & 0xFFu], c2 = lut[(x >> 16) & 0xFFu], c3 = lut[x >> 24];
        if (c0 & c1 & c2 & c3 & 8u) {
            e += 4;
            continue;
        }
        e += (c0 & 8u) ? ((c1 & 8u) ? ((c2 & 8u) ? 3u : 2u) : 1u) : 0u;
        break;
    }
    return e < h1 ? e : h1;
}

// Running totals of fasta_trim_by_quality.rs:33-36 over the aligned 8-byte block at window offset a,
// in the order the reference examines the bytes: T[i] is the total after byte a+7-i, starting from
// `total`.  Bytes outside the quality string [L3,E) contribute nothing (they are replaced by the
// byte whose contribution is zero, sub = 33 + min_baseq <= 255).  A byte below '!' takes the wrapping
// u8 subtraction (:35) on the byte-wise path.
__device__ __forceinline__ void blk8_totals(const uint8_t *b, uint32_t a, uint32_t L3, uint32_t E, int sub, int minq,
                                            int total, int (&T)[8]) {
    uint2 v = *(const uint2 *)(b + a);
    if (a < L3 || a + 8 > E) {
        const uint32_t nlow = a < L3 ? L3 - a : 0u, nhigh = a + 8 > E ? a + 8 - E : 0u;
        unsigned long long m = nlow >= 8u ? 0ull : (~0ull << (8u * nlow));
        m = nhigh >= 8u ? 0ull : (m & (~0ull >> (8u * nhigh)));
        const uint32_t mlo = (uint32_t)m, mhi = (uint32_t)(m >> 32), sub4 = (uint32_t)sub * 0x01010101u;
        v.x = (v.x & mlo) | (sub4 & ~mlo);
        v.y = (v.y & mhi) | (sub4 & ~mhi);
    }
    const uint32_t H = 0x80808080u, C = 0x21212121u;
    const uint32_t bad = (~((v.x | H) - C) & ~v.x & H) | (~((v.y | H) - C) & ~v.y & H);  // bytes below '!'
    if (!bad) {
        T[0] = (int)__dp4a(v.y, 0x01000000u, (uint32_t)(total - sub));
        T[1] = (int)__dp4a(v.y, 0x01010000u, (uint32_t)(total - 2 * sub));
        T[2] = (int)__dp4a(v.y, 0x01010100u, (uint32_t)(total - 3 * sub));
        T[3] = (int)__dp4a(v.y, 0x01010101u, (uint32_t)(total - 4 * sub));
        T[4] = (int)__dp4a(v.x, 0x01000000u, (uint32_t)(T[3] - sub));
        T[5] = (int)__dp4a(v.x, 0x01010000u, (uint32_t)(T[3] - 2 * sub));
        T[6] = (int)__dp4a(v.x, 0x01010100u, (uint32_t)(T[3] - 3 * sub));
        T[7] = (int)__dp4a(v.x, 0x01010101u, (uint32_t)(T[3] - 4 * sub));
    } else {
        int t = total;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t q = ((i < 4 ? v.y : v.x) >> (8 * (3 - (i & 3)))) & 0xFFu;
            t += q >= 33u ? (int)q - sub : (int)((q - 33u) & 0xFFu) - minq;
            T[i] = t;
        }
    }
}

// fasta_trim_by_quality.rs:28-48 for one record per lane, all lanes of the warp in step: every lane
// walks its quality string down in aligned 8-byte blocks inside one warp-synchronous loop and only
// notes (a) the block in which the running total first exceeds 0 (:37) and (b) the block holding the
// minimum so far (:38); the positions inside those two blocks are resolved once after the loop.
// Must be called by all 32 lanes (`has` = this lane carries a record).  min_baseq <= 222.
__device__ __forceinline__ bool plan_trim_warp(const uint8_t *b, bool has, uint32_t L1, uint32_t L2, uint32_t L3, uint32_t L4,
                                               int minq, uint8_t &mode, uint32_t &kk, uint32_t &body_len) {
    const uint32_t NONE = 0xFFFFFFFFu;
    uint32_t k = has ? L4 - L3 : 0u;
    while (k > 0 && is_ws(b[L3 + k - 1])) k--;  // qual.trim_end().len()  (:31)
    __syncwarp();
    const uint32_t E = L3 + k;
    const int sub = 33 + minq;
    int total = -50, lowest = -50;  // :28-29
    uint32_t low_a = NONE, brk_a = NONE;
    int low_total = 0, brk_total = 0;
    uint32_t a = k ? ((E - 1u) & ~7u) : 0u;
    bool active = k > 0;
    while (__any_sync(0xffffffffu, active)) {
        if (active) {
            int T[8];
            blk8_totals(b, a, L3, E, sub, minq, total, T);
            const int mx = max(max(max(T[0], T[1]), max(T[2], T[3])), max(max(T[4], T[5]), max(T[6], T[7])));
            const int mn = min(min(min(T[0], T[1]), min(T[2], T[3])), min(min(T[4], T[5]), min(T[6], T[7])));
            if (mx > 0) {  // the break is inside this block
                brk_a = a;
                brk_total = total;
                active = false;
            } else {
                if (mn < lowest) {  // strict '<': an earlier block keeps a tie
                    lowest = mn;
                    low_a = a;
                    low_total = total;
                }
                total = T[7];
                if (a <= L3) active = false;
                else a -= 8;
            }
        }
    }
    // resolve positions: first the break block (its totals before the break may lower the minimum),
    // then the block that holds the minimum
    uint32_t lowest_k = k;
    bool placed = false;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const uint32_t ra = pass == 0 ? brk_a : low_a;
        const bool go = ra != NONE && !placed;
        if (go) {
            int T[8];
            blk8_totals(b, ra, L3, E, sub, minq, pass == 0 ? brk_total : low_total, T);
            bool ok = true;
            int best = 0x7FFFFFFF;
            uint32_t at = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                ok = ok && T[i] <= 0;           // totals after the break are never looked at
                if (ok && T[i] < best) {        // first (highest address) of equal totals wins
                    best = T[i];
                    at = ra + 7u - (uint32_t)i;
                }
            }
            if (pass == 0 ? best < lowest : best == lowest) {
                lowest = best;
                lowest_k = at - L3;
                placed = true;
            }
        }
        __syncwarp();
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

// plan_trim_warp with two adjacent lanes per record (sub = 0 / 1): the pair walks the quality string
// down sixteen bytes per step, lane 0 on the upper 8-byte block and lane 1 on the one below it; each
// lane computes its block's totals relative to 0, one shuffle gives lane 1 the sum of lane 0's block
// (its entering total), another tells the pair whether either block contains the break (:37).  Each
// lane keeps the minimum over its own blocks; the pair's minimum is the lower of the two, the block
// examined first winning a tie (:38).  The positions are resolved after the loop, the break block by
// lane 0 and the minimum block by lane 1 at the same time.  Halves the length of the serial chain.
// Must be called by all 32 lanes; both lanes of a pair pass the same arguments.  min_baseq <= 222.
__device__ __forceinline__ bool plan_trim_pair(const uint8_t *b, bool has, uint32_t sub, uint32_t L1, uint32_t L2,
                                               uint32_t L3, uint32_t L4, int minq, uint8_t &mode, uint32_t &kk,
                                               uint32_t &body_len) {
    const uint32_t FULL = 0xffffffffu;
    const int NONE = -1;
    uint32_t k = has ? L4 - L3 : 0u;
    while (k > 0 && is_ws(b[L3 + k - 1])) k--;  // qual.trim_end().len()  (:31)
    __syncwarp();
    const uint32_t E = L3 + k;
    const int sq = 33 + minq;
    int tot = -50, lowest = -50;  // :28-29; tot = total entering lane 0's block of this step
    int low_a = NONE, low_total = 0, brk_a = NONE, brk_total = 0;
    int a = (k ? (int)((E - 1u) & ~7u) : 0) - 8 * (int)sub;
    bool active = k > 0;
    while (__any_sync(FULL, active)) {
        const bool mine = active && a >= 0 && a + 8 > (int)L3;
        int S = 0, mx = -0x40000000, mn = 0x40000000;
        if (mine) {
            int T[8];
            blk8_totals(b, (uint32_t)a, L3, E, sq, minq, 0, T);
            S = T[7];
            mx = max(max(max(T[0], T[1]), max(T[2], T[3])), max(max(T[4], T[5]), max(T[6], T[7])));
            mn = min(min(min(T[0], T[1]), min(T[2], T[3])), min(min(T[4], T[5]), min(T[6], T[7])));
        }
        const int S_o = __shfl_xor_sync(FULL, S, 1);
        const int enter = tot + (sub ? S_o : 0);
        const bool brk_me = mine && enter + mx > 0;
        const bool brk_o = __shfl_xor_sync(FULL, (int)brk_me, 1) != 0;
        const bool reached = mine && !(sub && brk_o);  // lane 0's block comes first
        if (reached) {
            if (brk_me) {
                brk_a = a;
                brk_total = enter;
            } else if (enter + mn < lowest) {
                lowest = enter + mn;
                low_a = a;
                low_total = enter;
            }
        }
        tot += S + S_o;
        const int a_first = a + 8 * (int)sub;  // lane 0's block of this step
        active = active && !(brk_me || brk_o) && a_first - 8 > (int)L3;
        a -= 16;
    }
    // the pair's break block (lane 0's if it broke, else lane 1's) and minimum block
    {
        const int ba_o = __shfl_xor_sync(FULL, brk_a, 1), bt_o = __shfl_xor_sync(FULL, brk_total, 1);
        const int b0 = sub ? ba_o : brk_a, b1 = sub ? brk_a : ba_o;
        const int t0 = sub ? bt_o : brk_total, t1 = sub ? brk_total : bt_o;
        brk_a = b0 != NONE ? b0 : b1;
        brk_total = b0 != NONE ? t0 : t1;
        const int lo_o = __shfl_xor_sync(FULL, lowest, 1), la_o = __shfl_xor_sync(FULL, low_a, 1);
        const int lt_o = __shfl_xor_sync(FULL, low_total, 1);
        if (la_o != NONE && (lo_o < lowest || (lo_o == lowest && (low_a == NONE || la_o > low_a)))) {
            lowest = lo_o;
            low_a = la_o;
            low_total = lt_o;
        }
    }
    // resolve: lane 0 looks into the break block, lane 1 into the minimum block
    const int ra = sub ? low_a : brk_a;
    int best = 0x7FFFFFFF, at = 0;
    if (ra != NONE) {
        int T[8];
        blk8_totals(b, (uint32_t)ra, L3, E, sq, minq, sub ? low_total : brk_total, T);
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ok = ok && T[i] <= 0;     // totals after the break are never looked at
            if (ok && T[i] < best) {  // first (highest address) of equal totals wins
                best = T[i];
                at = ra + 7 - i;
            }
        }
    }
    __syncwarp();
    const int best0 = __shfl_sync(FULL, best, (threadIdx.x & 31) & ~1), at0 = __shfl_sync(FULL, at, (threadIdx.x & 31) & ~1);
    const int at1 = __shfl_sync(FULL, at, (threadIdx.x & 31) | 1);
    uint32_t lowest_k = k;
    if (brk_a != NONE && best0 < lowest) lowest_k = (uint32_t)at0 - L3;  // a lower total just before the break
    else if (low_a != NONE) lowest_k = (uint32_t)at1 - L3;
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

// Pigeonhole barcode match on the compact tables (FastIdx): both half-key probes of a class are
// issued before either is consumed; a probe stops at the first slot whose tag matches (tags are
// unique per table, checked when the sheet is packed).  Same contract as hidx_match.
template <int NR>
__device__ __forceinline__ void fidx_match(const uint32_t (&raw)[NR], const HalfIdx &H, const FastIdx &F,
                                           const uint32_t *hcls, uint32_t S, uint32_t &lowest, uint32_t &best,
                                           uint32_t &last) {
    constexpr int NWMAX = NR - 1;
    const uint32_t nw = H.nw, nwp = H.nwp, tmask = H.tsize - 1u;
    lowest = 0xFFFFFFFFu;
    best = 0xFFFFFFFFu;
    last = 0;
    for (uint32_t c = 0; c < H.n_classes; c++) {
        const uint32_t *care = hcls + c * HIDX_CLS_ROWS * nwp;
        uint32_t tag[2] = {0u, 0u};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t *hm = care + (1 + 3 * h) * nwp;
#pragma unroll
            for (int w = 0; w < NWMAX; w++)
                if (w < (int)nw) tag[h] += (raw[w] & hm[w]) * hm[nwp + w];
        }
        const uint2 *tab0 = F.table + (size_t)(c * 2) * H.tsize, *tab1 = tab0 + H.tsize;
        uint32_t sl0 = (tag[0] ^ (tag[0] >> 15)) & tmask, sl1 = (tag[1] ^ (tag[1] >> 15)) & tmask;
        uint2 e0 = __ldg(&tab0[sl0]), e1 = __ldg(&tab1[sl1]);
        while (e0.y && e0.x != tag[0]) {
            sl0 = (sl0 + 1) & tmask;
            e0 = __ldg(&tab0[sl0]);
        }
        while (e1.y && e1.x != tag[1]) {
            sl1 = (sl1 + 1) & tmask;
            e1 = __ldg(&tab1[sl1]);
        }
        // candidate chains of the two halves (usually one sample each, usually the same one)
        uint32_t s = e0.y ? (e0.y & 0xFFFFu) - 1u : 0xFFFFFFFFu;
        bool more = (e0.y >> 16) != 0;
        uint32_t s_other = e1.y ? (e1.y & 0xFFFFu) - 1u : 0xFFFFFFFFu;
        bool more_other = (e1.y >> 16) != 0;
        if (s == s_other && !more && !more_other) s_other = 0xFFFFFFFFu;  // same single sample twice
        int h = 0;
        for (;;) {
            if (s == 0xFFFFFFFFu) {
                if (h) break;
                h = 1;
                s = s_other;
                more = more_other;
                if (s == 0xFFFFFFFFu) break;
            }
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
            uint32_t n = 0xFFFFu;
            if (more) n = __ldg(&F.next[(size_t)(c * 2 + h) * S + s]);
            s = n == 0xFFFFu ? 0xFFFFFFFFu : n;
        }
    }
    if (lowest > 1u) lowest = 0xFFFFFFFFu;  // farther samples were not enumerated completely
}

// dst[i] = ((u8)(qual[i] - 33) < minq) ? 'N' : seq[i]   (fasta_mask_by_quality.rs:40-43), thread-serial,
// four bytes per step once dst is word aligned.  7-bit input (see nl_flags7); minq <= 223 on the word
// path (bytes below '!' wrap to >= 223 and are then never masked).
__device__ __forceinline__ void mask_copy(uint8_t *dst, const uint8_t *seq, const uint8_t *qual, uint32_t len,
                                          uint32_t minq) {
#define SK_MASK_BYTE()                                                      \
    {                                                                       \
        const uint8_t q = (uint8_t)(*qual++ - 33u);                         \
        const uint8_t s = *seq++;                                           \
        *dst++ = q < minq ? (uint8_t)'N' : s;                               \
        len--;                                                              \
    }
    while (len && ((uint32_t)(uintptr_t)dst & 3u)) SK_MASK_BYTE()
    if (len >= 4 && minq <= 223u) {
        const uint32_t shs = ((uint32_t)(uintptr_t)seq & 3u) * 8u, shq = ((uint32_t)(uintptr_t)qual & 3u) * 8u;
        const uint32_t *sw = (const uint32_t *)((uintptr_t)seq & ~(uintptr_t)3);
        const uint32_t *qw = (const uint32_t *)((uintptr_t)qual & ~(uintptr_t)3);
        uint32_t *dw = (uint32_t *)dst;
        const uint32_t hi_c = 33u + minq;  // masked <=> 33 <= q < 33 + minq
        const uint32_t k_hi = hi_c >= 128u ? 0u : (128u - hi_c) * 0x01010101u;
        const bool no_hi = hi_c >= 128u;
        uint32_t slo = *sw++, qlo = *qw++;
        uint32_t nwords = len >> 2;
        for (uint32_t i = 0; i < nwords; i++) {
            const uint32_t shi = *sw++, qhi = *qw++;
            const uint32_t s = __funnelshift_r(slo, shi, shs), q = __funnelshift_r(qlo, qhi, shq);
            slo = shi;
            qlo = qhi;
            const uint32_t ge = (q + 0x5F5F5F5Fu);                       // bit 7 <=> q >= 33
            const uint32_t lt = no_hi ? 0xFFFFFFFFu : ~(q + k_hi);        // bit 7 <=> q < 33 + minq
            const uint32_t f = ge & lt & 0x80808080u;
            const uint32_t m = (f >> 7) * 0xFFu;
            *dw++ = (s & ~m) | (0x4E4E4E4Eu & m);
        }
        const uint32_t done = nwords * 4u;
        dst += done;
        seq += done;
        qual += done;
        len -= done;
    }
    while (len) SK_MASK_BYTE()
#undef SK_MASK_BYTE
}

template <class G, int OP, int NWMAX>
__global__ void __launch_bounds__(G::NT, G::MIN_CTAS) sk_fast_kernel(const __grid_constant__ KParams p) {
    constexpr int NT = G::NT, MAXREC = G::MAXREC, MAXLINES = G::MAXLINES;
    constexpr int TB = G::PPL * 16;              // bytes per thread in the scan
    constexpr int TCH = G::CHUNK / TB;           // threads whose bytes lie inside the chunk
    static_assert(G::CHUNK % TB == 0 && G::WIN_MAX == NT * TB && G::PPL == 5, "fast geometry");
    static_assert(MAXREC == NT && NT % 64 == 0 && NT <= 256, "one thread per record");
    constexpr bool IS_DEMUX = (OP == OP_DEMUX1 || OP == OP_DEMUX2);
    constexpr bool ORDERED = !IS_DEMUX;
    using FL = FLayout<G>;

    uint8_t *win = sk_smem + FL::win;
    uint8_t *stage = sk_smem + FL::stage;
    uint16_t *ls = (uint16_t *)(sk_smem + FL::ls);
    uint16_t *r_outoff = (uint16_t *)(sk_smem + FL::rec);
    uint16_t *r_outlen = r_outoff + MAXREC;
    uint16_t *r_alen = r_outlen + MAXREC;
    uint16_t *r_blen = r_alen + MAXREC;
    uint16_t *r_cut0 = r_blen + MAXREC;
    uint16_t *r_cut1 = r_cut0 + MAXREC;
    uint16_t *r_k = r_cut1 + MAXREC;
    uint16_t *r_body = r_k + MAXREC;
    int16_t *r_sample = (int16_t *)(r_body + MAXREC);
    uint8_t *r_taglen = (uint8_t *)(r_sample + MAXREC);
    uint8_t *r_mode = r_taglen + MAXREC;
    uint8_t *sh_lut = sk_smem + FL::lut;
    FMisc *M = (FMisc *)(sk_smem + FL::misc);
    uint32_t *hcls = (uint32_t *)(sk_smem + FL::dyn);
    const uint32_t S = IS_DEMUX ? p.sheet.S : 0u;
    const uint32_t hcls_words = (OP == OP_DEMUX1) ? p.sheet.hidx.n_classes * HIDX_CLS_ROWS * p.sheet.hidx.nwp : 0u;
    uint32_t *ccount = hcls + ((hcls_words + 3u) & ~3u);
    const bool cc_smem = (OP == OP_DEMUX1) && S <= (uint32_t)FAST_CCOUNT_MAX;
    uint8_t *sh_ulen = (uint8_t *)(ccount + (cc_smem ? ((S + 3u) & ~3u) : 0u));  // UMI length of every sample
    constexpr int NW = NT / 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevStats *st = p.stats;

    if (tid == 0) mbar_init(&M->mbar, 1);
    if (IS_DEMUX) {
        for (uint32_t i = tid; i < 256; i += NT) sh_lut[i] = p.sheet.lut[i];
        for (uint32_t i = tid; i < S; i += NT)
            sh_ulen[i] = (uint8_t)(p.sheet.wide ? __popcll(((const unsigned long long *)p.sheet.umask)[i]) : __popc(p.sheet.umask[i]));
        if (OP == OP_DEMUX1) {
            for (uint32_t i = tid; i < hcls_words; i += NT) hcls[i] = p.sheet.hidx.cls[i];
            if (cc_smem)
                for (uint32_t i = tid; i < S; i += NT) ccount[i] = 0;
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tid == 0) M->chunk = atomicAdd(&st->ticket, 1u);
    __syncthreads();

#ifdef SK_PHASE_TIMING
    unsigned long long ph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long ph_x = 0;  // tid 0: trim warp's plan; tid NT/2: header warp's plan; last warp lane 0: look-back
    long long t_prev = clock64();
#endif
    uint32_t parity = 0, flip = 0;
    bool store_pending = false;
    uint32_t my_total = 0, my_ident = 0;  // DEMUX1 counters of this thread's records (flushed at the end)
    const bool fused = IS_DEMUX && p.fused_trim >= 0;

    // Chunks are handed out by a ticket counter (a static round-robin deal was measured slower: one
    // lagging SM then holds up the look-back of every CTA behind it).
    for (;;) {
        const uint32_t c = M->chunk;
        if (c >= p.n_chunks) break;
        FK_T(0);
        const uint64_t c0 = (uint64_t)c * G::CHUNK;
        uint64_t wend = c0 + G::WIN_MAX;
        if (wend > p.n) wend = p.n;
        const uint32_t wlen = (uint32_t)(wend - c0);
        const bool at_end = (wend == p.n);

        // ---- P1 window load (TMA bulk copy; plain loads for the ragged tail of the stream)
        const uint32_t bulk = wlen & ~15u;
        if (tid == 0 && bulk) {
            fence_proxy_async();
            mbar_expect_tx(&M->mbar, bulk);
            bulk_g2s(win, p.in + c0, bulk, &M->mbar);
        }
        if (bulk != (uint32_t)G::WIN_MAX && tid < 16) {
            const uint32_t o = bulk + tid;
            if (o < (uint32_t)G::WIN_MAX) win[o] = (o < wlen) ? p.in[c0 + o] : (uint8_t)0;
        }
        if (bulk) {
            mbar_wait_parked(&M->mbar, parity);
            parity ^= 1;
        }
        if (bulk != (uint32_t)G::WIN_MAX) __syncthreads();
        FK_T(1);

        // ---- P2 newline scan.  A '\n' at window offset q starts a line at q+1; the chunk owns the
        // line starts of the newlines inside its CHUNK bytes (and the line at byte 0 of the stream).
        // A '\n' that is the last byte of the stream starts nothing.
        const uint32_t ls_hi = at_end ? (wlen ? wlen - 1 : 0) : wlen;
        const uint32_t o0 = (uint32_t)tid * TB;
        uint32_t m0 = 0, m1 = 0, m2 = 0, hib = 0;
        if (o0 + TB <= ls_hi) {
            const uint4 v0 = *(const uint4 *)(win + o0), v1 = *(const uint4 *)(win + o0 + 16);
            const uint4 v2 = *(const uint4 *)(win + o0 + 32), v3 = *(const uint4 *)(win + o0 + 48);
            const uint4 v4 = *(const uint4 *)(win + o0 + 64);
            hib = v0.x | v0.y | v0.z | v0.w | v1.x | v1.y | v1.z | v1.w | v2.x | v2.y | v2.z | v2.w | v3.x | v3.y | v3.z |
                  v3.w | v4.x | v4.y | v4.z | v4.w;
            m0 = nl_map_nat(v0) | (nl_map_nat(v1) << 16);
            m1 = nl_map_nat(v2) | (nl_map_nat(v3) << 16);
            m2 = nl_map_nat(v4);
        } else {
            uint32_t y[5];
#pragma unroll
            for (int q = 0; q < 5; q++) {
                const uint32_t o = o0 + q * 16;
                y[q] = 0;
                if (o < wlen) {
                    const uint4 v = *(const uint4 *)(win + o);  // bytes past wlen in the last piece are zero
                    hib |= v.x | v.y | v.z | v.w;
                    y[q] = nl_map_nat(v) & bits_below((int)ls_hi - (int)o);
                }
            }
            m0 = y[0] | (y[1] << 16);
            m1 = y[2] | (y[3] << 16);
            m2 = y[4];
        }
        if (__any_sync(0xffffffffu, (hib & 0x80808080u) != 0) && lane == 0) atomicOr(&st->flags, F_NON_ASCII);
        const uint32_t cnt_all = __popc(m0) + __popc(m1) + __popc(m2);
        const uint32_t cnt_chunk = tid < TCH ? cnt_all : 0u;
        uint32_t tot;
        const uint32_t pre = block_scan_fast<NT>((cnt_chunk << 16) | cnt_all, M->scratch, flip, tot);
        const uint32_t extra = (c == 0) ? 1u : 0u;
        const uint32_t nls = (tot & 0xFFFFu) + extra;
        const uint32_t nls_chunk = (tot >> 16) + extra;
        FK_T(2);

        // ---- P3 publish this chunk's line count for the look-back; everybody writes the line table
        if (tid == NT - 32) lookback_publish(p.tile_lines, c, nls_chunk);
        {
            uint32_t idx = (pre & 0xFFFFu) + extra;
            if (tid == 0 && extra) ls[0] = 0;
            uint32_t base = o0 + 1u;
#pragma unroll
            for (int wi = 0; wi < 3; wi++) {
                uint32_t m = wi == 0 ? m0 : (wi == 1 ? m1 : m2);
                while (m) {
                    const uint32_t i = (uint32_t)__ffs((int)m) - 1u;
                    m &= m - 1;
                    if (idx < (uint32_t)MAXLINES) ls[idx] = (uint16_t)(base + i);
                    idx++;
                }
                base += 32u;
            }
            if (tid >= NT - 8) {  // sentinels: lines past the last line start read as "end of window"
                const uint32_t k = nls + (uint32_t)(tid - (NT - 8));
                if (k < (uint32_t)MAXLINES + 8u) ls[k] = (uint16_t)wlen;
            }
        }
        __syncthreads();
        FK_T(3);
#define LB(x) ((uint32_t)ls[(x)])

        // ---- P4 framing.  Record i is lines 4i..4i+3 of the stream (common.rs:106-112), so the chunk needs
        // the global index g0 of its first line -- the look-back result, which every resident CTA would
        // otherwise wait for at the same point of every chunk.  The framing is therefore guessed from the
        // text (a line that starts with '@' whose second successor starts with '+'), the plan runs on the
        // guess while the last warp collects the real prefix, and the guess is checked against it
        // afterwards; a wrong guess (possible on malformed or adversarial text only) repeats the plan
        // with the true framing.  The plan has no side effects outside the per-record arrays.
        bool spec = false;
        uint32_t j0 = 0;
        if (c != 0 && p.rec_limit == ~0ull && nls >= 6u) {
            const uint4 l8 = *(const uint4 *)ls;  // the first eight line starts
            const uint32_t s0 = l8.x & 0xFFFFu, s1 = l8.x >> 16, s2 = l8.y & 0xFFFFu, s3 = l8.y >> 16;
            const uint32_t s4 = l8.z & 0xFFFFu, s5 = l8.z >> 16;
            const uint32_t b0 = win[s0], b1 = win[s1], b2 = win[s2], b3 = win[s3], b4 = win[s4], b5 = win[s5];
            const bool k0 = b0 == '@' && b2 == '+', k1 = b1 == '@' && b3 == '+';
            const bool k2 = b2 == '@' && b4 == '+', k3 = b3 == '@' && b5 == '+';
            j0 = k0 ? 0u : k1 ? 1u : k2 ? 2u : 3u;
            spec = (k0 || k1 || k2 || k3) && j0 < nls_chunk;
        }
        const bool speculated = spec;
#ifdef SK_PHASE_TIMING
        const long long t_plan0 = clock64();
#endif
        if (!spec) {
            if (warp == NW - 1) {
                const uint64_t excl = lookback_consume(p.tile_lines, c, nls_chunk, lane);
                if (lane == 0) M->g0 = excl;
            }
            __syncthreads();
            j0 = (4u - (uint32_t)(M->g0 & 3u)) & 3u;
        }
        const uint32_t tl = fused ? (uint32_t)tid % (NT / 2) : (uint32_t)tid;
        const uint32_t tstride = fused ? NT / 2 : NT;
        const bool do_trim = (fused && tid < NT / 2) || OP == OP_TRIM;
        const bool do_main = (!fused || tid >= NT / 2) && OP != OP_TRIM;
        const int trim_q = OP == OP_TRIM ? (int)p.min_baseq : p.fused_trim;
        uint32_t nrec = 0;
        uint64_t g0 = 0;
        bool bail = false;
        for (;;) {
            nrec = j0 < nls_chunk ? (nls_chunk - 1 - j0) / 4u + 1u : 0u;
            if (!spec) {
                g0 = M->g0;
                const uint64_t first = (g0 + j0) >> 2;
                if (first >= p.rec_limit) nrec = 0;
                else if ((uint64_t)nrec > p.rec_limit - first) nrec = (uint32_t)(p.rec_limit - first);
            }
            bail = false;
            if (nrec) {
                uint32_t jend = j0 + nrec * 4u;
                const bool eof_ok = at_end && p.final_batch;
                if (jend >= nls && !eof_ok) {
                    if (at_end) {  // non-final batch: the trailing incomplete record(s) stay for the next batch
                        nrec = nls > j0 + 4u ? (nls - j0 - 5u) / 4u + 1u : 0u;
                        jend = j0 + nrec * 4u;
                    } else {
                        bail = true;  // a record runs past the overhang
                    }
                }
                if (!bail && nrec) {
                    const uint32_t need = jend < nls ? jend : nls - 1;
                    if (need >= (uint32_t)MAXLINES || nrec > (uint32_t)MAXREC) bail = true;
                }
                if (bail) nrec = 0;
            }

            // The last warp collects the real prefix while the others plan (it carries records of its own
            // only in very dense chunks), so the inclusive prefix is also published as early as possible.
            const bool lb_early = (fused ? (uint32_t)(NT / 2 - 32) : (uint32_t)(NT - 32)) >= nrec;
            if (spec && warp == NW - 1 && lb_early) {
                const uint64_t excl = lookback_consume(p.tile_lines, c, nls_chunk, lane);
                if (lane == 0) M->g0 = excl;
#ifdef SK_PHASE_TIMING
                if (lane == 0) ph_x += (unsigned long long)(clock64() - t_plan0);  // look-back done
#endif
            }

            // ---- P5 plan, one lane per record, whole warps per task and every step warp-synchronous
            // (lanes that took different branches meet again at the __syncwarp that closes the step).
            // Fused trim+demultiplex: the quality trim runs on the lower half of the CTA, the header work
            // on the upper half.  Failures are left in the record arrays and reported in P6.
            if (do_trim) {
                for (uint32_t r0 = 0; r0 < nrec; r0 += tstride / 2) {
                    const uint32_t r = r0 + (tl >> 1), sub = tl & 1u;  // two adjacent lanes per record
                    const bool has = r < nrec;
                    const uint32_t j = j0 + (has ? r : 0u) * 4u;
                    const uint32_t L0 = LB(j), L1 = LB(j + 1), L2 = LB(j + 2), L3 = LB(j + 3), L4 = LB(j + 4);
                    uint8_t mode = OP == OP_TRIM ? B_NONE : B_FAIL;
                    uint32_t kk = 0, body = 0, errk = 0;
                    bool ok = has;
                    if (OP == OP_TRIM) {
                        if (has && win[L0] != '@') {  // fasta_trim_by_quality.rs:20-22
                            errk = K_BAD_HEADER;
                            ok = false;
                        }
                    } else {
                        ok = has && L1 > L0 && win[L1 - 1] == '\n';
                    }
                    bool fine;
                    if (trim_q <= 222) {
                        fine = plan_trim_pair(win, ok, sub, L1, L2, L3, L4, trim_q, mode, kk, body);
                    } else {
                        fine = ok ? plan_trim_body(win, L1, L2, L3, L4, trim_q, mode, kk, body) : true;
                        __syncwarp();
                    }
                    if (has && sub == 0) {
                        if (OP == OP_TRIM) {
                            uint32_t outlen = 0;
                            if (!ok) {
                                mode = B_NONE;
                            } else if (!fine) {
                                errk = K_SEQ_SHORT;
                                mode = B_NONE;
                            } else {
                                outlen = (L1 - L0) + body;  // header verbatim (:23) + body
                            }
                            if (mode == B_NONE) kk = errk;  // a failed record keeps its failure kind here
                            r_outlen[r] = (uint16_t)(outlen > 0x3FFFu ? 0x3FFFu : outlen);
                        } else if (!ok || !fine) {
                            mode = B_FAIL;
                        }
                        r_k[r] = (uint16_t)kk;
                        r_body[r] = (uint16_t)(body > 0xFFFFu ? 0xFFFFu : body);
                        r_mode[r] = mode;
                    }
                }
            }
            if (do_main) {
                for (uint32_t r0 = 0; r0 < nrec; r0 += tstride) {
                    const uint32_t r = r0 + tl;
                    const bool has = r < nrec;
                    const uint32_t j = j0 + (has ? r : 0u) * 4u;
                    const uint32_t L0 = LB(j), L1 = LB(j + 1);
                    if (OP == OP_MASK) {
                        if (!has) continue;
                        uint8_t mode = B_NONE;
                        uint32_t kk = 0, outlen = 0;
                        if (win[L0] != '@') {  // fasta_mask_by_quality.rs:21-23
                            kk = K_BAD_HEADER;
                        } else {
                            const uint32_t L2 = LB(j + 2), L3 = LB(j + 3), L4 = LB(j + 4);
                            uint32_t sl = L2 - L1, ql = L4 - L3;
                            if (sl && win[L2 - 1] == '\n') sl--;  // :32
                            if (ql && win[L4 - 1] == '\n') ql--;  // :33
                            if (sl != ql) {                       // :35-37
                                kk = K_LEN_MISMATCH;
                            } else {
                                mode = B_MASK;
                                kk = sl;
                                outlen = (L1 - L0) + 2 * sl + 4;  // header, masked, "\n+\n", qual, "\n"  (:26,:44)
                            }
                        }
                        r_k[r] = (uint16_t)kk;  // a failed record (B_NONE) keeps its failure kind here
                        r_mode[r] = mode;
                        r_outlen[r] = (uint16_t)(outlen > 0x3FFFu ? 0x3FFFu : outlen);
                    } else if (OP == OP_DEMUX1) {
                        // fasta_demultiplex.rs:117-194: validate, locate the barcode, match, decide.
                        // Outcome in the record arrays: sample >= 0 assigned; -2 ambiguous (best, last,
                        // mismatches in alen, blen, taglen); -1 unassigned (taglen = failure kind, or 0xFF
                        // for a record that never reached the match).
                        int sample = -1;
                        uint32_t alen = 0, blen = 0, cut0 = 0, cut1 = 0, taglen = 0xFFu;
                        // step 1: header checks and the leftmost " BC:x" (:118-120, :138-141)
                        bool live = has;
                        uint32_t stp = 0;
                        if (live) {
                            if (win[L0] != '@') {
                                taglen = K_BAD_HEADER;
                                live = false;
                            } else if (fused && !(L1 > L0 && win[L1 - 1] == '\n')) {
                                taglen = K_TRUNC_FUSED;
                                live = false;
                            } else if (!bc_find16(win, sh_lut, L0, L1, stp)) {
                                taglen = K_NO_BC;
                                live = false;
                            }
                        }
                        __syncwarp();
                        // step 2: the greedy class run must be exactly L long (:38, :148-150)
                        const uint32_t Lb = p.sheet.L;
                        uint32_t raw[NWMAX + 1];
                        const uint32_t bs = stp + 4;
                        if (live) {
                            cut0 = stp - L0;
                            cut1 = cut0 + 4 + Lb;
                            load_raw<NWMAX + 1>(win, bs, (Lb + 4u) >> 2, raw);
                            if (!class_run_is<NWMAX + 1>(raw, sh_lut, Lb, L1 - bs)) {
                                taglen = K_BC_LEN;
                                live = false;
                            }
                        }
                        __syncwarp();
                        // step 3: match against the sheet and decide (:154-194)
                        if (live) {
                            uint32_t lowest, best, last;
                            fidx_match<NWMAX + 1>(raw, p.sheet.hidx, p.sheet.fidx, hcls, S, lowest, best, last);
                            taglen = 0;
                            if (lowest <= 1u) {        // :172
                                if (best == last) {    // :173-178
                                    sample = (int)best;
                                } else {  // :184-188
                                    sample = -2;
                                    alen = best;
                                    blen = last;
                                    taglen = lowest;
                                }
                            }
                        }
                        __syncwarp();
                        if (sample >= 0) {
                            header_pieces(win, L0, L1, L0 + cut0, L0 + cut1, alen, blen);  // drain (:145) + trim_end (:206)
                            const uint32_t ul = sh_ulen[sample];
                            taglen = ul ? 5 + ul : 0;  // " UMI:" + umi (:207)
                        }
                        __syncwarp();
                        if (has) {
                            r_sample[r] = (int16_t)sample;
                            r_alen[r] = (uint16_t)alen;
                            r_blen[r] = (uint16_t)blen;
                            r_cut0[r] = (uint16_t)cut0;
                            r_cut1[r] = (uint16_t)cut1;
                            r_taglen[r] = (uint8_t)taglen;
                        }
                    } else if (OP == OP_DEMUX2) {
                        // fasta_demultiplex.rs:215-229: the header of mate 2 without its " BC:" field; whether
                        // the pair was assigned is looked up in P6, once the record index is known
                        uint32_t alen = 0, blen = 0;
                        uint32_t c0h = L1, c1h = L1, fa = 0;
                        const bool found = has && p.out && bc_find16(win, sh_lut, L0, L1, fa);  // :219-227
                        __syncwarp();
                        if (found) {
                            c0h = fa;
                            c1h = class_run_end(win, sh_lut, fa + 4, L1);
                        }
                        __syncwarp();
                        if (has && p.out) header_pieces(win, L0, L1, c0h, c1h, alen, blen);  // :229
                        __syncwarp();
                        if (has) {
                            r_alen[r] = (uint16_t)alen;
                            r_blen[r] = (uint16_t)blen;
                            r_cut1[r] = (uint16_t)(c1h - L0);
                        }
                    }
                }
            }
#ifdef SK_PHASE_TIMING
            if (tid == 0 || tid == NT / 2) ph_x += (unsigned long long)(clock64() - t_plan0);  // this warp's plan done
#endif
            if (spec && warp == NW - 1 && !lb_early) {  // ... or after its own part of the plan
                const uint64_t excl = lookback_consume(p.tile_lines, c, nls_chunk, lane);
                if (lane == 0) M->g0 = excl;
            }
            __syncthreads();  // the parts of a record's plan meet; g0 is known
            if (!spec) break;
            spec = false;
            g0 = M->g0;
            const uint32_t jt = (4u - (uint32_t)(g0 & 3u)) & 3u;
            if (jt == j0) break;
            j0 = jt;  // wrong guess: plan again on the true framing
            __syncthreads();
        }
        const uint64_t rec0 = (g0 + j0) >> 2;
        if (tid == 0) {
            if (c == p.n_chunks - 1) st->n_lines = g0 + nls_chunk;
            if (nrec) {
                atomicAdd(&st->n_records, (unsigned long long)nrec);
                atomicMax(&st->consumed, (unsigned long long)(c0 + LB(j0 + nrec * 4u)));
            }
        }
        FK_T(4);

        // ---- P6 outcome of every record (counters, deferred failures and ambiguity events), its output
        // length and its place in the chunk's output (input order)
        uint32_t outlen = 0;
        {
            const uint32_t r = (uint32_t)tid;
            if (r < nrec) {
                if (IS_DEMUX) {
                    int sample;
                    if (OP == OP_DEMUX1) {
                        sample = r_sample[r];
                        const uint32_t tg = r_taglen[r];
                        if (sample == -1 && tg != 0xFFu && tg != 0u) report_err(st, rec0 + r, tg);
                        if (sample != -1 || tg == 0u) my_total++;  // :169
                        if (sample >= 0) {                          // :177-178
                            my_ident++;
                            if (cc_smem) atomicAdd(&ccount[sample], 1u);
                            else atomicAdd(&p.counts[sample], 1ull);
                        } else if (sample == -2) {  // :184-188
                            const uint32_t ei = atomicAdd(&st->n_events, 1u);
                            if (ei < p.events_cap) {
                                Event ev;
                                ev.record = (uint32_t)(rec0 + r);
                                ev.bc_off = (uint32_t)(c0 + LB(j0 + r * 4u) + r_cut0[r] + 4u);
                                ev.bc_off2 = 0xFFFFFFFFu;
                                ev.best = (int16_t)r_alen[r];
                                ev.last = (int16_t)r_blen[r];
                                ev.mismatches = tg;
                                p.events[ei] = ev;
                            } else {
                                atomicOr(&st->flags, F_EVENTS_OVERFLOW);
                            }
                        }
                    } else {
                        const uint64_t rec = rec0 + r;
                        sample = (p.out && rec < p.r1_stats->n_records) ? (int)p.assign[rec] : -1;
                        if (sample >= 0) {
                            const uint32_t ul = sh_ulen[sample];
                            r_taglen[r] = (uint8_t)(ul ? 5 + ul : 0);
                            const uint32_t j = j0 + r * 4u;
                            const uint32_t L0 = LB(j), L1 = LB(j + 1);
                            if (fused && !(L1 > L0 && win[L1 - 1] == '\n')) {
                                report_err(st, rec, K_TRUNC_FUSED);
                                sample = -1;
                            }
                        } else {
                            sample = -1;
                        }
                        r_sample[r] = (int16_t)sample;
                    }
                    if (sample >= 0) {
                        const uint32_t j = j0 + r * 4u;
                        uint32_t body = LB(j + 4) - LB(j + 1);  // three lines verbatim (:209-212)
                        uint8_t mode = B_VERBATIM;
                        bool fine = true;
                        if (fused) {
                            mode = r_mode[r];
                            body = r_body[r];
                            fine = mode != B_FAIL;
                        } else {
                            r_k[r] = 0;
                        }
                        if (!fine) {
                            report_err(st, rec0 + r, K_SEQ_SHORT);
                            if (OP == OP_DEMUX1) r_sample[r] = -1;
                            mode = B_NONE;
                        } else if (!p.out) {
                            mode = B_NONE;  // dry run: count only (:77-78,:179)
                        } else {
                            outlen = (uint32_t)r_alen[r] + r_blen[r] + r_taglen[r] + 1u + body;
                            if (outlen > 0x3FFFu) outlen = 0x3FFFu;
                        }
                        r_mode[r] = mode;
                    }
                } else {
                    outlen = r_outlen[r];
                    if (r_mode[r] == B_NONE && r_k[r]) report_err(st, rec0 + r, r_k[r]);
                }
            }
        }
        uint32_t t2;
        const uint32_t sc = block_scan_fast<NT>(outlen | (outlen ? 1u << 22 : 0u), M->scratch, flip, t2);
        const uint32_t chunk_out = t2 & 0x3FFFFFu, n_emit = t2 >> 22;
        const uint32_t my_off = sc & 0x3FFFFFu, my_rank = sc >> 22;
        FK_T(5);

        // ---- P7 reserve output space.  In-order operators chain a second look-back on output bytes (the
        // staging image must be aligned like its destination).  Demultiplex chunks take 16-byte aligned
        // space from a bump allocator: thread 0 issues the atomic here and needs its result only when it
        // stores the image, so the round trip overlaps the assembly.
        unsigned long long demux_base = 0;
        if (ORDERED) {
            if (warp == 0) {
                const uint64_t excl = lookback_wide(p.tile_out, c, chunk_out, lane);
                if (lane == 0) {
                    M->out_base = excl;
                    if (c == p.n_chunks - 1) {
                        st->out_bytes = excl + chunk_out;
                        st->out_extent = excl + chunk_out;
                    }
                }
            }
        } else if (tid == 0 && p.out) {
            demux_base = atomicAdd(&st->out_cursor, (unsigned long long)((chunk_out + 15u) & ~15u));
            if (chunk_out) atomicAdd(&st->out_bytes, (unsigned long long)chunk_out);
        }
        if (tid == 0 && store_pending) {  // the previous chunk's TMA store must have read the staging image
            bulk_wait_read0();
            store_pending = false;
        }
        if (tid < (int)nrec) {
            r_outoff[tid] = (uint16_t)my_off;
            r_outlen[tid] = (uint16_t)outlen;
        }
        __syncthreads();
        FK_T(6);
        const uint64_t out_base = ORDERED ? M->out_base : 0ull;
        const uint32_t shift = (uint32_t)(out_base & 15u);
        bool writable = p.out != nullptr && chunk_out > 0;
        if (writable && shift + chunk_out > (uint32_t)G::STAGE) {
            writable = false;
            bail = true;  // output does not fit the staging image
        }
        if (ORDERED && writable && out_base + ((chunk_out + 15u) & ~15u) > p.out_cap) {
            if (tid == 0) report_err(st, rec0, K_OUT_OVERFLOW);
            writable = false;
        }
        if (bail && tid == 0) atomicOr(&st->flags, F_NEED_GENERAL);

        // ---- P8 assemble the chunk's output image in shared memory, aligned like its destination.
        // Five jobs per record (header, tag and literals | two halves of each of the two body pieces),
        // numbered job-type major and dealt to all threads: a warp's lanes mostly share a job type,
        // every job is a short word-wise copy, and the halves meet on a 16-byte boundary of the image.
        // (Five, not six: 5 x 22 and 5 x 44 records fit one round of 128 / 256 threads.)
        if (writable) {
            uint8_t *sb = stage + shift;
            const uint32_t njobs = 5u * nrec;
            for (uint32_t jb = (uint32_t)tid; jb < njobs; jb += NT) {
                const uint32_t type = (jb >= nrec) + (jb >= 2u * nrec) + (jb >= 3u * nrec) + (jb >= 4u * nrec);
                const uint32_t r = jb - type * nrec;
                const uint32_t ol = r_outlen[r];
                uint8_t *jd = nullptr;
                const uint8_t *js = nullptr, *jq = nullptr;  // jq: qualities of a masked copy
                uint32_t jl = 0;
                if (ol) {
                    const uint32_t j = j0 + r * 4u;
                    const uint32_t L0 = LB(j), L1 = LB(j + 1);
                    uint8_t *d0 = sb + r_outoff[r];
                    const uint32_t kk = r_k[r];
                    const uint8_t mode = r_mode[r];
                    uint32_t hlen, alen = 0, blen = 0, taglen = 0;
                    if (ORDERED) {
                        hlen = L1 - L0;
                    } else {
                        alen = r_alen[r];
                        blen = r_blen[r];
                        taglen = r_taglen[r];
                        hlen = alen + blen + taglen + 1;
                    }
                    uint8_t *db = d0 + hlen;
                    if (type == 0) {
                        jd = d0;
                        js = win + L0;
                        jl = ORDERED ? hlen : alen;
                    }
                    if (type != 0) {
                        // body piece 0 / 1: verbatim -> the two halves of the three lines; trim and mask ->
                        // sequence and qualities
                        const uint32_t piece = (type - 1u) >> 1, hi = (type - 1u) & 1u;
                        uint8_t *pd = nullptr;
                        uint32_t ps = 0, pn = 0;
                        if (mode == B_VERBATIM) {
                            const uint32_t n = LB(j + 4) - L1;
                            uint32_t m = (uint32_t)((((uintptr_t)db + n / 2 + 15u) & ~(uintptr_t)15) - (uintptr_t)db);
                            if (m > n) m = n;
                            pd = piece ? db + m : db;
                            ps = piece ? L1 + m : L1;
                            pn = piece ? n - m : m;
                        } else if (mode == B_TRIM || mode == B_MASK) {
                            pd = piece ? db + kk + 3 : db;
                            ps = piece ? LB(j + 3) : L1;
                            pn = kk;
                            if (mode == B_MASK && !piece) jq = win + LB(j + 3);
                        }
                        uint32_t h = (uint32_t)((((uintptr_t)pd + pn / 2 + 15u) & ~(uintptr_t)15) - (uintptr_t)pd);
                        if (h > pn) h = pn;
                        jd = hi ? pd + h : pd;
                        js = win + (hi ? ps + h : ps);
                        jl = hi ? pn - h : h;
                        if (jq) jq += hi ? h : 0u;
                    } else {  // type 0 also writes the record's literals
                        if (!ORDERED) {  // the (usually empty) piece after the cut, the tag and the newline
                            uint8_t *d = d0 + alen;
                            const uint8_t *sB = win + L0 + r_cut1[r];
                            for (uint32_t i = 0; i < blen; i++) d[i] = sB[i];
                            d += blen;
                            if (taglen) {
                                d[0] = ' '; d[1] = 'U'; d[2] = 'M'; d[3] = 'I'; d[4] = ':';
                                uint8_t *gu = p.umi + (rec0 + r) * p.sheet.Umax;
                                const uint32_t ul = taglen - 5;
                                if (OP == OP_DEMUX1) {
                                    // UMI = observed chars where the sheet barcode has 'U' (:200-203); also
                                    // parked in the side table for mate 2
                                    const uint8_t *ob = win + L0 + r_cut0[r] + 4;
                                    const int sm = r_sample[r];
                                    unsigned long long m = p.sheet.wide ? ((const unsigned long long *)p.sheet.umask)[sm]
                                                                        : (unsigned long long)p.sheet.umask[sm];
                                    const uint32_t u0 = (uint32_t)__ffsll((long long)m) - 1u;
                                    if ((m >> u0) == ((1ull << ul) - 1ull)) {  // the U positions are one run (the usual sheet)
                                        for (uint32_t t = 0; t < ul; t++) {
                                            const uint8_t ch = ob[u0 + t];
                                            d[5 + t] = ch;
                                            gu[t] = ch;
                                        }
                                    } else {
                                        uint32_t t = 0;
                                        while (m) {
                                            const uint32_t q = (uint32_t)__ffsll((long long)m) - 1u;
                                            m &= m - 1;
                                            const uint8_t ch = ob[q];
                                            d[5 + t] = ch;
                                            gu[t] = ch;
                                            t++;
                                        }
                                    }
                                } else {
                                    for (uint32_t i = 0; i < ul; i += 8) {
                                        uint8_t tmp[8];
#pragma unroll
                                        for (int k = 0; k < 8; k++) tmp[k] = (i + k < ul) ? gu[i + k] : (uint8_t)0;
#pragma unroll
                                        for (int k = 0; k < 8; k++)
                                            if (i + k < ul) d[5 + i + k] = tmp[k];
                                    }
                                }
                                d += taglen;
                            }
                            d[0] = '\n';
                        }
                        if (mode == B_TRIM || mode == B_MASK) {
                            uint8_t *d = db + kk;
                            d[0] = '\n'; d[1] = '+'; d[2] = '\n';
                            d[3 + kk] = '\n';
                        } else if (mode == B_GARBAGE) {
                            db[0] = 'N'; db[1] = '\n'; db[2] = '+'; db[3] = '\n'; db[4] = '!'; db[5] = '\n';
                        }
                    }
                }
                if (jq) mask_copy(jd, js, jq, jl, p.min_baseq);
                else tcopy(jd, js, jl);
            }
            fence_proxy_async();  // staging writes -> visible to the TMA store issued after the barrier
        } else if (OP == OP_DEMUX1 && p.sheet.Umax) {
            // no output assembled here (dry run / overflow): mate 2 still needs the UMI side table
            for (uint32_t r = tid; r < nrec; r += NT) {
                const int sm = r_sample[r];
                if (sm < 0) continue;
                const uint8_t *ob = win + LB(j0 + r * 4u) + r_cut0[r] + 4;
                uint8_t *gu = p.umi + (rec0 + r) * p.sheet.Umax;
                unsigned long long m = p.sheet.wide ? ((const unsigned long long *)p.sheet.umask)[sm]
                                                    : (unsigned long long)p.sheet.umask[sm];
                uint32_t t = 0;
                while (m) {
                    const uint32_t q = (uint32_t)__ffsll((long long)m) - 1u;
                    m &= m - 1;
                    gu[t++] = ob[q];
                }
            }
        }
        FK_T(7);

        // ---- P9 side tables, next ticket, store
        if (tid < (int)nrec) {
            if (OP == OP_DEMUX1) p.assign[rec0 + tid] = r_sample[tid];
            if (IS_DEMUX && outlen && p.out) {
                Group g;
                g.sample = (uint16_t)r_sample[tid];
                g.len = (uint16_t)outlen;
                p.groups[rec0 + my_rank] = g;
            }
        }
        if (tid == 0) M->chunk = atomicAdd(&st->ticket, 1u);
        __syncthreads();  // window, staging image and record arrays are reused by the next chunk
        if (ORDERED) {
            if (writable) {
                const uint32_t span = shift + chunk_out;
                uint8_t *g16 = p.out + (out_base - shift);
                const uint32_t a = shift ? 16u : 0u;  // first whole 16-byte unit
                const uint32_t b2 = span & ~15u;      // end of the last whole unit
                if (b2 > a) {
                    if (tid == 0) {
                        bulk_s2g(g16 + a, stage + a, b2 - a);
                        bulk_commit();
                        store_pending = true;
                    }
                    if (tid < 32) {  // ragged first/last bytes live in units shared with the neighbours
                        const uint32_t t = (uint32_t)tid;
                        if (t < 16u) {
                            if (t >= shift && t < a && t < span) g16[t] = stage[t];
                        } else {
                            const uint32_t o = b2 + (t - 16u);
                            if (o < span) g16[o] = stage[o];
                        }
                    }
                } else {
                    for (uint32_t o = shift + tid; o < span; o += NT) g16[o] = stage[o];
                }
            }
        } else if (tid == 0 && p.out) {
            ChunkRow row;
            row.base = demux_base;
            row.first_group = (uint32_t)rec0;
            row.n_groups = writable ? n_emit : 0u;
            if (writable && demux_base + ((chunk_out + 15u) & ~15u) > p.out_cap) {
                report_err(st, rec0, K_OUT_OVERFLOW);
                row.n_groups = 0;
            } else if (writable) {
                bulk_s2g(p.out + demux_base, stage, (chunk_out + 15u) & ~15u);  // demux chunks own whole 16-byte units
                bulk_commit();
                store_pending = true;
            }
            p.rows[c] = row;
        }
        FK_T(8);
    }
    if (tid == 0 && store_pending) bulk_wait_read0();  // shared memory must outlive the last TMA store
#ifdef SK_PHASE_TIMING
    if (tid == 0)
        for (int i = 0; i < 16; i++)
            if (ph[i]) atomicAdd(&st->phase_cycles[i], ph[i]);
    if (tid == 0 && ph_x) atomicAdd(&st->phase_cycles[12], ph_x);
    if (tid == NT / 2 && ph_x) atomicAdd(&st->phase_cycles[13], ph_x);
    if (tid == NT - 32 && ph_x) atomicAdd(&st->phase_cycles[14], ph_x);
#endif
    if (OP == OP_DEMUX1) {  // fasta_demultiplex.rs:108-109,169,177-178
        const uint32_t wt = __reduce_add_sync(0xffffffffu, my_total), wi = __reduce_add_sync(0xffffffffu, my_ident);
        if (lane == 0 && wt) atomicAdd(&p.counts[S], (unsigned long long)wt);
        if (lane == 0 && wi) atomicAdd(&p.counts[S + 1], (unsigned long long)wi);
        if (cc_smem) {
            __syncthreads();
            for (uint32_t s = tid; s < S; s += NT)
                if (ccount[s]) atomicAdd(&p.counts[s], (unsigned long long)ccount[s]);
        }
    }
#undef LB
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
template <class G>
static uint32_t fast_smem(uint32_t S, uint32_t n_classes, uint32_t nwp, bool d1) {
    uint32_t o = FLayout<G>::dyn;
    if (d1) {
        o += ((n_classes * HIDX_CLS_ROWS * nwp + 3u) & ~3u) * 4u;
        if (S <= (uint32_t)FAST_CCOUNT_MAX) o += ((S + 3u) & ~3u) * 4u;
    }
    o += (S + 15u) & ~15u;  // per-sample UMI lengths
    return o;
}

template <class G, int OP, int NWMAX>
static int launch_fast_one(const KParams &p, int sm_count, cudaStream_t stream, const char **err) {
    auto kfn = sk_fast_kernel<G, OP, NWMAX>;
    const int smem = (int)fast_smem<G>(p.sheet.S, p.sheet.hidx.n_classes, p.sheet.hidx.nwp, OP == OP_DEMUX1);
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, G::NT, smem);
    if (e != cudaSuccess || per_sm < 1) {
        *err = e != cudaSuccess ? cudaGetErrorString(e) : "kernel does not fit on an SM";
        return -1;
    }
    long long grid = (long long)sm_count * per_sm;
    if (grid > (long long)p.n_chunks) grid = p.n_chunks;
    if (grid < 1) return 0;
    kfn<<<(unsigned)grid, G::NT, smem, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    return 1;
}

int fast_chunk_bytes(int geo) { return geo == GeoM::ID ? GeoM::CHUNK : GeoS::CHUNK; }

bool fast_supported(int geo, int op, const KParams &p) {
    if (op != OP_TRIM && op != OP_MASK && op != OP_DEMUX1 && op != OP_DEMUX2) return false;
    if (p.lpr != 4) return false;
    if (op == OP_DEMUX1 || op == OP_DEMUX2) {
        if (p.n_index || !p.sheet.hidx.n_classes || !p.sheet.fidx.table) return false;
        const uint32_t need = geo == GeoM::ID ? fast_smem<GeoM>(p.sheet.S, p.sheet.hidx.n_classes, p.sheet.hidx.nwp, true)
                                              : fast_smem<GeoS>(p.sheet.S, p.sheet.hidx.n_classes, p.sheet.hidx.nwp, true);
        if (need > 100u * 1024u) return false;
    }
    return true;
}

template <class G>
static int launch_fast_geo(int op, const KParams &p, int sm_count, cudaStream_t stream, const char **err) {
    const bool wide = p.sheet.wide != 0;
    switch (op) {
        case OP_TRIM: return launch_fast_one<G, OP_TRIM, 8>(p, sm_count, stream, err);
        case OP_MASK: return launch_fast_one<G, OP_MASK, 8>(p, sm_count, stream, err);
        case OP_DEMUX1:
            return wide ? launch_fast_one<G, OP_DEMUX1, 16>(p, sm_count, stream, err)
                        : launch_fast_one<G, OP_DEMUX1, 8>(p, sm_count, stream, err);
        case OP_DEMUX2: return launch_fast_one<G, OP_DEMUX2, 8>(p, sm_count, stream, err);
    }
    *err = "operator not handled by the lean engine";
    return -1;
}

int launch_fast_kernel(int geo, int op, const KParams &p, int sm_count, void *stream_, const char **err) {
    cudaStream_t stream = (cudaStream_t)stream_;
    return geo == GeoM::ID ? launch_fast_geo<GeoM>(op, p, sm_count, stream, err)
                           : launch_fast_geo<GeoS>(op, p, sm_count, stream, err);
}

}  // namespace sk
