// sk_record.cuh -- per-record device helpers of the warp engine (sk_warp.cu): newline maps, " BC:" search,
// class-run checks, block-wise quality trim, pigeonhole barcode match, masked copy.  Include after
// sk_device.cuh.  (Round 1's lean engine, sk_fast.cu, shared them; it was retired in round 2.)
#pragma once
#include "sk_device.cuh"

namespace sk {

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

// Running totals of fasta_trim_by_quality.rs:33-36 over the aligned 16-byte block at window offset a, in
// the order the reference examines the bytes, as keys K[i] = 16 * T[i] + i, T[i] = the total after byte
// a+15-i starting from `total`: the minimum key names the lowest total and, among equal totals, the byte
// examined first (:38); a key above 15 is a total above 0 (:37).  Bytes outside the quality string
// [L3,E) contribute nothing (they are replaced by the byte whose contribution is zero, sub = 33 +
// min_baseq <= 255).  A byte below '!' takes the wrapping u8 subtraction (:35): its contribution is 256
// more than q - sub, added through a second set of dot products over the flags of those bytes.
__device__ __forceinline__ void blk16_keys(const uint8_t *b, uint32_t a, uint32_t L3, uint32_t E, int sub, int total,
                                           int (&K)[16]) {
    uint4 v = *(const uint4 *)(b + a);
    if (a < L3 || a + 16u > E) {
        const uint32_t lo = a < L3 ? L3 - a : 0u;                      // first kept byte of the block
        const uint32_t hi = a + 16u > E ? (E > a ? E - a : 0u) : 16u;  // one past the last kept byte
        const uint32_t sub4 = (uint32_t)sub * 0x01010101u;
        auto keep = [&](uint32_t w, int q) {  // bytes [lo, hi) of the block keep their value, the others become `sub`
            const int l = (int)lo - 4 * q, h = (int)hi - 4 * q;
            const uint32_t ml = l <= 0 ? 0xFFFFFFFFu : (l >= 4 ? 0u : 0xFFFFFFFFu << (8 * l));
            const uint32_t mh = h >= 4 ? 0xFFFFFFFFu : (h <= 0 ? 0u : 0xFFFFFFFFu >> (8 * (4 - h)));
            const uint32_t m = ml & mh;
            return (w & m) | (sub4 & ~m);
        };
        v.x = keep(v.x, 0);
        v.y = keep(v.y, 1);
        v.z = keep(v.z, 2);
        v.w = keep(v.w, 3);
    }
    const uint32_t H = 0x80808080u, C = 0x21212121u;
    const uint32_t gx = ((v.x | H) - C) | v.x, gy = ((v.y | H) - C) | v.y;  // bit 7 clear <=> byte below '!'
    const uint32_t gz = ((v.z | H) - C) | v.z, gw = ((v.w | H) - C) | v.w;  // (a filler byte may be >= 0x80)
    const int s16 = 16 * sub;
    int base = 16 * total;
#define SK_KEY4(w, q)                                                                       \
    K[4 * q + 0] = (int)__dp4a(w, 0x10000000u, (uint32_t)(base - s16 + 4 * q));             \
    K[4 * q + 1] = (int)__dp4a(w, 0x10100000u, (uint32_t)(base - 2 * s16 + 4 * q + 1));     \
    K[4 * q + 2] = (int)__dp4a(w, 0x10101000u, (uint32_t)(base - 3 * s16 + 4 * q + 2));     \
    K[4 * q + 3] = (int)__dp4a(w, 0x10101010u, (uint32_t)(base - 4 * s16 + 4 * q + 3));     \
    base = K[4 * q + 3] - (4 * q + 3);
    SK_KEY4(v.w, 0)
    SK_KEY4(v.z, 1)
    SK_KEY4(v.y, 2)
    SK_KEY4(v.x, 3)
#undef SK_KEY4
    if ((~(gx & gy & gz & gw)) & H) {  // rare: add 256 per byte below '!' examined so far
        const uint32_t fx = (~gx & H) >> 7, fy = (~gy & H) >> 7, fz = (~gz & H) >> 7, fw = (~gw & H) >> 7;
        uint32_t n = 0;
#define SK_BAD4(f, q)                                               \
    K[4 * q + 0] += 4096 * (int)__dp4a(f, 0x01000000u, n);          \
    K[4 * q + 1] += 4096 * (int)__dp4a(f, 0x01010000u, n);          \
    K[4 * q + 2] += 4096 * (int)__dp4a(f, 0x01010100u, n);          \
    n = __dp4a(f, 0x01010101u, n);                                  \
    K[4 * q + 3] += 4096 * (int)n;
        SK_BAD4(fw, 0)
        SK_BAD4(fz, 1)
        SK_BAD4(fy, 2)
        SK_BAD4(fx, 3)
#undef SK_BAD4
    }
}
__device__ __forceinline__ int max3i(int a, int b, int c) { return max(max(a, b), c); }
__device__ __forceinline__ int min3i(int a, int b, int c) { return min(min(a, b), c); }

// fasta_trim_by_quality.rs:28-48 for one record per lane, sixteen quality bytes per step: every lane
// walks its quality string down in aligned 16-byte blocks; the minimum key of a block gives the lowest
// total and its position at once (:38); in the block in which the running total first exceeds 0 (:37)
// only the totals before the break count.  Must be
// called by all 32 lanes (`has` = this lane carries a record).  min_baseq <= 222.
__device__ __forceinline__ bool plan_trim_lane16(const uint8_t *b, bool has, uint32_t L1, uint32_t L2, uint32_t L3, uint32_t L4,
                                                 int minq, uint8_t &mode, uint32_t &kk, uint32_t &body_len) {
    uint32_t k = has ? L4 - L3 : 0u;
#pragma unroll 1
    while (k > 0 && is_ws(b[L3 + k - 1])) k--;  // qual.trim_end().len()  (:31)
    __syncwarp();
    const uint32_t E = L3 + k;
    const int sub = 33 + minq;
    int total = -50, lowest = -50;  // :28-29
    uint32_t lowest_k = k;
    uint32_t a = k ? ((E - 1u) & ~15u) : 0u;
    bool active = k > 0;
#pragma unroll 1
    while (__any_sync(0xffffffffu, active)) {
        if (active) {
            int K[16];
            blk16_keys(b, a, L3, E, sub, total, K);
            const int mx = max3i(max3i(max3i(K[0], K[1], K[2]), max3i(K[3], K[4], K[5]), max3i(K[6], K[7], K[8])),
                                 max3i(K[9], K[10], K[11]), max3i(max3i(K[12], K[13], K[14]), K[15], K[15]));
            if (mx > 15) {  // the break (:37) is inside this block: its totals up to the break may still lower the minimum
                bool ok = true;
                int best = 0x7FFFFFFF;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    ok = ok && K[i] <= 15;  // totals after the break are never looked at
                    best = (ok && K[i] < best) ? K[i] : best;
                }
                if (best != 0x7FFFFFFF && (best >> 4) < lowest) {
                    lowest = best >> 4;
                    lowest_k = a + 15u - (uint32_t)(best & 15) - L3;
                }
                active = false;
            } else {
                const int mn = min3i(min3i(min3i(K[0], K[1], K[2]), min3i(K[3], K[4], K[5]), min3i(K[6], K[7], K[8])),
                                     min3i(K[9], K[10], K[11]), min3i(min3i(K[12], K[13], K[14]), K[15], K[15]));
                if ((mn >> 4) < lowest) {  // strict '<': an earlier block keeps a tie
                    lowest = mn >> 4;
                    lowest_k = a + 15u - (uint32_t)(mn & 15) - L3;
                }
                total = K[15] >> 4;
                if (a <= L3) active = false;
                else a -= 16;
            }
        }
    }
    __syncwarp();
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

// Bytes [lo, hi) of the aligned 16-byte block at window offset a keep their value, the others become
// `sub` (the byte whose contribution to the running total is zero); [lo, hi) = the block's part of the
// quality string [L3, E).
__device__ __forceinline__ uint4 blk16_keep(uint4 v, uint32_t a, uint32_t L3, uint32_t E, int sub) {
    const uint32_t lo = a < L3 ? L3 - a : 0u;
    const uint32_t hi = a + 16u > E ? (E > a ? E - a : 0u) : 16u;
    const uint32_t sub4 = (uint32_t)sub * 0x01010101u;
    auto keep = [&](uint32_t w, int q) {
        const int l = (int)lo - 4 * q, h = (int)hi - 4 * q;
        const uint32_t ml = l <= 0 ? 0xFFFFFFFFu : (l >= 4 ? 0u : 0xFFFFFFFFu << (8 * l));
        const uint32_t mh = h >= 4 ? 0xFFFFFFFFu : (h <= 0 ? 0u : 0xFFFFFFFFu >> (8 * (4 - h)));
        const uint32_t m = ml & mh;
        return (w & m) | (sub4 & ~m);
    };
    return make_uint4(keep(v.x, 0), keep(v.y, 1), keep(v.z, 2), keep(v.w, 3));
}
// Summary of one aligned 16-byte block of a quality string, relative to an entering total of 0, in the
// order the reference examines the bytes (highest address first, fasta_trim_by_quality.rs:33): keys
// 16 * T + i, T = the total after the i-th byte examined.  mx / mn = the largest / smallest key of the
// block (the smallest names the lowest total and, among equal totals, the byte examined first, :38),
// last = the key after all sixteen bytes.  The four words are summarised on their own (four dot products
// with constant accumulators, one min and one max of four) and then offset by the words before them.
// 7-bit bytes; c0..c3 = -(j+1) * 16 * sub + j.
struct Blk16 {
    int mx, mn, last;
};
__device__ __forceinline__ Blk16 blk16_summary(const uint4 v, int c0, int c1, int c2, int c3) {
#define SK_W4(w, k0, k1, k2, k3)                                 \
    const int k0 = (int)__dp4a(w, 0x10000000u, (uint32_t)c0);    \
    const int k1 = (int)__dp4a(w, 0x10100000u, (uint32_t)c1);    \
    const int k2 = (int)__dp4a(w, 0x10101000u, (uint32_t)c2);    \
    const int k3 = (int)__dp4a(w, 0x10101010u, (uint32_t)c3);
    SK_W4(v.w, a0, a1, a2, a3)
    SK_W4(v.z, b0, b1, b2, b3)
    SK_W4(v.y, d0, d1, d2, d3)
    SK_W4(v.x, e0, e1, e2, e3)
#undef SK_W4
    // a word's last key is 16 * sum + 3: the next word's keys start 16 * sum + 4 higher
    const int B1 = a3 + 1, B2 = B1 + b3 + 1, B3 = B2 + d3 + 1;
    Blk16 r;
    r.mx = max3i(max3i(a0, a1, max(a2, a3)), max3i(b0, b1, max(b2, b3)) + B1,
                 max(max3i(d0, d1, max(d2, d3)) + B2, max3i(e0, e1, max(e2, e3)) + B3));
    r.mn = min3i(min3i(a0, a1, min(a2, a3)), min3i(b0, b1, min(b2, b3)) + B1,
                 min(min3i(d0, d1, min(d2, d3)) + B2, min3i(e0, e1, min(e2, e3)) + B3));
    r.last = B3 + e3;
    return r;
}
// 0x80 in every byte position <=> none of the sixteen (7-bit) bytes is below '!'
__device__ __forceinline__ uint32_t blk16_ge33(const uint4 v) {
    const uint32_t A = 0x5F5F5F5Fu;
    return (v.x + A) & (v.y + A) & (v.z + A) & (v.w + A);
}

// fasta_trim_by_quality.rs:28-48 for one record per lane with block summaries that do not depend on each
// other: every lane goes down its quality string two aligned 16-byte blocks per step; a block's summary
// (blk16_summary) is relative to an entering total of 0, so the sixteen running totals of a block wait
// for nothing but the block's bytes, and the walk itself -- break (:37), minimum (:38), entering total of
// the next block -- is a handful of instructions per block.  The warp leaves the loop when every lane has
// met its break or the start of its string; only then does a lane that broke look into its break block,
// whose totals before the break may still lower the minimum.  Bytes below '!' take the wrapping u8
// subtraction (:35): a lane that saw one says so in `cold` and the caller repeats the record byte-wise.
// Must be called by all 32 lanes (`has` = this lane carries a record).  33 + min_baseq <= 127.
__device__ __forceinline__ bool plan_trim_blocks(const uint8_t *b, bool has, uint32_t L1, uint32_t L2, uint32_t L3, uint32_t L4,
                                                 int minq, uint8_t &mode, uint32_t &kk, uint32_t &body_len, bool &cold) {
    constexpr uint32_t FULL = 0xffffffffu, NONE = 0xFFFFFFFFu;
    uint32_t k = has ? L4 - L3 : 0u;
#pragma unroll 1
    while (k > 0 && is_ws(b[L3 + k - 1])) k--;  // qual.trim_end().len()  (:31)
    __syncwarp();
    const uint32_t E = L3 + k;
    const int sub = 33 + minq, s16 = 16 * sub;
    const int c0 = -s16, c1 = -2 * s16 + 1, c2 = -3 * s16 + 2, c3 = -4 * s16 + 3;
    const uint32_t a_bot = L3 & ~15u;
    uint32_t a = k ? ((E - 1u) & ~15u) : 0u;  // block examined next (a dead lane keeps re-reading its last block)
    int total = -50, lowest = -50;            // :28-29
    uint32_t lowest_k = k;
    uint32_t brk_a = NONE;
    int brk_total = 0;
    uint32_t ge33 = 0x80808080u;
    bool live = k > 0;
    // one block of the walk: `on` = the lane has reached this block
    auto walk = [&](bool on, uint32_t at, const Blk16 &s) {
        if (on) {
            const int e16 = 16 * total;
            if (e16 + s.mx > 15) {  // a total above 0 (:37) inside this block
                brk_a = at;
                brk_total = total;
                live = false;
            } else {
                const int cand = e16 + s.mn;
                if ((cand >> 4) < lowest) {  // strict '<': an earlier block keeps a tie
                    lowest = cand >> 4;
                    lowest_k = at + 15u - (uint32_t)(cand & 15) - L3;
                }
                total += (s.last - 15) >> 4;
                if (at == a_bot) live = false;
            }
        }
    };
#pragma unroll 1
    while (__any_sync(FULL, live)) {
        const bool two = live && a != a_bot;
        const uint32_t a2 = two ? a - 16u : a;
        uint4 v1 = *(const uint4 *)(b + a), v2 = *(const uint4 *)(b + a2);
        const bool edge = a + 16u > E || a2 < L3;
        if (__any_sync(FULL, live && edge)) {  // the two ends of the string only
            if (edge) {
                v1 = blk16_keep(v1, a, L3, E, sub);
                v2 = blk16_keep(v2, a2, L3, E, sub);
            }
        }
        ge33 &= live ? blk16_ge33(v1) & blk16_ge33(v2) : 0xFFFFFFFFu;  // (a dead lane's block is not masked)
        const Blk16 s1 = blk16_summary(v1, c0, c1, c2, c3), s2 = blk16_summary(v2, c0, c1, c2, c3);
        walk(live, a, s1);
        walk(live && two, a2, s2);
        if (live) a = a2 - 16u;
    }
    if (__any_sync(FULL, brk_a != NONE)) {
        if (brk_a != NONE) {  // the totals of the break block up to the break
            const uint4 v = blk16_keep(*(const uint4 *)(b + brk_a), brk_a, L3, E, sub);
            int K[16];
            int base = 16 * brk_total;
#define SK_KEY4(w, q)                                                                   \
    K[4 * q + 0] = (int)__dp4a(w, 0x10000000u, (uint32_t)(base + c0 + 4 * q));          \
    K[4 * q + 1] = (int)__dp4a(w, 0x10100000u, (uint32_t)(base + c1 + 4 * q));          \
    K[4 * q + 2] = (int)__dp4a(w, 0x10101000u, (uint32_t)(base + c2 + 4 * q));          \
    K[4 * q + 3] = (int)__dp4a(w, 0x10101010u, (uint32_t)(base + c3 + 4 * q));          \
    base = K[4 * q + 3] - (4 * q + 3);
            SK_KEY4(v.w, 0)
            SK_KEY4(v.z, 1)
            SK_KEY4(v.y, 2)
            SK_KEY4(v.x, 3)
#undef SK_KEY4
            bool ok = true;
            int best = 0x7FFFFFFF;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                ok = ok && K[i] <= 15;  // totals after the break are never looked at
                best = (ok && K[i] < best) ? K[i] : best;
            }
            if (best != 0x7FFFFFFF && (best >> 4) < lowest) {
                lowest = best >> 4;
                lowest_k = brk_a + 15u - (uint32_t)(best & 15) - L3;
            }
        }
    }
    __syncwarp();
    cold = k > 0 && (ge33 & 0x80808080u) != 0x80808080u;
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
// The two table probes of one class, issued (fidx_issue) as soon as the barcode bytes are in registers so
// that their latency is covered by whatever the caller does before it needs the match.
struct FProbe {
    uint32_t tag0, tag1, sl0, sl1;
    uint2 e0, e1;
};
template <int NR>
__device__ __forceinline__ FProbe fidx_issue(const uint32_t (&raw)[NR], const HalfIdx &H, const FastIdx &F, const uint32_t *hcls,
                                             uint32_t c) {
    constexpr int NWMAX = NR - 1;
    const uint32_t nw = H.nw, nwp = H.nwp, tmask = H.tsize - 1u;
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
    FProbe pr;
    pr.tag0 = tag[0];
    pr.tag1 = tag[1];
    pr.sl0 = (tag[0] ^ (tag[0] >> 15)) & tmask;
    pr.sl1 = (tag[1] ^ (tag[1] >> 15)) & tmask;
    pr.e0 = __ldg(&tab0[pr.sl0]);
    pr.e1 = __ldg(&tab1[pr.sl1]);
    return pr;
}
template <int NR>
__device__ __forceinline__ void fidx_match(const uint32_t (&raw)[NR], const HalfIdx &H, const FastIdx &F,
                                           const uint32_t *hcls, uint32_t S, uint32_t &lowest, uint32_t &best,
                                           uint32_t &last, const FProbe *first = nullptr) {
    constexpr int NWMAX = NR - 1;
    const uint32_t nw = H.nw, nwp = H.nwp, tmask = H.tsize - 1u;
    lowest = 0xFFFFFFFFu;
    best = 0xFFFFFFFFu;
    last = 0;
    for (uint32_t c = 0; c < H.n_classes; c++) {
        const uint32_t *care = hcls + c * HIDX_CLS_ROWS * nwp;
        const FProbe pr = (c == 0 && first) ? *first : fidx_issue<NR>(raw, H, F, hcls, c);  // class 0 may come pre-issued
        const uint32_t tag[2] = {pr.tag0, pr.tag1};
        const uint2 *tab0 = F.table + (size_t)(c * 2) * H.tsize, *tab1 = tab0 + H.tsize;
        uint32_t sl0 = pr.sl0, sl1 = pr.sl1;
        uint2 e0 = pr.e0, e1 = pr.e1;
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

// ------------------------------------------------------------------------------------------------
// window -> global memory copies (the emit of sk_warp.cu, the move of sk_compact.cu)
// ------------------------------------------------------------------------------------------------
// Up to eight bytes of the window from any byte offset, as aligned words and funnel shifts.  The window
// is addressed as base + offset throughout, so that the compiler keeps the accesses in the shared space
// (a pointer rebuilt from an integer turns them into generic loads).
__device__ __forceinline__ uint32_t lds_un32(const uint8_t *win, uint32_t off) {
    const uint32_t *w = (const uint32_t *)(win + (off & ~3u));
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}
__device__ __forceinline__ uint2 lds_un64(const uint8_t *win, uint32_t off) {
    const uint32_t *w = (const uint32_t *)(win + (off & ~3u));
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
// Sixteen bytes of the window from any byte offset.
__device__ __forceinline__ uint4 lds_unaligned16(const uint8_t *win, uint32_t off) {
    const uint32_t *w = (const uint32_t *)(win + (off & ~3u));
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
// Thirty-two bytes of the window from any byte offset >= -32.
struct U256 {
    uint32_t w[8];
};
__device__ __forceinline__ U256 lds_unaligned32(const uint8_t *win, int off) {
    const uint32_t *p = (const uint32_t *)(win + (off & ~3));
    const uint32_t sh = ((uint32_t)off & 3u) * 8u;
    U256 r;
    uint32_t lo = p[0];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t hi = p[k + 1];
        r.w[k] = __funnelshift_r(lo, hi, sh);
        lo = hi;
    }
    return r;
}
// bytes [0, f) of a, the others of b (0 <= f < 32)
__device__ __forceinline__ U256 merge_low32(const U256 &a, const U256 &b, uint32_t f) {
    // Three instructions a word: the bit position of byte f in word k, clamped below at 0 by the fused
    // add-max and above at 32 by shl (PTX clamps the shift amount), is where b takes over from a.
    U256 r;
    const int t = 8 * (int)f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int s = __viaddmax_s32(t, -32 * k, 0);
        uint32_t mb;
        asm("shl.b32 %0, %1, %2;" : "=r"(mb) : "r"(0xFFFFFFFFu), "r"(s));
        r.w[k] = (a.w[k] & ~mb) | (b.w[k] & mb);
    }
    return r;
}
// One 256-bit store to a 32-byte aligned global address (sm_100: STG.256).
__device__ __forceinline__ void stg256(void *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                       uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
                 "r"(f), "r"(g), "r"(h)
                 : "memory");
}
// Lane-serial copy of `len` window bytes from offset `so` to global memory, any alignment on either side.
// The destination is brought to a 16-byte boundary by at most one store of each size 1, 2, 4, 8 (no
// loops: the lanes of a warp copy runs of different alignment), one of 16 to reach a 32-byte sector
// boundary, then 32 bytes per step (8 LDS.32 + 8 funnel shifts + 1 STG.256), then at most one store of
// each size 16, 8, 4, 2, 1.  The window is only ever read as aligned words (up to seven bytes past the run).
__device__ __forceinline__ void gcopy(uint8_t *dst, const uint8_t *win, uint32_t so, uint32_t len) {
    if (((uint32_t)(uintptr_t)dst & 1u) && len >= 1u) {
        *dst = win[so];
        dst += 1, so += 1, len -= 1;
    }
    if (((uint32_t)(uintptr_t)dst & 2u) && len >= 2u) {
        *(uint16_t *)dst = (uint16_t)lds_un32(win, so);
        dst += 2, so += 2, len -= 2;
    }
    if (((uint32_t)(uintptr_t)dst & 4u) && len >= 4u) {
        *(uint32_t *)dst = lds_un32(win, so);
        dst += 4, so += 4, len -= 4;
    }
    if (((uint32_t)(uintptr_t)dst & 8u) && len >= 8u) {
        *(uint2 *)dst = lds_un64(win, so);
        dst += 8, so += 8, len -= 8;
    }
    // dst is 16-byte aligned here; one 16-byte step brings it to a 32-byte sector boundary, then whole
    // sectors go out with 256-bit stores (STG.256: half as many requests, none of them a partial sector)
    if ((((uint32_t)(uintptr_t)dst & 16u) && len >= 16u)) {
        *(uint4 *)dst = lds_unaligned16(win, so);
        dst += 16, so += 16, len -= 16;
    }
    if (len >= 32u) {
        const uint32_t sh = (so & 3u) * 8u;
        const uint32_t *sw = (const uint32_t *)(win + (so & ~3u));
        uint32_t lo = *sw;
        const uint32_t n32 = len >> 5;
#pragma unroll 1
        for (uint32_t i = 0; i < n32; i++) {
            const uint32_t w1 = sw[1], w2 = sw[2], w3 = sw[3], w4 = sw[4], w5 = sw[5], w6 = sw[6], w7 = sw[7], w8 = sw[8];
            stg256(dst, __funnelshift_r(lo, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                   __funnelshift_r(w3, w4, sh), __funnelshift_r(w4, w5, sh), __funnelshift_r(w5, w6, sh),
                   __funnelshift_r(w6, w7, sh), __funnelshift_r(w7, w8, sh));
            lo = w8;
            sw += 8;
            dst += 32;
        }
        so += n32 * 32u, len &= 31u;
    }
    if (len & 16u) {
        *(uint4 *)dst = lds_unaligned16(win, so);
        dst += 16, so += 16;
    }
    if (len & 8u) {
        if (((uint32_t)(uintptr_t)dst & 7u) == 0u) {
            *(uint2 *)dst = lds_un64(win, so);
        } else {  // (a run shorter than its head alignment)
#pragma unroll 1
            for (uint32_t i = 0; i < 8u; i++) dst[i] = win[so + i];
        }
        dst += 8, so += 8;
    }
    if (len & 4u) {
        if (((uint32_t)(uintptr_t)dst & 3u) == 0u) {
            *(uint32_t *)dst = lds_un32(win, so);
        } else {
#pragma unroll 1
            for (uint32_t i = 0; i < 4u; i++) dst[i] = win[so + i];
        }
        dst += 4, so += 4;
    }
    if (len & 2u) {
        if (((uint32_t)(uintptr_t)dst & 1u) == 0u) {
            *(uint16_t *)dst = (uint16_t)lds_un32(win, so);
        } else {
            dst[0] = win[so];
            dst[1] = win[so + 1];
        }
        dst += 2, so += 2;
    }
    if (len & 1u) *dst = win[so];
}

// ------------------------------------------------------------------------------------------------
// global -> global copy of one piece by a whole warp (sk_compact.cu's fallback, sk_lineops.cu's emit)
// ------------------------------------------------------------------------------------------------
// len bytes from src + so to dst + d_o (any alignment on either side), the whole warp: a lane takes the
// 16-byte units of the destination, reads the two aligned 16-byte pieces of the source that hold a unit's
// bytes and shifts them into place (the shift is the same for every unit of a piece).
__device__ __forceinline__ void warp_copy_piece(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, unsigned long long so,
                                                unsigned long long d_o, uint32_t len, int lane) {
    uint8_t *d = dst + d_o;
    const uint32_t a = (uint32_t)(uintptr_t)d & 15u;  // d - a is 16-byte aligned
    const uint32_t units = (a + len + 15u) >> 4;
    const long long s0 = (long long)so - (long long)a;  // source offset of the first unit's first byte (may be < 0)
    const uint32_t sh = (uint32_t)(s0 & 15), wo = sh >> 2, bs = (sh & 3u) * 8u;
    for (uint32_t u = (uint32_t)lane; u < units; u += 32u) {
        const long long su = (s0 + 16ll * u) & ~15ll;  // aligned piece that holds the unit's first byte
        uint4 A = make_uint4(0u, 0u, 0u, 0u), B = make_uint4(0u, 0u, 0u, 0u);
        if (su >= 0) A = *(const uint4 *)(src + su);
        if (sh && su + 16 >= 0) B = *(const uint4 *)(src + su + 16);  // (the buffers end with slack: reading past a piece is fine)
        const uint32_t W[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
        uint4 o4;
        switch (wo) {  // uniform over the piece
            case 0: o4 = make_uint4(__funnelshift_r(W[0], W[1], bs), __funnelshift_r(W[1], W[2], bs), __funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs)); break;
            case 1: o4 = make_uint4(__funnelshift_r(W[1], W[2], bs), __funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs)); break;
            case 2: o4 = make_uint4(__funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs), __funnelshift_r(W[5], W[6], bs)); break;
            default: o4 = make_uint4(__funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs), __funnelshift_r(W[5], W[6], bs), __funnelshift_r(W[6], W[7], bs)); break;
        }
        uint8_t *du = d - a + 16u * u;
        const uint32_t first = u == 0 ? a : 0u;  // valid bytes of the unit: [first, last)
        const uint32_t last = 16u * u + 16u > a + len ? a + len - 16u * u : 16u;
        if (first == 0u && last == 16u) {
            *(uint4 *)du = o4;
        } else {  // the piece's first and last unit are shared with its neighbours in the sample's run
            const uint32_t ow[4] = {o4.x, o4.y, o4.z, o4.w};
            for (uint32_t i = first; i < last; i++) du[i] = (uint8_t)(ow[i >> 2] >> (8u * (i & 3u)));
        }
    }
}

}  // namespace sk
