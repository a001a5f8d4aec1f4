// sk_warp.cu -- the warp engine: header-route demultiplex (mate 1 / mate 2, with or without the fused
// quality trim), trim by quality and mask by quality, with one warp per tile and one lane per record
// (DESIGN.md section 3.0).
//
// A CTA-per-chunk engine (round 1's lean engine, retired) spends a third of its warp time at CTA barriers:
// its per-record phase runs one lane per record on the few warps a 16 KiB chunk fills, while the other warps
// of the CTA wait.  Here a warp owns a whole tile of the stream from load to store (the warps of a CTA only
// start their tiles together, for the instruction cache):
//   * tile = 29 lanes x 400 B of input (about 31 records of 2x150 bp FASTQ) + 3 lanes of overhang,
//     loaded by one TMA bulk copy into the warp's private window; 16 such warps per SM sit in
//     different phases, so nobody waits at a barrier and every per-record step has its 32 lanes busy;
//   * newline scan: 25 conflict-free LDS.128 per lane, newline maps in registers, one shuffle scan,
//     line starts written by predicated stores;
//   * record framing by global line index (common.rs:106-112): tiles publish their line counts as
//     16-bit words, a tile sums the 1024 counts before it with four 16-byte loads per lane and adds
//     the inclusive prefix of the tile before those (wlb_consume); the framing is guessed from the
//     text first and the guess verified before anything is written;
//   * per record, in registers: '@' check, leftmost " BC:x", class run, pigeonhole match, header
//     surgery, quality trim (fasta_demultiplex.rs:117-212, fasta_trim_by_quality.rs:28-48);
//   * output: a round (<= 32 records) takes its space with one atomicAdd, the record's edits are
//     patched into the window in place (" UMI:x\n" over the deleted " BC:x", "\n+\n" and "\n" behind
//     the kept bases and qualities), so that a record is two or three contiguous runs which the lane
//     copies straight to global memory with 16-byte stores.  No staging image, no second pass.
// Trim and mask by quality write one stream in input order: mask through a second look-back (on output
// bytes, known right after the line table); trim, whose sizes are known only after the plan, lets its
// tiles write wherever the output cursor puts them and restores the order with a scan and a gather
// kernel (bottom of this file).
// Anything outside the engine's limits (a record longer than the overhang, more than 128 records in
// a tile) raises F_NEED_GENERAL and sk_wait re-runs the operator on the general engine.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "sk_internal.h"

namespace sk {
extern __shared__ __align__(128) unsigned char sk_smem[];
}
#include "sk_device.cuh"
#include "sk_record.cuh"

#ifndef SKW_LOCKSTEP
#define SKW_LOCKSTEP 1  // 0: every warp takes its own tickets
#endif
#ifndef SKW_GROUP_WARPS
#define SKW_GROUP_WARPS 8  // warps that take consecutive tiles and start them together (8 = the CTA, or 4)
#endif
#ifndef SKW_EMIT_RUNS
#define SKW_EMIT_RUNS 1  // sector emit run by run (0: one piece per step, every step through the run table)
#endif
#ifndef SKW_EARLY_PROBE
#define SKW_EARLY_PROBE 0  // mate 1: 1 = barcode search, hash and table probes before the quality trim (measured: no gain)
#endif
#ifndef SKW_ADDBC_LS
#define SKW_ADDBC_LS 0  // add barcode: 1 = the warps of a CTA start their tiles together
#endif
#ifndef SKW_ADDBC_NOLOAD
#define SKW_ADDBC_NOLOAD 0  // timing experiment only (wrong output): add barcode without the loads of the barcode table and bytes
#endif
#ifndef SKW_TRIM_FREE
#define SKW_TRIM_FREE 1  // trim by quality: 1 = every warp on its own ticket (like mask; 2.68 vs 3.03 ms per 8 M reads)
#endif
#ifndef SKW_TRIM_BLOCKS
#define SKW_TRIM_BLOCKS 0  // quality trim with independent 16-byte block summaries (0: plan_trim_lane16)
#endif

namespace sk {

struct WLayout {
    static constexpr uint32_t win = 32;                                              // the sector emit reads up to 31 bytes below a run
    static constexpr uint32_t ls = win + GeoW::WIN + 32;                             // patches may touch win[wlen]
    static constexpr uint32_t misc = ls + (((GeoW::MAXLINES + 8) * 2 + 15) / 16) * 16;
    static constexpr uint32_t per_warp = misc + 16;
    static constexpr uint32_t lut = GeoW::WARPS * per_warp;
    static constexpr uint32_t dyn = lut + 256;  // hcls rows, per-sample counters, UMI lengths
};
static_assert(WLayout::per_warp % 16 == 0, "warp areas are 16-byte aligned");

// min_baseq > 222 (never in practice): the byte-wise trim of sk_device.cuh, out of line.  Returns
// kk | mode << 16 | fine << 24 (no reference parameters: they would pin the caller's variables to the stack).
static __device__ __noinline__ uint32_t plan_trim_cold(const uint8_t *b, uint32_t L1, uint32_t L2, uint32_t L3, uint32_t L4, int minq) {
    uint8_t mode = B_FAIL;
    uint32_t kk = 0, body_len = 0;
    const bool fine = plan_trim_body(b, L1, L2, L3, L4, minq, mode, kk, body_len);
    return (kk & 0xFFFFu) | ((uint32_t)mode << 16) | (fine ? 1u << 24 : 0u);
}

// fasta_trim_by_quality.rs:28-48 for one record per lane (all 32 lanes call; `ok` = this lane carries a
// record): block summaries in the usual case, the byte-wise form for a lane whose quality string holds a
// byte below '!' (wrapping u8 subtraction, :35) and for thresholds outside the block form's range.
__device__ __forceinline__ bool plan_trim_any(const uint8_t *win, bool ok, uint32_t L1, uint32_t L2, uint32_t L3, uint32_t L4,
                                              int trim_q, uint8_t &mode, uint32_t &kk, uint32_t &body) {
    bool fine = true, cold = ok;
#if SKW_TRIM_BLOCKS
    if (trim_q <= 94) fine = plan_trim_blocks(win, ok, L1, L2, L3, L4, trim_q, mode, kk, body, cold);
#else
    if (trim_q <= 222) {
        fine = plan_trim_lane16(win, ok, L1, L2, L3, L4, trim_q, mode, kk, body);
        cold = false;
    }
#endif
    if (__any_sync(0xffffffffu, cold)) {
        if (cold) {
            const uint32_t pk = plan_trim_cold(win, L1, L2, L3, L4, trim_q);
            kk = pk & 0xFFFFu;
            mode = (uint8_t)(pk >> 16);
            fine = (pk >> 24) != 0u;
            body = mode == B_GARBAGE ? 6u : 2u * kk + 4u;
        }
        __syncwarp();
    }
    return fine;
}

// Look-back of the warp engine over p.tile_lines: inc[c] (u64: bit 63 | lines through tile c) and, behind
// those, agg[c] (u16: lines owned by tile c, plus one; 0 = not yet known).  A tile's exclusive prefix is
// the inclusive prefix of tile 8(b-128)-1 plus the 1024 + (c & 7) counts after it, b = c / 8: four aligned
// 16-byte loads per lane.  The window is that wide because the inclusive prefixes form a chain (tile c
// needs the one of tile c-1024 or so): with n tiles the chain has n/1024 links of a few microseconds of
// memory latency each, which must stay far below the kernel's run time.  Tiles are handed out by a ticket counter, so every predecessor is owned by a
// running warp that never waits on a later tile.
__device__ __forceinline__ uint16_t *wlb_agg(uint64_t *tile_lines, uint32_t n_tiles) {
    return (uint16_t *)(tile_lines + ((n_tiles + 1u) & ~1u));
}
__device__ __forceinline__ void wlb_publish(uint16_t *agg, uint32_t c, uint32_t count) {
    const unsigned short v = (unsigned short)(count + 1u);
    asm volatile("st.relaxed.gpu.global.u16 [%0], %1;" ::"l"(agg + c), "h"(v) : "memory");
}
__device__ __forceinline__ uint32_t zero_half(uint32_t x) { return (x - 0x00010001u) & ~x & 0x80008000u; }
static __device__ __noinline__ uint64_t wlb_consume(uint64_t *inc, const uint16_t *agg, uint32_t c, uint32_t own, int lane) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t BPL = 4;           // blocks of eight counts per lane: a window of 1024 tiles
    constexpr uint32_t NB = 32 * BPL;
    uint64_t excl = 0;
    const uint32_t b = c >> 3;  // block of eight counts that holds c
    if (b >= NB) {
        const uint16_t *blk = agg + 8u * (b - NB + BPL * (uint32_t)lane);
        const uint16_t *part = agg + 8u * b;
        const uint32_t npart = c & 7u;
        uint32_t sum;
        for (;;) {
            uint32_t x[4 * BPL];
#pragma unroll
            for (uint32_t q = 0; q < BPL; q++)
                asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(x[4 * q]), "=r"(x[4 * q + 1]), "=r"(x[4 * q + 2]), "=r"(x[4 * q + 3])
                             : "l"(blk + 8u * q)
                             : "memory");
            uint32_t z = 0;
            sum = 0;
#pragma unroll
            for (uint32_t q = 0; q < 4 * BPL; q++) {
                z |= zero_half(x[q]);
                sum += (x[q] & 0xFFFFu) + (x[q] >> 16);
            }
            sum -= 8u * BPL;
            bool ok = z == 0u;
            if (lane == 0 && npart) {
                uint32_t y[4];
                asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]) : "l"(part) : "memory");
#pragma unroll
                for (int k = 0; k < 7; k++)
                    if ((uint32_t)k < npart) {
                        const uint32_t e = (y[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                        ok = ok && e != 0u;
                        sum += e - 1u;
                    }
            }
            if (__all_sync(FULL, ok)) break;
            __nanosleep(40);
        }
        sum = __reduce_add_sync(FULL, sum);
        const uint32_t t = 8u * (b - NB);
        uint64_t w = 1ull << 63;
        if (t > 0u && lane == 0) {
            w = ts_load(&inc[t - 1u]);
            while (!(w >> 63)) {
                __nanosleep(40);
                w = ts_load(&inc[t - 1u]);
            }
        }
        w = __shfl_sync(FULL, w, 0);
        excl = (w & ~(1ull << 63)) + sum;
    } else {
        uint32_t sum = 0;
#pragma unroll 1
        for (uint32_t i = (uint32_t)lane; i < c; i += 32u) {
            unsigned short e;
            for (;;) {
                asm volatile("ld.relaxed.gpu.global.u16 %0, [%1];" : "=h"(e) : "l"(agg + i) : "memory");
                if (e) break;
                __nanosleep(40);
            }
            sum += (uint32_t)e - 1u;
        }
        excl = __reduce_add_sync(FULL, sum);
    }
    if (lane == 0) {
        const uint64_t v = (1ull << 63) | (excl + own);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(inc + c), "l"(v) : "memory");
    }
    return excl;
}

template <int OP, int NWMAX>
__global__ void __launch_bounds__(GeoW::NT, GeoW::MIN_CTAS) sk_warp_kernel(const __grid_constant__ KParams p) {
    constexpr bool D1 = OP == OP_DEMUX1;
    constexpr bool IS_DEMUX = OP == OP_DEMUX1 || OP == OP_DEMUX2;  // else OP_TRIM / OP_MASK / OP_ADDBC: one output stream, input order
    static_assert(IS_DEMUX || OP == OP_TRIM || OP == OP_MASK || OP == OP_ADDBC, "warp engine: demultiplex passes, trim, mask, add barcode");
    constexpr int UPL = GeoW::UPL, LANE_BYTES = GeoW::LANE_BYTES, WIN = GeoW::WIN;
    constexpr int MAXREC = GeoW::MAXREC, MAXLINES = GeoW::MAXLINES;
    constexpr uint32_t FULL = 0xffffffffu;
    using WL = WLayout;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *warea = sk_smem + (uint32_t)warp * WL::per_warp;
    uint8_t *win = warea + WL::win;
    uint16_t *ls = (uint16_t *)(warea + WL::ls);
    uint64_t *mbar = (uint64_t *)(warea + WL::misc);
    uint8_t *sh_lut = sk_smem + WL::lut;
    uint32_t *hcls = (uint32_t *)(sk_smem + WL::dyn);
    const uint32_t S = p.sheet.S;
    const uint32_t hcls_words = D1 ? p.sheet.hidx.n_classes * HIDX_CLS_ROWS * p.sheet.hidx.nwp : 0u;
    uint32_t *ccount = hcls + ((hcls_words + 3u) & ~3u);
    const bool cc_smem = D1 && S <= (uint32_t)FAST_CCOUNT_MAX;
    uint8_t *sh_ulen = (uint8_t *)(ccount + (cc_smem ? ((S + 3u) & ~3u) : 0u));  // UMI length of every sample
    DevStats *st = p.stats;

    if (lane == 0) mbar_init(mbar, 1);
    if (IS_DEMUX) {
#pragma unroll 1
        for (uint32_t i = tid; i < 256; i += GeoW::NT) sh_lut[i] = p.sheet.lut[i];
#pragma unroll 1
        for (uint32_t i = tid; i < S; i += GeoW::NT)
            sh_ulen[i] = (uint8_t)(p.sheet.wide ? __popcll(((const unsigned long long *)p.sheet.umask)[i]) : __popc(p.sheet.umask[i]));
    }
    if (D1) {
#pragma unroll 1
        for (uint32_t i = tid; i < hcls_words; i += GeoW::NT) hcls[i] = p.sheet.hidx.cls[i];
        if (cc_smem) {
#pragma unroll 1
            for (uint32_t i = tid; i < S; i += GeoW::NT) ccount[i] = 0;
        }
    }
    constexpr int GW = SKW_GROUP_WARPS;  // warps of a group; its first warp is the group's leader
    static_assert(GW == GeoW::WARPS || GW == 4, "groups of 4 warps, or the whole CTA");
    const bool g_lead = (tid % (GW * 32)) == 0;
    const int wg = warp % GW;
    if (g_lead) {  // the group's ticket slots (below): 0 is never newer than a ticket
        ((volatile uint32_t *)(warea + WL::misc + 8))[0] = 0u;
        ((volatile uint32_t *)(warea + WL::misc + 8))[1] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t tile_lanes = p.tile_lanes;
    const uint32_t TILE = tile_lanes * (uint32_t)LANE_BYTES;
    const bool fused = IS_DEMUX && p.fused_trim >= 0;
    const int trim_q = IS_DEMUX ? p.fused_trim : (int)p.min_baseq;
    uint16_t *oagg16 = wlb_agg(p.tile_out, p.n_chunks);  // trim / mask: second look-back, on output bytes
    const uint32_t Lb = p.sheet.L;
    uint16_t *agg16 = wlb_agg(p.tile_lines, p.n_chunks);
    const unsigned long long r1_nrec = (D1 || !IS_DEMUX) ? 0ull : p.r1_stats->n_records;  // mate 2: records of the mate-1 pass
    // add barcode: records of the barcode stream's table, and the growth of a record when every barcode is as long
    // as the first and no header ends in white space (" BC:" + barcode; the in-place form's assumption)
    const unsigned long long bc_n = (OP == OP_ADDBC && p.ext_stats[0]) ? p.ext_stats[0]->n_records : 0ull;
    const uint32_t addD = 4u + (bc_n ? (uint32_t)p.ext_tab[0][0].seq_len : 0u);
    unsigned long long my_outmax = 0;      // lane 0, add barcode in place: end of the output of this warp's tiles
    uint32_t parity = 0;
    uint32_t my_total = 0, my_ident = 0;   // DEMUX1 counters of this lane's records
    unsigned long long my_out = 0;         // lane 0: payload bytes of this warp's tiles
    unsigned long long my_nrec = 0, my_upto = 0;  // lane 0: records of this warp's tiles, end of the last one
#define LB(x) ((uint32_t)ls[(x)])

    // Starts the load of tile cc into the warp's window (TMA bulk copy; plain loads for the ragged tail of
    // the stream).  The window must be dead: called at the top of a tile or, when the CTA's next ticket is
    // already known, right after the previous tile's emit -- the copy then runs while the warp waits for
    // its CTA at the meeting point.
    auto issue_load = [&](uint32_t cc) {
        const uint64_t b0 = (uint64_t)cc * TILE;
        uint64_t bend = b0 + (uint64_t)WIN;
        if (bend > p.n) bend = p.n;
        const uint32_t blen = (uint32_t)(bend - b0);
        const uint32_t bbulk = blen & ~15u;
        fence_proxy_async();  // this lane's patches of the previous window before the next TMA write
        __syncwarp();         // every lane is done with the previous window
        if (lane == 0 && bbulk) {
            fence_proxy_async();
            mbar_expect_tx(mbar, bbulk);
            bulk_g2s(win, p.in + b0, bbulk, mbar);
        }
        if (bbulk != (uint32_t)WIN) {  // last window of the stream: ragged tail, zeros up to the end of the window
            if (lane < 16) {
                const uint32_t o = bbulk + (uint32_t)lane;
                win[o] = (o < blen) ? p.in[b0 + o] : (uint8_t)0;
            }
            for (uint32_t o = bbulk + 16u + 16u * (uint32_t)lane; o < (uint32_t)WIN + 32u; o += 512u)
                *(uint4 *)(win + o) = make_uint4(0u, 0u, 0u, 0u);
        }
    };
    bool early = false;  // the load of this tile was started at the end of the previous one

    uint32_t c = 0;
    // Demultiplex and trim: the warps of a CTA take eight consecutive tiles at a time and start them
    // together: they then run the same code at about the same time, which the SM's instruction cache needs
    // (the kernels are larger than that cache, and sixteen warps spread over them saturate the GPC-level
    // instruction cache).  Mask (a small kernel whose tiles wait on each other's output sizes) runs faster
    // with every warp on its own ticket, taken when the tile starts.
    constexpr bool LS = SKW_LOCKSTEP && OP != OP_MASK && !(!SKW_ADDBC_LS && OP == OP_ADDBC) && !(SKW_TRIM_FREE && OP == OP_TRIM);
    volatile uint32_t *cta_ticket =
        (volatile uint32_t *)(sk_smem + (uint32_t)(warp - wg) * WL::per_warp + WL::misc + 8);  // two slots in the leader's misc area
    uint32_t cta_next = 0, flipk = 0;
    bool cta_have = false;
    if (!LS) {
        if (lane == 0) c = atomicAdd(&st->ticket, 1u);
        c = __shfl_sync(FULL, c, 0);
    }
    for (;;) {
        uint32_t cb = 0;
        if (LS) {
            if (g_lead) cta_ticket[flipk] = cta_have ? cta_next : atomicAdd(&st->ticket, (uint32_t)GW);
            cta_have = false;
            if (GW == GeoW::WARPS) __syncthreads();
            else asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / GW), "r"(GW * 32) : "memory");
            cb = cta_ticket[flipk];
            flipk ^= 1u;
            if (cb >= p.n_chunks) break;
            c = cb + (uint32_t)wg;
            if (c >= p.n_chunks) continue;
        } else if (c >= p.n_chunks) {
            break;
        }
        const uint64_t c0 = (uint64_t)c * TILE;
        uint64_t wend = c0 + (uint64_t)WIN;
        if (wend > p.n) wend = p.n;
        const uint32_t wlen = (uint32_t)(wend - c0);
        const bool at_end = (wend == p.n);

        // ---- load the window
        const uint32_t bulk = wlen & ~15u;
        if (!early) issue_load(c);
        early = false;
        if (bulk) {
            mbar_wait_parked(mbar, parity);
            parity ^= 1;
        }
        __syncwarp();

        // ---- newline scan.  A '\n' at window offset q starts a line at q+1; the tile owns the line
        // starts of the newlines inside its TILE bytes (and the line at byte 0 of the stream).  A '\n'
        // that is the last byte of the stream starts nothing.
        const uint32_t ls_hi = at_end ? (wlen ? wlen - 1 : 0) : wlen;
        const uint32_t o0 = (uint32_t)lane * LANE_BYTES;
        // 25 conflict-free LDS.128 per lane; the newline maps stay in registers (13 words of 32 bytes each).
        constexpr int NMW = (UPL + 1) / 2;
        uint32_t mw[NMW];
        uint32_t hib = 0;
#pragma unroll
        for (int i = 0; i < NMW; i++) mw[i] = 0;
#pragma unroll
        for (int q = 0; q < UPL; q++) {
            const uint4 v = *(const uint4 *)(win + o0 + q * 16);
            hib |= v.x | v.y | v.z | v.w;
            mw[q >> 1] |= nl_map_nat(v) << (16 * (q & 1));
        }
        if (o0 + LANE_BYTES > ls_hi) {  // last window of the stream only
#pragma unroll
            for (int i = 0; i < NMW; i++) mw[i] &= bits_below((int)ls_hi - (int)(o0 + 32u * i));
        }
        if (__any_sync(FULL, (hib & 0x80808080u) != 0) && lane == 0) atomicOr(&st->flags, F_NON_ASCII);
        uint32_t cnt_all = 0;
#pragma unroll
        for (int i = 0; i < NMW; i++) cnt_all += __popc(mw[i]);
        const uint32_t cnt_own = (uint32_t)lane < tile_lanes ? cnt_all : 0u;
        uint32_t incl = (cnt_own << 16) | cnt_all;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
        }
        const uint32_t tot = __shfl_sync(FULL, incl, 31);
        const uint32_t extra = (c == 0) ? 1u : 0u;
        const uint32_t nls = (tot & 0xFFFFu) + extra;
        const uint32_t nls_own = (tot >> 16) + extra;
        if (lane == 0) wlb_publish(agg16, c, nls_own);
        {
            // line starts: the first two newlines of every 32-byte word by predicated stores (a third one in
            // 32 bytes is rare), no branch in the common case
            uint32_t idx = ((incl & 0xFFFFu) - cnt_all) + extra;
            if (lane == 0 && extra) ls[0] = 0;
            uint32_t base = o0;
#pragma unroll
            for (int wi = 0; wi < NMW; wi++) {
                uint32_t m = mw[wi];
                const bool h1 = m != 0u;
                const uint32_t p1 = base + (uint32_t)__ffs((int)m);
                if (h1 && idx < (uint32_t)MAXLINES) ls[idx] = (uint16_t)p1;
                idx += h1 ? 1u : 0u;
                m &= m - 1;
                const bool h2 = m != 0u;
                const uint32_t p2 = base + (uint32_t)__ffs((int)m);
                if (h2 && idx < (uint32_t)MAXLINES) ls[idx] = (uint16_t)p2;
                idx += h2 ? 1u : 0u;
                m &= m - 1;
                if (__any_sync(FULL, m != 0u)) {
#pragma unroll 1
                    while (m) {
                        if (idx < (uint32_t)MAXLINES) ls[idx] = (uint16_t)(base + (uint32_t)__ffs((int)m));
                        idx++;
                        m &= m - 1;
                    }
                }
                base += 32u;
            }
            if (lane < 8) {  // sentinels: lines past the last line start read as "end of window"
                const uint32_t k = nls + (uint32_t)lane;
                if (k < (uint32_t)MAXLINES + 8u) ls[k] = (uint16_t)wlen;
            }
        }
        __syncwarp();

        // ---- framing.  Record i is lines 4i..4i+3 of the stream (common.rs:106-112): the tile needs the
        // global index g0 of its first line, i.e. the counts of every tile before it -- and the nearest of
        // those are published only moments before this point is reached.  The framing is therefore guessed
        // from the text (a line that starts with '@' whose second successor starts with '+'), the first
        // round is planned on the guess, and the guess is checked against the look-back before the round
        // writes anything; a wrong guess (malformed or adversarial text only) repeats the plan.
        bool spec = false;
        uint32_t j0 = 0;
        uint64_t g0 = 0;
        if (c != 0 && p.rec_limit == ~0ull && nls >= 6u) {
            const uint32_t s0 = LB(0), s1 = LB(1), s2 = LB(2), s3 = LB(3), s4 = LB(4), s5 = LB(5);
            const uint32_t b0 = win[s0], b1 = win[s1], b2 = win[s2], b3 = win[s3], b4 = win[s4], b5 = win[s5];
            const bool k0 = b0 == '@' && b2 == '+', k1 = b1 == '@' && b3 == '+';
            const bool k2 = b2 == '@' && b4 == '+', k3 = b3 == '@' && b5 == '+';
            j0 = k0 ? 0u : k1 ? 1u : k2 ? 2u : 3u;
            spec = (k0 || k1 || k2 || k3) && j0 < nls_own;
        }
        if (!spec) {
            g0 = wlb_consume(p.tile_lines, agg16, c, nls_own, lane);
            j0 = (4u - (uint32_t)(g0 & 3u)) & 3u;
        }

        uint32_t nrec = 0, c_next = 0;
        bool bail = false, have_next = false;
        bool out_done = false;  // trim / mask: this tile's output bytes are in the second look-back
        for (;;) {  // repeated only when the guess was wrong
            nrec = j0 < nls_own ? (nls_own - 1 - j0) / 4u + 1u : 0u;
            if (!spec) {
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

            // trim / mask by quality for one record per lane (fasta_trim_by_quality.rs:19-48,
            // fasta_mask_by_quality.rs:20-45): returns the record's output length, failure kind in errk
            auto stream_plan = [&](bool has, uint32_t L0, uint32_t L1, uint32_t L2, uint32_t L3, uint32_t L4, uint8_t &mode,
                                   uint32_t &kk, uint32_t &errk) -> uint32_t {
                uint32_t slen = 0;
                errk = 0;
                mode = B_NONE;
                if (OP == OP_TRIM) {
                    bool ok = has;
                    if (has && win[L0] != '@') {  // :20-22
                        errk = K_BAD_HEADER;
                        ok = false;
                    }
                    bool fine = true;
                    uint32_t body = 0;
                    fine = plan_trim_any(win, ok, L1, L2, L3, L4, trim_q, mode, kk, body);
                    if (!ok) {
                        mode = B_NONE;
                    } else if (!fine) {  // &seq[..k] would panic (:47)
                        errk = K_SEQ_SHORT;
                        mode = B_NONE;
                    } else {
                        slen = (L1 - L0) + body;  // header verbatim (:23) + body
                    }
                }
                if (OP == OP_MASK && has) {
                    if (win[L0] != '@') {  // fasta_mask_by_quality.rs:21-23
                        errk = K_BAD_HEADER;
                    } else {
                        uint32_t sl = L2 - L1, ql = L4 - L3;
                        if (sl && win[L2 - 1] == '\n') sl--;  // :32
                        if (ql && win[L4 - 1] == '\n') ql--;  // :33
                        if (sl != ql) {                       // :35-37
                            errk = K_LEN_MISMATCH;
                        } else {
                            mode = B_MASK;
                            kk = sl;
                            slen = (L1 - L0) + 2u * sl + 4u;  // header, masked, "\n+\n", qual, "\n"  (:26,:44)
                        }
                    }
                }
                return slen;
            };
            // add barcode for one record per lane (fasta_add_barcode.rs:20-43): the record's output length; the barcode of
            // record `rec` is the sequence line of barcode record `rec`, the last one once that file is exhausted
            auto addbc_plan = [&](bool has, uint64_t rec, uint32_t L0, uint32_t L1, uint32_t L4, uint32_t &alen, uint32_t &bl,
                                  uint32_t &bo, uint32_t &errk) -> uint32_t {
                errk = 0;
                alen = bl = bo = 0;
                if (!has) return 0u;
                uint32_t e = L1;
                while (e > L0 && is_ws(win[e - 1])) e--;  // header.trim_end()  (:33)
                alen = e - L0;
                if (bc_n) {
                    const RecRef rr = p.ext_tab[0][rec < bc_n ? rec : bc_n - 1ull];
                    bl = rr.seq_len;
                    bo = rr.seq_off;
                    if (rr.flags & RR_LONG) {
                        errk = K_TOO_LONG;
                        return 0u;
                    }
                }
                const uint8_t h = win[L0];
                if (h != '@') {  // the reference prints the BC'd header and stops (:33 before :41-43): the host reproduces that
                    errk = h == '>' ? K_MIXED : K_BAD_FASTX_LINE;
                    return 0u;
                }
                return alen + 4u + bl + 1u + (L4 - L1);
            };
            uint32_t s_l0 = 0;      // add barcode in place: window offset of the tile's first record
            uint64_t s_obase = 0;   // trim / mask: the tile's place in the output stream ...
            uint32_t s_done = 0;    // ... and the bytes its earlier rounds have written
            bool s_writable = false;
            bool wrong = false;
            uint64_t rec0 = 0;
            for (uint32_t r0 = 0; r0 < nrec; r0 += 32u) {
                // ---- plan: one lane per record, nothing is written
                const uint32_t r = r0 + (uint32_t)lane;
                const bool has = r < nrec;
                const uint32_t j = j0 + (has ? r : 0u) * 4u;
                const uint32_t L0 = LB(j), L1 = LB(j + 1), L2 = LB(j + 2), L3 = LB(j + 3), L4 = LB(j + 4);
                uint8_t mode = B_VERBATIM;
                uint32_t kk = 0, body = L4 - L1;  // three lines verbatim (:209-212)
                int sample = -1;
                unsigned long long um = 0;  // positions where the sample's sheet barcode has 'U'
                uint32_t alen = 0, blen = 0, cut0 = 0, cut1 = 0, taglen = 0xFFu;
                // mate 1, first half (fasta_demultiplex.rs:117-150): validate, locate the barcode, hash its two
                // halves and send the two table probes on their way -- they come back while the quality trim runs
                bool live = D1 && has;
                uint32_t raw[NWMAX + 1];
                FProbe probe;
                uint32_t bs = 0;
                auto d1_locate = [&]() {
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
                    bs = stp + 4;
                    if (live) {
                        cut0 = stp - L0;
                        cut1 = cut0 + 4 + Lb;
                        load_raw<NWMAX + 1>(win, bs, (Lb + 4u) >> 2, raw);
                        probe = fidx_issue<NWMAX + 1>(raw, p.sheet.hidx, p.sheet.fidx, hcls, 0u);  // consumed after the class check
                    }
                    __syncwarp();
                };
                if (D1 && SKW_EARLY_PROBE) d1_locate();
                if (fused) {
                    const bool ok = has && L1 > L0 && win[L1 - 1] == '\n';
                    bool fine;
                    mode = B_FAIL;
                    fine = plan_trim_any(win, ok, L1, L2, L3, L4, trim_q, mode, kk, body);
                    if (!ok || !fine) mode = B_FAIL;
                }
                // trim / mask by quality: failure kind in errk, output length in slen
                uint32_t errk = 0, slen = 0;
                if (!IS_DEMUX && OP != OP_ADDBC) slen = stream_plan(has, L0, L1, L2, L3, L4, mode, kk, errk);
                if (!IS_DEMUX) {
                } else if (D1) {
                    // fasta_demultiplex.rs:148-194, second half: barcode length, match, decide.  Outcome:
                    // sample >= 0 assigned; -2 ambiguous (best, last, mismatches in alen, blen, taglen);
                    // -1 unassigned (taglen = failure kind, or 0xFF for a record that never reached the match)
                    if (!SKW_EARLY_PROBE) d1_locate();
                    if (live) {
                        if (!class_run_is<NWMAX + 1>(raw, sh_lut, Lb, L1 - bs)) {  // :38, :148-150
                            taglen = K_BC_LEN;
                            live = false;
                        }
                    }
                    __syncwarp();
                    if (live) {  // :154-194
                        uint32_t lowest, best, last;
                        fidx_match<NWMAX + 1>(raw, p.sheet.hidx, p.sheet.fidx, hcls, S, lowest, best, last, &probe);
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
                        um = p.sheet.u_uniform ? p.sheet.u_mask
                             : p.sheet.wide  ? ((const unsigned long long *)p.sheet.umask)[sample]
                                             : (unsigned long long)p.sheet.umask[sample];
                        header_pieces(win, L0, L1, L0 + cut0, L0 + cut1, alen, blen);  // drain (:145) + trim_end (:206)
                        const uint32_t ul = sh_ulen[sample];
                        taglen = ul ? 5 + ul : 0;  // " UMI:" + umi (:207)
                    }
                    __syncwarp();
                } else {
                    // fasta_demultiplex.rs:215-229: the header of mate 2 without its " BC:" field
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
                    cut0 = c0h - L0;
                    cut1 = c1h - L0;
                }

                // ---- the guess is verified against the line count before the first round writes
                if (spec) {
                    g0 = wlb_consume(p.tile_lines, agg16, c, nls_own, lane);
                    spec = false;
                    const uint32_t jt = (4u - (uint32_t)(g0 & 3u)) & 3u;
                    if (jt != j0) {
                        j0 = jt;
                        wrong = true;
                        break;
                    }
                }
                rec0 = (g0 + j0) >> 2;
                // the per-record tables (assign, umi, groups; the barcode stream's record table) hold sk_limits.max_records
                // entries: a batch with more records is refused before anything is written or read past them
                if ((IS_DEMUX && rec0 + nrec > p.max_records) ||
                    (OP == OP_ADDBC && bc_n && (rec0 + nrec - 1u < bc_n ? rec0 + nrec - 1u : bc_n - 1ull) >= p.max_records)) {
                    if (lane == 0) report_err(st, p.max_records, K_TOO_MANY);
                    break;
                }
                const uint64_t rec = rec0 + r;
                if (!IS_DEMUX) {
                    // ---- trim / mask: one output stream in input order.  The tile's output bytes go through a
                    // second look-back (p.tile_out) to its place in the stream; a record is patched in place
                    // into two runs of the window -- header + bases + "\n+\n", qualities + "\n" -- and copied.
                    // (no early ticket for ordered output: a tile's output bytes are published only after its plan
                    // and every later tile waits for them before it writes, so tickets must be taken when tiles start)
                    if (p.unordered && LS && r0 + 32u >= nrec && tid % (GW * 32) == 0) {
                        cta_next = atomicAdd(&st->ticket, (uint32_t)GW);
                        cta_have = true;
                    }
                    if (OP == OP_ADDBC && p.inplace) {
                        // Add barcode, first form: every barcode is taken to be as long as the first and no header to end
                        // in white space, so record r of the stream moves up by r * addD bytes and nothing has to wait
                        // for the barcode table: the rest of the record goes out while the table entry is on its way,
                        // then " BC:" + barcode + "\n" is written behind the header, over the head of the sequence line
                        // (already copied), and header + tag go out as one run.  A record that does not fit the assumption
                        // raises F_NEED_ORDERED: sk_wait repeats the operator in the ordered form below.
                        uint32_t e = L1;
                        while (has && e > L0 && is_ws(win[e - 1])) e--;  // header.trim_end()  (:33)
                        const uint32_t alen = e - L0;
                        RecRef rr;
                        rr.seq_off = 0, rr.seq_len = (uint16_t)(addD - 4u), rr.flags = 0;
                        uint4 bq0 = make_uint4(0u, 0u, 0u, 0u), bq1 = bq0;
                        const bool binl = p.bc_inline != nullptr && addD <= 36u;  // the barcode's bytes come with the table entry
#if !SKW_ADDBC_NOLOAD
                        if (has && bc_n) {
                            const unsigned long long bi = rec < bc_n ? rec : bc_n - 1ull;
                            rr = p.ext_tab[0][bi];
                            if (binl) {
                                bq0 = p.bc_inline[2ull * bi];
                                if (addD > 20u) bq1 = p.bc_inline[2ull * bi + 1ull];
                            }
                        }
#endif
                        if (r0 == 0) {
                            s_l0 = __shfl_sync(FULL, L0, 0);
                            s_obase = c0 + s_l0 + rec0 * addD;
                            const uint32_t tile_outb = (LB(j0 + nrec * 4u) - s_l0) + nrec * addD;
                            out_done = true;
                            s_writable = p.out != nullptr;
                            if (s_writable && s_obase + tile_outb > p.out_cap) {
                                if (lane == 0) report_err(st, rec0, K_OUT_OVERFLOW);
                                s_writable = false;
                            }
                        }
                        uint8_t *gd = p.out + s_obase + (L0 - s_l0) + r * addD;
                        const uint32_t hl = alen + 1u + addD;
                        const bool fits = hl <= L4 - L0;
                        if (s_writable) gcopy(gd + hl, win, L1, has ? L4 - L1 : 0u);
                        __syncwarp();
                        const bool uni = !has || (win[L0] == '@' && !(rr.flags & RR_LONG) && alen + 1u == L1 - L0 && (uint32_t)rr.seq_len + 4u == addD);
                        if (!__all_sync(FULL, uni)) {
                            if (lane == 0) atomicOr(&st->flags, F_NEED_ORDERED);
                        } else if (s_writable) {
#if SKW_ADDBC_NOLOAD
                            const uint8_t *bsrc = win + L3;
#else
                            const uint8_t *bsrc = p.ext_data[0] + rr.seq_off;
#endif
                            const uint32_t bl = addD - 4u;
                            if (has && fits) {
                                uint8_t *d = win + L0 + alen;
                                d[0] = ' '; d[1] = 'B'; d[2] = 'C'; d[3] = ':';
                                if (binl) {
                                    const uint32_t bw[8] = {bq0.x, bq0.y, bq0.z, bq0.w, bq1.x, bq1.y, bq1.z, bq1.w};
#pragma unroll
                                    for (uint32_t t = 0; t < 32u; t++)
                                        if (t < bl) d[4u + t] = (uint8_t)(bw[t >> 2] >> (8u * (t & 3u)));
                                } else {
#pragma unroll 4
                                    for (uint32_t t = 0; t < bl; t++) d[4u + t] = bsrc[t];
                                }
                                d[4u + bl] = '\n';
                            }
                            __syncwarp();
                            gcopy(gd, win, L0, has && fits ? hl : 0u);
                            if (has && !fits) {  // rare: a record shorter than its new header line
                                uint8_t *d = gd;
#pragma unroll 1
                                for (uint32_t i = 0; i < alen; i++) *d++ = win[L0 + i];
                                d[0] = ' '; d[1] = 'B'; d[2] = 'C'; d[3] = ':';
                                d += 4;
#pragma unroll 1
                                for (uint32_t t = 0; t < bl; t++) *d++ = bsrc[t];
                                *d = '\n';
                            }
                        }
                        __syncwarp();
                        continue;
                    }
                    uint32_t a_alen = 0, a_bl = 0, a_bo = 0;
                    if (OP == OP_ADDBC) slen = addbc_plan(has, rec, L0, L1, L4, a_alen, a_bl, a_bo, errk);
                    if (has && errk) report_err(st, rec, errk);
                    uint32_t oincl = slen;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y = __shfl_up_sync(FULL, oincl, o);
                        if (lane >= o) oincl += y;
                    }
                    const uint32_t round_outb = __shfl_sync(FULL, oincl, 31);
                    const uint32_t my_off = s_done + oincl - slen;
                    uint32_t tile_outb = round_outb;
                    if (r0 == 0) {
                        // the look-back wants the whole tile's output: a tile of several rounds (short or uneven
                        // records) plans its later rounds twice, once here for their lengths only
#pragma unroll 1
                        for (uint32_t rr = 32u; rr < nrec; rr += 32u) {
                            const bool has2 = rr + (uint32_t)lane < nrec;
                            const uint32_t j2 = j0 + (has2 ? rr + (uint32_t)lane : 0u) * 4u;
                            uint8_t m2;
                            uint32_t k2 = 0, e2 = 0, x0, x1, x2;
                            const uint32_t l2 = OP == OP_ADDBC
                                                    ? addbc_plan(has2, rec0 + rr + (uint32_t)lane, LB(j2), LB(j2 + 1), LB(j2 + 4), x0, x1, x2, e2)
                                                    : stream_plan(has2, LB(j2), LB(j2 + 1), LB(j2 + 2), LB(j2 + 3), LB(j2 + 4), m2, k2, e2);
                            tile_outb += __reduce_add_sync(FULL, l2);
                        }
                        if (lane == 0 && !p.unordered && !p.inplace) wlb_publish(oagg16, c, tile_outb);
                    }
                    // patches (and the mask itself, in place) while the predecessors' counts arrive
                    uint32_t run1 = 0, run2 = 0;
                    bool slow = false;
                    if (OP != OP_ADDBC && slen && p.out) {
                        if (mode == B_MASK) {
                            if (L1 + kk + 3u <= L3) {
                                mask_copy(win + L1, win + L1, win + L3, kk, p.min_baseq);  // :40-43, in place
                                uint8_t *d = win + L1 + kk;
                                d[0] = '\n'; d[1] = '+'; d[2] = '\n';
                                win[L3 + kk] = '\n';
                                run1 = (L1 - L0) + kk + 3u;
                                run2 = kk + 1u;
                            } else {
                                slow = true;
                            }
                        } else if (mode == B_TRIM) {
                            if (L1 + kk + 3u <= L3) {
                                uint8_t *d = win + L1 + kk;
                                d[0] = '\n'; d[1] = '+'; d[2] = '\n';
                                win[L3 + kk] = '\n';
                                run1 = (L1 - L0) + kk + 3u;
                                run2 = kk + 1u;
                            } else {
                                slow = true;
                            }
                        } else {  // B_GARBAGE (:44-45)
                            if (L1 + 6u <= L4) {
                                uint8_t *d = win + L1;
                                d[0] = 'N'; d[1] = '\n'; d[2] = '+'; d[3] = '\n'; d[4] = '!'; d[5] = '\n';
                                run1 = (L1 - L0) + 6u;
                            } else {
                                slow = true;
                            }
                        }
                    }
                    __syncwarp();
                    if (r0 == 0 && p.unordered) {
                        // space from the cursor; (base, length) noted for the scan and gather passes
                        unsigned long long b = 0;
                        if (lane == 0 && tile_outb) b = atomicAdd(&st->out_cursor, (unsigned long long)((tile_outb + 31u) & ~31u));
                        s_obase = __shfl_sync(FULL, b, 0);
                        out_done = true;
                        s_writable = p.out != nullptr && tile_outb > 0;
                        if (s_writable && s_obase + ((tile_outb + 31u) & ~31u) > p.out_cap) {
                            if (lane == 0) report_err(st, rec0, K_OUT_OVERFLOW);
                            s_writable = false;
                        }
                        if (lane == 0) {
                            p.tile_out[c] = s_obase;
                            ((uint32_t *)(p.tile_out + p.n_chunks))[c] = s_writable ? tile_outb : 0u;
                        }
                    } else if (r0 == 0 && OP == OP_MASK && p.inplace) {
                        // mask of a regular file: the tile's records go where they came from (verified per round below)
                        s_obase = c0 + __shfl_sync(FULL, L0, 0);
                        out_done = true;
                        s_writable = p.out != nullptr && tile_outb > 0;
                        if (s_writable && s_obase + tile_outb > p.out_cap) {
                            if (lane == 0) report_err(st, rec0, K_OUT_OVERFLOW);
                            s_writable = false;
                        }
                    } else if (r0 == 0) {
                        s_obase = wlb_consume(p.tile_out, oagg16, c, tile_outb, lane);
                        out_done = true;
                        if (lane == 0 && c == p.n_chunks - 1) {
                            st->out_bytes = s_obase + tile_outb;
                            st->out_extent = s_obase + tile_outb;
                        }
                        s_writable = p.out != nullptr && tile_outb > 0;
                        if (s_writable && s_obase + tile_outb > p.out_cap) {
                            if (lane == 0) report_err(st, rec0, K_OUT_OVERFLOW);
                            s_writable = false;
                        }
                    }
                    if (LS && g_lead && cta_have) cta_ticket[flipk] = cta_next;  // the group's next ticket, for early loads
                    s_done += round_outb;
                    // Mask by quality of a regular file (bare "+" lines, final newline) changes bytes, not lengths:
                    // every record's two runs are the record's own bytes in the window and the round's output is
                    // one piece of the window, first header to last newline.  The warp then copies that piece as
                    // a whole, 16 destination-aligned bytes per lane and step (512 contiguous bytes per store
                    // instruction), instead of two runs per lane.
                    bool whole = false;
                    if (OP == OP_MASK && (s_writable || p.inplace)) {
                        const bool idl = !slen || (!slow && L0 + run1 == L3 && L3 + run2 == L4);
                        const uint32_t src0 = __shfl_sync(FULL, L0, 0);
                        const uint32_t srce = __reduce_max_sync(FULL, slen ? L4 : 0u);
                        whole = __all_sync(FULL, idl) && srce == src0 + round_outb;
                        if (p.inplace && (!whole || !s_writable || __any_sync(FULL, has && !slen))) {
                            // a record that changes its length, fails or cannot be written: the input offsets are
                            // not the output offsets; nothing is written and sk_wait runs the ordered form
                            if (lane == 0) atomicOr(&st->flags, F_NEED_ORDERED);
                            whole = true;
                        } else if (whole) {
                            uint8_t *g0 = p.out + s_obase + (s_done - round_outb);
                            uint32_t so = src0, len = round_outb;
                            const uint32_t head = min((16u - ((uint32_t)(uintptr_t)g0 & 15u)) & 15u, len);
                            if ((uint32_t)lane < head) g0[lane] = win[so + (uint32_t)lane];
                            g0 += head, so += head, len -= head;
                            const uint32_t nv = len >> 4;
#pragma unroll 2
                            for (uint32_t u = (uint32_t)lane; u < nv; u += 32u) {
                                const uint4 v = lds_unaligned16(win, so + 16u * u);
                                *(uint4 *)(g0 + 16u * u) = v;
                            }
                            const uint32_t tl = len & 15u;
                            if ((uint32_t)lane < tl) g0[16u * nv + (uint32_t)lane] = win[so + 16u * nv + (uint32_t)lane];
                        }
                    }
                    if (OP == OP_ADDBC) {
                        if (s_writable) {  // ordered form
                            // the rest of the record goes out first; then " BC:" + barcode + "\n" is written behind the
                            // trimmed header, over the head of the sequence line, and header + tag go out as one run
                            uint8_t *gd = p.out + s_obase + my_off;
                            const uint32_t hl = a_alen + 5u + a_bl;
                            const bool fits = slen && hl <= L4 - L0;
                            gcopy(gd + hl, win, L1, slen ? L4 - L1 : 0u);
                            __syncwarp();
                            const uint8_t *bsrc = p.ext_data[0] + a_bo;
                            if (fits) {
                                uint8_t *d = win + L0 + a_alen;
                                d[0] = ' '; d[1] = 'B'; d[2] = 'C'; d[3] = ':';
#pragma unroll 4
                                for (uint32_t t = 0; t < a_bl; t++) d[4u + t] = bsrc[t];
                                d[4u + a_bl] = '\n';
                            }
                            __syncwarp();
                            gcopy(gd, win, L0, fits ? hl : 0u);
                            if (slen && !fits) {  // rare: a record shorter than its new header line
                                uint8_t *d = gd;
#pragma unroll 1
                                for (uint32_t i = 0; i < a_alen; i++) *d++ = win[L0 + i];
                                d[0] = ' '; d[1] = 'B'; d[2] = 'C'; d[3] = ':';
                                d += 4;
#pragma unroll 1
                                for (uint32_t t = 0; t < a_bl; t++) *d++ = bsrc[t];
                                *d = '\n';
                            }
                        }
                        __syncwarp();
                        continue;
                    }
                    if (s_writable && !whole) {
                        uint8_t *gd = p.out + s_obase + my_off;
#pragma unroll 1
                        for (int q = 0; q < 2; q++) {
                            gcopy(q == 0 ? gd : gd + run1, win, q == 0 ? L0 : L3, q == 0 ? run1 : run2);
                            __syncwarp();
                        }
                        if (slow) {  // rare: a '+' line too short to hold the patch, a record cut off by the end of the stream
                            uint8_t *d = gd;
#pragma unroll 1
                            for (uint32_t i = L0; i < L1; i++) *d++ = win[i];
                            if (mode == B_GARBAGE) {
                                d[0] = 'N'; d[1] = '\n'; d[2] = '+'; d[3] = '\n'; d[4] = '!'; d[5] = '\n';
                            } else {
#pragma unroll 1
                                for (uint32_t i = 0; i < kk; i++) {
                                    const uint8_t sq = win[L1 + i];
                                    *d++ = (mode == B_MASK && (uint8_t)(win[L3 + i] - 33u) < p.min_baseq) ? (uint8_t)'N' : sq;
                                }
                                d[0] = '\n'; d[1] = '+'; d[2] = '\n';
                                d += 3;
#pragma unroll 1
                                for (uint32_t i = 0; i < kk; i++) *d++ = win[L3 + i];
                                *d = '\n';
                            }
                        }
                    }
                    __syncwarp();
                    continue;
                }

                // ---- loads whose results are needed further down go out first: mate 2 reads the pair's
                // sample and UMI from mate 1's tables; the last round takes the warp's next ticket
                uint8_t *gu = p.umi + rec * p.sheet.Umax;
                int a16 = -1;
                uint2 guv = make_uint2(0u, 0u);
                if (!D1 && has && p.out && rec < r1_nrec) {
                    a16 = (int)p.assign[rec];
                    if (p.sheet.Umax == 8u) guv = *(const uint2 *)gu;
                }
                if (r0 + 32u >= nrec && lane == 0) {
                    if (LS) {
                        if (wg == 0) {
                            cta_next = atomicAdd(&st->ticket, (uint32_t)GW);
                            cta_have = true;
                        }
                    } else {
                        c_next = atomicAdd(&st->ticket, 1u);
                        have_next = true;
                    }
                }

                // ---- body: "\n+\n" and "\n" are patched in behind the kept bases / qualities (:47), whether or
                // not the record is written in the end (the window is private to the tile)
                uint32_t run1 = 0, run2 = 0;
                bool bslow = false;
                if (has && p.out) {
                    if (mode == B_VERBATIM) {
                        run1 = body;
                    } else if (mode == B_TRIM) {
                        if (L1 + kk + 3u <= L3) {
                            uint8_t *d = win + L1 + kk;
                            d[0] = '\n'; d[1] = '+'; d[2] = '\n';
                            win[L3 + kk] = '\n';
                            run1 = kk + 3u;
                            run2 = kk + 1u;
                        } else {
                            bslow = true;
                        }
                    } else if (mode == B_GARBAGE) {  // :44-45
                        if (L1 + 6u <= L4) {
                            uint8_t *d = win + L1;
                            d[0] = 'N'; d[1] = '\n'; d[2] = '+'; d[3] = '\n'; d[4] = '!'; d[5] = '\n';
                            run1 = 6u;
                        } else {
                            bslow = true;
                        }
                    }
                }

                // ---- outcome of every record: counters, failures, ambiguity events (:169-194)
                if (has) {
                    if (D1) {
                        if (sample == -1 && taglen != 0xFFu && taglen != 0u) report_err(st, rec, taglen);
                        if (sample != -1 || taglen == 0u) my_total++;  // :169
                        if (sample >= 0) {                              // :177-178
                            my_ident++;
                            if (cc_smem) atomicAdd(&ccount[sample], 1u);
                            else atomicAdd(&p.counts[sample], 1ull);
                        } else if (sample == -2) {  // :184-188
                            const uint32_t ei = atomicAdd(&st->n_events, 1u);
                            if (ei < p.events_cap) {
                                Event ev;
                                ev.record = (uint32_t)rec;
                                ev.bc_off = (uint32_t)(c0 + L0 + cut0 + 4u);
                                ev.bc_off2 = 0xFFFFFFFFu;
                                ev.best = (int16_t)alen;
                                ev.last = (int16_t)blen;
                                ev.mismatches = taglen;
                                p.events[ei] = ev;
                            } else {
                                atomicOr(&st->flags, F_EVENTS_OVERFLOW);
                            }
                        }
                    } else {
                        sample = a16;
                        if (sample >= 0) {
                            const uint32_t ul = sh_ulen[sample];
                            taglen = ul ? 5 + ul : 0;
                            if (fused && !(L1 > L0 && win[L1 - 1] == '\n')) {
                                report_err(st, rec, K_TRUNC_FUSED);
                                sample = -1;
                            }
                        } else {
                            sample = -1;
                        }
                    }
                }
                uint32_t outlen = 0;
                if (has && sample >= 0) {
                    if (mode == B_FAIL) {  // &seq[..k] would panic (fasta_trim_by_quality.rs:47)
                        report_err(st, rec, K_SEQ_SHORT);
                        if (D1) sample = -1;
                        mode = B_NONE;
                    } else if (!p.out) {
                        mode = B_NONE;  // dry run: count only (:77-78,:179)
                    } else {
                        outlen = alen + blen + taglen + 1u + body;
                    }
                }
                if (D1 && has) p.assign[rec] = (int16_t)sample;

                // ---- place of the record in the round's output (input order), space for the round
                uint32_t oincl = outlen;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(FULL, oincl, o);
                    if (lane >= o) oincl += y;
                }
                const uint32_t round_out = __shfl_sync(FULL, oincl, 31);
                const uint32_t my_off = oincl - outlen;
                const uint32_t emit_mask = __ballot_sync(FULL, outlen != 0);
                const uint32_t my_rank = __popc(emit_mask & ((1u << lane) - 1u));
                const uint32_t n_emit = __popc(emit_mask);
                unsigned long long rbase = 0;
                if (lane == 0 && p.out && round_out) {  // the reply is awaited after the header patches
                    rbase = atomicAdd(&st->out_cursor, (unsigned long long)((round_out + 31u) & ~31u));  // whole sectors
                    my_out += round_out;
                }
                const bool emit = outlen != 0;
                const uint32_t ul = taglen ? taglen - 5u : 0u;

                // ---- header: " UMI:x\n" goes over the deleted " BC:x" when the header ends with it
                const bool hpatch = emit && blen == 0 && alen + taglen + 1u <= L1 - L0;
                uint32_t hrun = 0;
                uint32_t ulo = 0, uhi = 0;  // the UMI in registers (up to eight characters)
                bool ureg = false;
                if (D1) {
                    if (has && sample >= 0 && p.sheet.Umax) {
                        // UMI = observed chars where the sheet barcode has 'U' (:200-203), parked in the side
                        // table for mate 2 (also on a dry run); read before the patch overwrites the barcode
                        const uint8_t *ob = win + L0 + cut0 + 4;
                        unsigned long long m = um;
                        const uint32_t ulen = sh_ulen[sample];
                        const uint32_t u0 = (uint32_t)__ffsll((long long)m) - 1u;
                        if (ulen && ulen <= 8u && (m >> u0) == ((1ull << ulen) - 1ull)) {  // one run of U (the usual sheet)
                            const uint2 uv = lds_un64(win, L0 + cut0 + 4u + u0);
                            ulo = uv.x;
                            uhi = uv.y;
                            ureg = true;
                            if (p.sheet.Umax == 8u && ulen == 8u) {
                                *(uint2 *)gu = make_uint2(ulo, uhi);
                            } else {
#pragma unroll 1
                                for (uint32_t t = 0; t < ulen; t++)
                                    gu[t] = (uint8_t)(t < 4u ? ulo >> (8u * t) : uhi >> (8u * (t - 4u)));
                            }
                        } else {
                            uint32_t t = 0;
#pragma unroll 1
                            while (m) {
                                const uint32_t q = (uint32_t)__ffsll((long long)m) - 1u;
                                m &= m - 1;
                                gu[t++] = ob[q];
                            }
                        }
                    }
                } else if (p.sheet.Umax == 8u) {
                    ulo = guv.x;
                    uhi = guv.y;
                    ureg = true;
                }
                __syncwarp();
#define UMI_BYTE(t) (ureg ? (uint8_t)((t) < 4u ? ulo >> (8u * (t)) : uhi >> (8u * ((t) - 4u))) : gu[(t)])
                if (hpatch) {
                    uint8_t *d = win + L0 + alen;
                    if (taglen) {
#pragma unroll 1
                        for (uint32_t t = 0; t < ul; t++) d[5 + t] = UMI_BYTE(t);
                        d[0] = ' '; d[1] = 'U'; d[2] = 'M'; d[3] = 'I'; d[4] = ':';
                    }
                    d[taglen] = '\n';
                    hrun = alen + taglen + 1u;
                }
                __syncwarp();

                // ---- emit
                rbase = __shfl_sync(FULL, rbase, 0);
                if (LS && g_lead && cta_have) cta_ticket[flipk] = cta_next;  // the group's next ticket, for early loads
                bool writable = p.out != nullptr && round_out > 0;
                if (writable && rbase + ((round_out + 31u) & ~31u) > p.out_cap) {
                    if (lane == 0) report_err(st, rec0 + r0, K_OUT_OVERFLOW);
                    writable = false;
                }
                // Sector emit (the usual round: every record patched in place, one round in the tile).  The
                // round's output is the concatenation of three runs per record (run table behind the line
                // table).  A lane writes the 32-byte sectors that start inside its record, each with one
                // STG.256: a sector inside one run is one unaligned 32-byte read of the window; at a seam the
                // lane walks on through the table -- into the next records' runs if need be -- reading each
                // piece at the offset that puts its bytes in place.  No partial-sector store, a third of the
                // store requests of the run copies below.
                const bool sect = writable && nrec <= 32u && !__any_sync(FULL, emit && (!hpatch || bslow));
                if (sect) {
                    uint32_t *rt = (uint32_t *)(ls + 144);
                    const uint32_t ro1 = my_off + (emit ? hrun : 0u), ro2 = ro1 + (emit ? run1 : 0u);
                    rt[3 * lane] = my_off | (L0 << 16);
                    rt[3 * lane + 1] = ro1 | (L1 << 16);
                    rt[3 * lane + 2] = ro2 | (L3 << 16);
                    if (lane == 0) rt[96] = round_out;
                    __syncwarp();
                    uint8_t *gbase = p.out + rbase;
#if SKW_EMIT_RUNS
                    // The lane goes through the three runs of its record in turn.  Sectors that lie inside one
                    // run are one unaligned 32-byte read of the window and one store (no merge, no table); the
                    // sector that straddles the end of a run -- at most one per run -- is put together from the
                    // tail of this run and the heads of the runs that follow (the next records' runs if need be).
                    {
                        uint32_t i = 3u * (uint32_t)lane;
                        uint32_t e = rt[i], en = rt[i + 1];
                        uint32_t o = (my_off + 31u) & ~31u;
#pragma unroll 1
                        for (int q = 0; q < 3; q++) {
                            const uint32_t rb = emit ? (en & 0xFFFFu) : 0u;  // end of this run in the round's output
                            const int delta = (int)(e >> 16) - (int)(e & 0xFFFFu);
                            bool in = o + 32u <= rb;
                            {   // consecutive sectors of a run share a word of the window: eight loads per sector, not nine
                                const int so = (int)o + delta;
                                const uint32_t *pw = (const uint32_t *)(win + (so & ~3));
                                const uint32_t sh = ((uint32_t)so & 3u) * 8u;
                                uint32_t lo = in ? pw[0] : 0u;
#pragma unroll 1
                                while (__any_sync(FULL, in)) {
                                    if (in) {
                                        const uint32_t w1 = pw[1], w2 = pw[2], w3 = pw[3], w4 = pw[4], w5 = pw[5], w6 = pw[6], w7 = pw[7], w8 = pw[8];
                                        stg256(gbase + o, __funnelshift_r(lo, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                                               __funnelshift_r(w3, w4, sh), __funnelshift_r(w4, w5, sh), __funnelshift_r(w5, w6, sh),
                                               __funnelshift_r(w6, w7, sh), __funnelshift_r(w7, w8, sh));
                                        lo = w8;
                                        pw += 8;
                                        o += 32u;
                                        in = o + 32u <= rb;
                                    }
                                }
                            }
                            bool seam = o < rb;
                            if (__any_sync(FULL, seam)) {
                                uint32_t si = i, se = e, sen = en, filled = 0;
                                U256 v;
#pragma unroll
                                for (int k = 0; k < 8; k++) v.w[k] = 0u;
#pragma unroll 1
                                while (__any_sync(FULL, seam)) {
                                    if (seam) {
#pragma unroll 1
                                        while ((sen & 0xFFFFu) <= o + filled && si < 95u) {  // runs that end before the next byte
                                            si++;
                                            se = sen;
                                            sen = rt[si + 1];
                                        }
                                        const U256 ld = lds_unaligned32(win, (int)(se >> 16) + (int)o - (int)(se & 0xFFFFu));
                                        v = merge_low32(v, ld, filled);
                                        const uint32_t upto = (sen & 0xFFFFu) - o;  // bytes of the sector known after this run
                                        if (upto >= 32u || (sen & 0xFFFFu) >= round_out) {
                                            stg256(gbase + o, v.w[0], v.w[1], v.w[2], v.w[3], v.w[4], v.w[5], v.w[6], v.w[7]);
                                            o += 32u;
                                            seam = false;
                                        } else {
                                            filled = upto;
                                        }
                                    }
                                }
                            }
                            i++;
                            e = en;
                            en = rt[i + 1];
                        }
                    }
#else
                    // One step = one piece: read 32 bytes of the current run at the offset that puts them in
                    // place, merge them behind the bytes the sector already has, then either store the sector
                    // (full, or the round's last) or move on to the next run.  Every lane takes the same path
                    // through a step, whatever its seams.
                    {
                        uint32_t i = 3u * (uint32_t)lane;
                        uint32_t e = rt[i], en = rt[i + 1];
                        const uint32_t oend = my_off + outlen;
                        uint32_t o = (my_off + 31u) & ~31u, filled = 0;
                        bool act = emit && o < oend;
                        U256 v;
#pragma unroll
                        for (int k = 0; k < 8; k++) v.w[k] = 0u;
#pragma unroll 1
                        while (__any_sync(FULL, act)) {
                            if (act) {
#pragma unroll 1
                                while ((en & 0xFFFFu) <= o + filled && i < 95u) {  // runs that end before the next byte
                                    i++;
                                    e = en;
                                    en = rt[i + 1];
                                }
                                const U256 ld = lds_unaligned32(win, (int)(e >> 16) + (int)o - (int)(e & 0xFFFFu));
                                v = merge_low32(v, ld, filled);
                                const uint32_t upto = (en & 0xFFFFu) - o;  // bytes of the sector known after this run
                                if (upto >= 32u || (en & 0xFFFFu) >= round_out) {
                                    stg256(gbase + o, v.w[0], v.w[1], v.w[2], v.w[3], v.w[4], v.w[5], v.w[6], v.w[7]);
                                    o += 32u;
                                    filled = 0;
                                    act = o < oend;
                                } else {
                                    filled = upto;
                                }
                            }
                        }
                    }
#endif
                    if (emit) {
                        Group g;
                        g.sample = (uint16_t)sample;
                        g.len = (uint16_t)outlen;
                        p.groups[rec0 + r0 + my_rank] = g;
                    }
                    __syncwarp();
                } else if (writable) {
                    uint8_t *gd = p.out + rbase + my_off;
                    if (emit && !hpatch) {  // rare: a header piece after the cut, or no room for the tag
                        uint8_t *d = gd;
#pragma unroll 1
                        for (uint32_t i = 0; i < alen; i++) *d++ = win[L0 + i];
#pragma unroll 1
                        for (uint32_t i = 0; i < blen; i++) *d++ = win[L0 + cut1 + i];
                        if (taglen) {
                            d[0] = ' '; d[1] = 'U'; d[2] = 'M'; d[3] = 'I'; d[4] = ':';
#pragma unroll 1
                            for (uint32_t t = 0; t < ul; t++) d[5 + t] = UMI_BYTE(t);
                            d += taglen;
                        }
                        *d = '\n';
                    }
                    __syncwarp();
                    const uint32_t hlen = alen + blen + taglen + 1u;
                    // the record's runs: header, bases (+ "\n+\n"), qualities (+ "\n")
#pragma unroll 1
                    for (int q = 0; q < 3; q++) {
                        const uint32_t so = q == 0 ? L0 : q == 1 ? L1 : L3;
                        const uint32_t ln = !emit ? 0u : q == 0 ? hrun : q == 1 ? run1 : run2;
                        uint8_t *dd = q == 0 ? gd : q == 1 ? gd + hlen : gd + hlen + run1;
                        gcopy(dd, win, so, ln);
                        __syncwarp();
                    }
                    if (emit && bslow) {  // rare: a '+' line too short to hold the patch
                        uint8_t *d = gd + hlen;
                        if (mode == B_TRIM) {
#pragma unroll 1
                            for (uint32_t i = 0; i < kk; i++) *d++ = win[L1 + i];
                            d[0] = '\n'; d[1] = '+'; d[2] = '\n';
                            d += 3;
#pragma unroll 1
                            for (uint32_t i = 0; i < kk; i++) *d++ = win[L3 + i];
                            *d = '\n';
                        } else {
                            d[0] = 'N'; d[1] = '\n'; d[2] = '+'; d[3] = '\n'; d[4] = '!'; d[5] = '\n';
                        }
                    }
                    if (emit) {
                        Group g;
                        g.sample = (uint16_t)sample;
                        g.len = (uint16_t)outlen;
                        p.groups[rec0 + r0 + my_rank] = g;
                    }
                }
                if (lane == 0 && p.out) {
                    ChunkRow row;
                    row.base = rbase;
                    row.first_group = (uint32_t)(rec0 + r0);
                    row.n_groups = writable ? n_emit : 0u;
                    p.rows[(size_t)c * GeoW::ROUNDS + (r0 >> 5)] = row;
                }
                __syncwarp();
            }
            if (wrong) continue;
            if (spec) {  // no round ran (the tile was given up): the prefix is still owed to the successors
                g0 = wlb_consume(p.tile_lines, agg16, c, nls_own, lane);
                spec = false;
                const uint32_t jt = (4u - (uint32_t)(g0 & 3u)) & 3u;
                if (jt != j0) {  // also re-evaluates a tile given up under the wrong framing
                    j0 = jt;
                    continue;
                }
            }
            break;
        }
        if (bail && lane == 0) atomicOr(&st->flags, F_NEED_GENERAL);
        if (!IS_DEMUX && !out_done && p.unordered) {
            if (lane == 0) {
                p.tile_out[c] = 0;
                ((uint32_t *)(p.tile_out + p.n_chunks))[c] = 0u;
            }
        } else if (!IS_DEMUX && !out_done && !p.inplace) {  // a tile without a round still owes its (empty) output to the look-back
            if (lane == 0) wlb_publish(oagg16, c, 0u);
            const uint64_t obase = wlb_consume(p.tile_out, oagg16, c, 0u, lane);
            if (lane == 0 && c == p.n_chunks - 1) {
                st->out_bytes = obase;
                st->out_extent = obase;
            }
        }
        if (IS_DEMUX && p.out) {  // rows of the rounds that did not run
            const uint32_t ran = (nrec + 31u) >> 5;
            if ((uint32_t)lane < (uint32_t)GeoW::ROUNDS && (uint32_t)lane >= ran) {
                ChunkRow row;
                row.base = 0;
                row.first_group = 0;
                row.n_groups = 0;
                p.rows[(size_t)c * GeoW::ROUNDS + lane] = row;
            }
        }
        if (lane == 0) {
            if (c == p.n_chunks - 1) st->n_lines = g0 + nls_own;
            if (nrec) {  // flushed once, after the last tile
                my_nrec += nrec;
                const unsigned long long upto = c0 + LB(j0 + nrec * 4u);
                my_upto = upto > my_upto ? upto : my_upto;
                if (OP == OP_ADDBC) {
                    const unsigned long long oend = upto + (((g0 + j0) >> 2) + nrec) * addD;
                    my_outmax = oend > my_outmax ? oend : my_outmax;
                }
            }
        }
        // The next ticket is taken during the last round (or now): a tile's count must appear soon after
        // its ticket, its successors wait for it.
        if (!LS) {
            if (lane == 0 && !have_next) c_next = atomicAdd(&st->ticket, 1u);
            c = __shfl_sync(FULL, c_next, 0);
        } else {  // the CTA's next ticket may be known already (tickets only grow): start the next load now
            // (An unsynchronised peek at a word only the group's leader writes -- compute-sanitizer's racecheck reports
            // it against the two stores above.  Either the slot still holds the ticket of two tiles ago, which is not
            // newer than this tile's and is ignored, or the leader's next ticket, which the barrier at the top of the
            // loop hands to every warp anyway.)
            const uint32_t v = cta_ticket[flipk];
            if (v > cb && v + (uint32_t)wg < p.n_chunks) {
                issue_load(v + (uint32_t)wg);
                early = true;
            }
        }
    }
#undef LB
#undef UMI_BYTE
    if (lane == 0 && my_out) atomicAdd(&st->out_bytes, my_out);
    if (lane == 0 && my_nrec) {
        atomicAdd(&st->n_records, my_nrec);
        atomicMax(&st->consumed, my_upto);
        if (OP == OP_MASK && p.inplace) {  // output == input bytes of the complete records
            atomicMax(&st->out_bytes, my_upto);
            atomicMax(&st->out_extent, my_upto);
        }
        if (OP == OP_ADDBC && p.inplace) {
            atomicMax(&st->out_bytes, my_outmax);
            atomicMax(&st->out_extent, my_outmax);
        }
    }
    if (D1) {  // fasta_demultiplex.rs:108-109,169,177-178
        const uint32_t wt = __reduce_add_sync(FULL, my_total), wi = __reduce_add_sync(FULL, my_ident);
        if (lane == 0 && wt) atomicAdd(&p.counts[S], (unsigned long long)wt);
        if (lane == 0 && wi) atomicAdd(&p.counts[S + 1], (unsigned long long)wi);
        if (cc_smem) {
            __syncthreads();
#pragma unroll 1
            for (uint32_t s = tid; s < S; s += GeoW::NT)
                if (ccount[s]) atomicAdd(&p.counts[s], (unsigned long long)ccount[s]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// unordered trim: scan of the tiles' output lengths, gather into input order
// ------------------------------------------------------------------------------------------------
// dst[c] = sum of len[0..c) in two small launches: sums of blocks of 1024 tiles, then every block adds the
// sums before it to the scan of its own 1024 lengths; the total goes into the outcome block.
__global__ void __launch_bounds__(1024) sk_tile_sum_kernel(const uint32_t *len, unsigned long long *bsum, uint32_t n) {
    __shared__ unsigned long long ws[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    unsigned long long v = i < n ? len[i] : 0u;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) ws[w] = v;
    __syncthreads();
    if (w == 0) {
        v = ws[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) bsum[blockIdx.x] = v;
    }
}
__global__ void __launch_bounds__(1024) sk_tile_scan_kernel(const uint32_t *len, const unsigned long long *bsum, uint64_t *dst,
                                                             uint32_t n, DevStats *st) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long s_before;
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    // sums of the blocks before this one (a few hundred values at most)
    unsigned long long b = 0;
    for (uint32_t j = threadIdx.x; j < blockIdx.x; j += 1024u) b += bsum[j];
#pragma unroll
    for (int o = 16; o; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane == 0) ws[w] = b;
    __syncthreads();
    if (w == 0) {
        b = ws[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) s_before = b;
    }
    __syncthreads();
    const unsigned long long before = s_before;
    // inclusive scan of this block's lengths
    const unsigned long long own = i < n ? len[i] : 0u;
    unsigned long long x = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
        if ((int)lane >= o) x += y;
    }
    __syncthreads();  // ws is reused
    if (lane == 31) ws[w] = x;
    __syncthreads();
    unsigned long long wb = 0;
    for (uint32_t k = 0; k < w; k++) wb += ws[k];
    const unsigned long long incl = before + wb + x;
    if (i < n) dst[i] = incl - own;
    if (i == n - 1) {
        st->out_bytes = incl;
        st->out_extent = incl;
    }
}

// One warp per tile: len[c] bytes from the 32-byte aligned scratch + base[c] to out + dst[c] (any alignment),
// 16 destination-aligned bytes per lane and step, read as two aligned 16-byte loads and shifted into place.
__global__ void __launch_bounds__(256) sk_tile_gather_kernel(const uint8_t *scratch, uint8_t *out, const uint64_t *base,
                                                             const uint32_t *len, const uint64_t *dst, uint32_t n) {
    const uint32_t lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    for (uint32_t c = blockIdx.x * wpb + (threadIdx.x >> 5); c < n; c += gridDim.x * wpb) {
        const uint32_t L = len[c];
        if (!L) continue;
        const uint8_t *src = scratch + base[c];
        uint8_t *d = out + dst[c];
        const uint32_t a = (uint32_t)(uintptr_t)d & 15u;  // d - a is 16-byte aligned
        const uint32_t r = (16u - a) & 15u, wo = r >> 2, bs = (r & 3u) * 8u;
        const uint32_t units = (a + L + 15u) >> 4;
        for (uint32_t u = lane; u < units; u += 32u) {
            const int s0 = (int)(16u * u) - (int)a;  // source offset of the unit's first byte
            const int k = s0 >> 4;                   // aligned 16-byte chunk that holds it
            uint4 A = make_uint4(0u, 0u, 0u, 0u), B = make_uint4(0u, 0u, 0u, 0u);
            if (k >= 0) A = *(const uint4 *)(src + 16 * k);
            if (r && 16u * (uint32_t)(k + 1) < ((L + 31u) & ~31u)) B = *(const uint4 *)(src + 16 * (k + 1));  // inside the tile's sectors
            const uint32_t W[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
            uint4 o;
            switch (wo) {  // uniform over the tile
                case 0: o = make_uint4(__funnelshift_r(W[0], W[1], bs), __funnelshift_r(W[1], W[2], bs), __funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs)); break;
                case 1: o = make_uint4(__funnelshift_r(W[1], W[2], bs), __funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs)); break;
                case 2: o = make_uint4(__funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs), __funnelshift_r(W[5], W[6], bs)); break;
                default: o = make_uint4(__funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs), __funnelshift_r(W[5], W[6], bs), __funnelshift_r(W[6], W[7], bs)); break;
            }
            uint8_t *du = d - a + 16u * u;
            const uint32_t first = u == 0 ? a : 0u;                                     // valid bytes of the unit: [first, last)
            const uint32_t last = 16u * u + 16u > a + L ? a + L - 16u * u : 16u;
            if (first == 0u && last == 16u) {
                *(uint4 *)du = o;
            } else {  // the tile's first and last unit are shared with its neighbours
                const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
                for (uint32_t i = first; i < last; i++) du[i] = (uint8_t)(ow[i >> 2] >> (8u * (i & 3u)));
            }
        }
    }
}

// (base, length, destination) of every tile in p.tile_out: u64 base[n], u32 len[n], u64 dst[n], then the block sums
static inline uint64_t *tt_base(const KParams &p) { return p.tile_out; }
static inline uint32_t *tt_len(const KParams &p) { return (uint32_t *)(p.tile_out + p.n_chunks); }
static inline uint64_t *tt_dst(const KParams &p) { return p.tile_out + p.n_chunks + (p.n_chunks + 1u) / 2u; }

int launch_tile_gather(const KParams &p, int sm_count, void *stream_, const char **err) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p.n_chunks) return 0;
    const unsigned nb = (p.n_chunks + 1023u) / 1024u;
    unsigned long long *bsum = (unsigned long long *)(tt_dst(p) + p.n_chunks);  // behind the three per-tile arrays
    sk_tile_sum_kernel<<<nb, 1024, 0, stream>>>(tt_len(p), bsum, p.n_chunks);
    sk_tile_scan_kernel<<<nb, 1024, 0, stream>>>(tt_len(p), bsum, tt_dst(p), p.n_chunks, p.stats);
    const unsigned grid = (unsigned)std::min<long long>((long long)sm_count * 8, ((long long)p.n_chunks + 7) / 8);
    sk_tile_gather_kernel<<<grid ? grid : 1u, 256, 0, stream>>>(p.out, p.final_out, tt_base(p), tt_len(p), tt_dst(p), p.n_chunks);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    return 3;
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
static uint32_t warp_smem(uint32_t S, uint32_t n_classes, uint32_t nwp, bool d1) {
    uint32_t o = WLayout::dyn;
    if (d1) {
        o += ((n_classes * HIDX_CLS_ROWS * nwp + 3u) & ~3u) * 4u;
        if (S <= (uint32_t)FAST_CCOUNT_MAX) o += ((S + 3u) & ~3u) * 4u;
    }
    o += (S + 15u) & ~15u;  // per-sample UMI lengths
    return o;
}

bool warp_supported(int op, const KParams &p) {
    if (p.lpr != 4 || p.tile_lanes < 8 || p.tile_lanes > 30) return false;
    if (op == OP_TRIM || op == OP_MASK) return true;
    if (op == OP_ADDBC) return p.head_char == '@' && p.ext_tab[0] && p.ext_data[0] && p.ext_stats[0];
    if (op != OP_DEMUX1 && op != OP_DEMUX2) return false;
    if (p.n_index || !p.sheet.hidx.n_classes || !p.sheet.fidx.table) return false;
    if (p.tile_lanes < 8 || p.tile_lanes > 30) return false;
    return warp_smem(p.sheet.S, p.sheet.hidx.n_classes, p.sheet.hidx.nwp, true) <= (GeoW::WARPS == 8 ? 115200u : 231000u);
}

template <int OP, int NWMAX>
static int launch_warp_one(const KParams &p, int sm_count, cudaStream_t stream, const char **err) {
    auto kfn = sk_warp_kernel<OP, NWMAX>;
    const bool stream_op = OP == OP_TRIM || OP == OP_MASK || OP == OP_ADDBC;  // no sheet tables
    const int smem = (int)warp_smem(stream_op ? 0u : p.sheet.S, stream_op ? 0u : p.sheet.hidx.n_classes, p.sheet.hidx.nwp, OP == OP_DEMUX1);
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, GeoW::NT, smem);
    if (e != cudaSuccess || per_sm < 1) {
        *err = e != cudaSuccess ? cudaGetErrorString(e) : "kernel does not fit on an SM";
        return -1;
    }
    const long long tiles_per_cta = GeoW::WARPS;
    long long grid = (long long)sm_count * per_sm;
    const long long want = ((long long)p.n_chunks + tiles_per_cta - 1) / tiles_per_cta;
    if (grid > want) grid = want;
    if (grid < 1) return 0;
    kfn<<<(unsigned)grid, GeoW::NT, smem, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    return 1;
}

int launch_warp_kernel(int op, const KParams &p, int sm_count, void *stream_, const char **err) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool wide = p.sheet.wide != 0;
    switch (op) {
        case OP_DEMUX1:
            return wide ? launch_warp_one<OP_DEMUX1, 16>(p, sm_count, stream, err)
                        : launch_warp_one<OP_DEMUX1, 8>(p, sm_count, stream, err);
        case OP_DEMUX2: return launch_warp_one<OP_DEMUX2, 8>(p, sm_count, stream, err);
        case OP_TRIM: return launch_warp_one<OP_TRIM, 8>(p, sm_count, stream, err);
        case OP_MASK: return launch_warp_one<OP_MASK, 8>(p, sm_count, stream, err);
        case OP_ADDBC: return launch_warp_one<OP_ADDBC, 8>(p, sm_count, stream, err);
    }
    *err = "operator not handled by the warp engine";
    return -1;
}

}  // namespace sk
