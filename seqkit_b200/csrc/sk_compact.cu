// sk_compact.cu -- device-side stable per-sample compaction of a demultiplex result.
//
// The reference appends every assigned record to its sample's file as it meets it
// (fasta_demultiplex.rs:196-238), so a sample's file is the sample's records in input order.  The
// demultiplex kernels write the emitted records of a batch in input order, with one slice-table row per
// chunk / round and one (sample, len) group per piece (sk_internal.h: ChunkRow, Group).  The four kernels
// here turn that into one contiguous run of bytes per sample -- a stable partition by sample -- so that
// the host appends S slices per batch and mate instead of one piece per record:
//
//   hist    per block of CB_ROWS consecutive rows: bytes of every sample (shared-memory histogram)
//   cols    per sample: exclusive prefix of its block totals (where the block's pieces start inside
//           the sample's run) and the sample's total
//   bases   one CTA: the samples' runs back to back, each starting on a 128-byte line -> slices[]
//   addr    per block, one warp walking its rows in order: destination of every piece = base of the
//           sample's run + bytes of the sample in earlier blocks + bytes of the sample earlier in this
//           block (match.any groups the lanes of a round by sample; ties keep lane = input order)
//   move    one warp per row: every piece goes from its place in the input-order stream to its
//           destination, 16 destination-aligned bytes per lane and step
//
// Stable by construction: blocks, rows inside a block, groups inside a row and lanes inside a match
// group are all taken in input order.  HBM traffic: the emitted bytes once more in and out, plus 12 bytes
// per piece of tables; the kernels are bandwidth work, nothing here is a contraction.
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

constexpr uint32_t CB_ROWS = 512;  // slice-table rows per compaction block (128 tiles of the warp engine)

// ---- hist ----------------------------------------------------------------------------------------
// One warp takes 32 rows at a time (one 16-byte row per lane, coalesced), then walks the non-empty ones.
template <class F>
__device__ __forceinline__ void for_rows_of_block(const ChunkRow *rows, uint32_t n_rows, uint32_t blk, int warp, int nwarps,
                                                  int lane, F f) {
    const uint32_t r0 = blk * CB_ROWS, r1 = min(r0 + CB_ROWS, n_rows);
    for (uint32_t rb = r0 + 32u * (uint32_t)warp; rb < r1; rb += 32u * (uint32_t)nwarps) {
        const uint32_t r = rb + (uint32_t)lane;
        ChunkRow row;
        row.base = 0, row.first_group = 0, row.n_groups = 0;
        if (r < r1) row = rows[r];
        uint32_t todo = __ballot_sync(0xffffffffu, row.n_groups != 0u);
        while (todo) {
            const int q = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const unsigned long long base = __shfl_sync(0xffffffffu, row.base, q);
            const uint32_t fg = __shfl_sync(0xffffffffu, row.first_group, q);
            const uint32_t ng = __shfl_sync(0xffffffffu, row.n_groups, q);
            f(base, fg, ng);
        }
    }
}

__global__ void __launch_bounds__(256) sk_compact_hist_kernel(const ChunkRow *__restrict__ rows, const Group *__restrict__ groups,
                                                              uint32_t n_rows, uint32_t S, uint32_t *__restrict__ hist) {
    uint32_t *sh_hist = (uint32_t *)sk_smem;
    for (uint32_t s = threadIdx.x; s < S; s += blockDim.x) sh_hist[s] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for_rows_of_block(rows, n_rows, blockIdx.x, warp, (int)(blockDim.x >> 5), lane,
                      [&](unsigned long long, uint32_t fg, uint32_t ng) {
                          for (uint32_t k = (uint32_t)lane; k < ng; k += 32u) {
                              const Group g = groups[fg + k];
                              if (g.len && g.sample < S) atomicAdd(&sh_hist[g.sample], (uint32_t)g.len);
                          }
                      });
    __syncthreads();
    uint32_t *h = hist + (size_t)blockIdx.x * S;
    for (uint32_t s = threadIdx.x; s < S; s += blockDim.x) h[s] = sh_hist[s];
}

// ---- cols ----------------------------------------------------------------------------------------
// offs[b][s] = sum of hist[b'][s] over b' < b, total[s].  A CTA owns 64 samples (threadIdx.x: consecutive samples,
// coalesced rows of the table) and cuts the blocks into 16 ranges (threadIdx.y): every thread sums its range, the
// ranges' sums are scanned through shared memory, and a second walk over the range writes the prefixes.
constexpr uint32_t CL_S = 64, CL_R = 16;
__global__ void __launch_bounds__(CL_S * CL_R) sk_compact_cols_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ offs,
                                                                    uint32_t nb, uint32_t S, unsigned long long *__restrict__ total) {
    __shared__ unsigned long long part[CL_R][CL_S];
    const uint32_t s = blockIdx.x * CL_S + threadIdx.x, y = threadIdx.y;
    const uint32_t per = (nb + CL_R - 1u) / CL_R;
    const uint32_t b0 = min(y * per, nb), b1 = min(b0 + per, nb);
    const bool live = s < S;
    unsigned long long sum = 0;
    if (live) {
        uint32_t b = b0;
        for (; b + 8u <= b1; b += 8u) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = hist[(size_t)(b + k) * S + s];
#pragma unroll
            for (int k = 0; k < 8; k++) sum += v[k];
        }
        for (; b < b1; b++) sum += hist[(size_t)b * S + s];
    }
    part[y][threadIdx.x] = sum;
    __syncthreads();
    unsigned long long run = 0;
    for (uint32_t k = 0; k < y; k++) run += part[k][threadIdx.x];
    if (live) {
        uint32_t b = b0;
        for (; b + 8u <= b1; b += 8u) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = hist[(size_t)(b + k) * S + s];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                offs[(size_t)(b + k) * S + s] = (uint32_t)run;
                run += v[k];
            }
        }
        for (; b < b1; b++) {
            const uint32_t v = hist[(size_t)b * S + s];
            offs[(size_t)b * S + s] = (uint32_t)run;
            run += v;
        }
        if (y == CL_R - 1u) total[s] = run;
    }
}

// ---- bases ---------------------------------------------------------------------------------------
// slices[s] = {offset, len}: the samples' runs back to back, each starting on a 128-byte line;
// slices[S] = {extent of the compacted buffer, payload bytes}.  One CTA; S is a few hundred.
__global__ void __launch_bounds__(1024) sk_compact_bases_kernel(const unsigned long long *__restrict__ total, uint32_t S,
                                                                unsigned long long *__restrict__ slices, unsigned long long dst_cap,
                                                                DevStats *st) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry_s, pay_s;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0, pay_s = 0;
    __syncthreads();
    for (uint32_t s0 = 0; s0 < S; s0 += 1024u) {
        const uint32_t s = s0 + threadIdx.x;
        const unsigned long long len = s < S ? total[s] : 0ull;
        const unsigned long long padded = (len + 127ull) & ~127ull;
        unsigned long long x = padded, y = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += t;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
        if (lane == 31) ws[w] = x;
        __syncthreads();
        unsigned long long before = carry_s;
        for (uint32_t k = 0; k < w; k++) before += ws[k];
        if (s < S) {
            slices[2 * s] = before + x - padded;
            slices[2 * s + 1] = len;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + x;
        if (lane == 0 && y) atomicAdd(&pay_s, y);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        slices[2 * S] = carry_s;
        slices[2 * S + 1] = pay_s;
        st->compact_extent = carry_s;
        if (carry_s > dst_cap) atomicMax(&st->err_key, ~((0ull << 8) | (unsigned long long)K_OUT_OVERFLOW));
    }
}

// ---- addr ----------------------------------------------------------------------------------------
// One warp per block walks the block's rows in order.  running[s] = where the next piece of sample s goes.
__global__ void __launch_bounds__(32) sk_compact_addr_kernel(const ChunkRow *__restrict__ rows, const Group *__restrict__ groups,
                                                             uint32_t n_rows, uint32_t S, const uint32_t *__restrict__ offs,
                                                             const unsigned long long *__restrict__ slices,
                                                             unsigned long long *__restrict__ piece_dst) {
    unsigned long long *running = (unsigned long long *)sk_smem;
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const uint32_t *o = offs + (size_t)blockIdx.x * S;
    for (uint32_t s = (uint32_t)lane; s < S; s += 32u) running[s] = slices[2 * s] + o[s];
    __syncwarp();
    // The rows are taken 32 at a time (one per lane, coalesced) and the non-empty ones walked in order; the groups
    // of the batch's first eight non-empty rows are fetched at once (and the next 32 rows meanwhile), so that the walk
    // itself waits for no load.
    const uint32_t r0 = blockIdx.x * CB_ROWS, r1 = min(r0 + CB_ROWS, n_rows);
    auto load_rows = [&](uint32_t rb) -> ChunkRow {
        ChunkRow row;
        row.base = 0, row.first_group = 0, row.n_groups = 0;
        const uint32_t r = rb + (uint32_t)lane;
        if (rb < r1 && r < r1) row = rows[r];
        return row;
    };
    auto load_group = [&](uint32_t fg, uint32_t ng, uint32_t k0) -> Group {
        Group g;
        g.sample = 0xFFFFu, g.len = 0;
        const uint32_t k = k0 + (uint32_t)lane;
        if (k < ng) g = groups[fg + k];
        return g;
    };
    ChunkRow row = load_rows(r0);
    for (uint32_t rb = r0; rb < r1; rb += 32u) {
        const ChunkRow row_next = load_rows(rb + 32u);
        uint32_t todo = __ballot_sync(FULL, row.n_groups != 0u);
        // the first groups of up to eight non-empty rows at once (a tile of the warp engine fills one row in four)
        constexpr int PF = 8;
        Group gq[PF];
        {
            uint32_t t = todo;
#pragma unroll
            for (int i = 0; i < PF; i++) {
                gq[i].sample = 0xFFFFu, gq[i].len = 0;
                if (t) {
                    const int q = __ffs((int)t) - 1;
                    t &= t - 1;
                    gq[i] = load_group(__shfl_sync(FULL, row.first_group, q), __shfl_sync(FULL, row.n_groups, q), 0u);
                }
            }
        }
        int qi = 0;
        while (todo) {
            const int q = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const uint32_t fg = __shfl_sync(FULL, row.first_group, q);
            const uint32_t ng = __shfl_sync(FULL, row.n_groups, q);
            Group g;
            g.sample = 0xFFFFu, g.len = 0;
            bool have = false;
#pragma unroll
            for (int i = 0; i < PF; i++)
                if (i == qi) g = gq[i], have = true;
            qi++;
            for (uint32_t k0 = 0; k0 < ng; k0 += 32u) {
                const uint32_t k = k0 + (uint32_t)lane;
                if (k0 || !have) g = load_group(fg, ng, k0);
                const bool act = g.len != 0 && g.sample < S;
                // lanes of one sample, lowest lane = earliest piece; an idle lane is a group of its own
                const uint32_t key = act ? (uint32_t)g.sample : 0x10000u + (uint32_t)lane;
                const uint32_t peers = __match_any_sync(FULL, key);
                uint32_t lower = peers & ((1u << lane) - 1u);
                uint32_t pre = 0;
                while (__any_sync(FULL, lower != 0u)) {  // bytes of the same sample in lower lanes (a few at most)
                    const int j = lower ? __ffs((int)lower) - 1 : lane;
                    const uint32_t v = __shfl_sync(FULL, (uint32_t)g.len, j);
                    if (lower) {
                        pre += v;
                        lower &= lower - 1;
                    }
                }
                unsigned long long base = 0;
                if (act) base = running[g.sample];
                __syncwarp();
                if (act && (peers >> lane) == 1u) running[g.sample] = base + pre + g.len;  // the group's last lane
                __syncwarp();
                if (act) piece_dst[fg + k] = base + pre;
            }
        }
        row = row_next;
    }
}

// ---- move ----------------------------------------------------------------------------------------
// One warp per row.  The row's pieces lie back to back in the input-order stream: the warp brings the whole
// row into its window of shared memory with one TMA bulk copy and then every lane copies one piece from the
// window to its destination (gcopy: whole 32-byte sectors inside the piece, small stores at its two ends,
// which share their sectors with the neighbouring pieces of the sample's run).  Rows that do not fit the
// window, hold more than 32 pieces or start off a 16-byte boundary take the piece-by-piece copy above.
#ifndef SKC_WARP_PIECE
#define SKC_WARP_PIECE 1  // 1: the warp copies a row's pieces one by one (contiguous stores); 0: one lane per piece (gcopy)
#endif
constexpr uint32_t MV_ROW = 12288;               // bytes of a row the window holds
constexpr uint32_t MV_WIN = MV_ROW + 96;         // + slack: gcopy reads whole aligned words past a piece
constexpr int MV_WARPS = 6;                      // 6 x 12.1 KB: three CTAs per SM
__global__ void __launch_bounds__(MV_WARPS * 32) sk_compact_move_kernel(const ChunkRow *__restrict__ rows, const Group *__restrict__ groups,
                                                                        uint32_t n_rows, uint32_t S,
                                                                        const unsigned long long *__restrict__ piece_dst,
                                                                        const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                                        const unsigned long long *__restrict__ slices,
                                                                        unsigned long long dst_cap, unsigned int *ticket) {
    constexpr uint32_t FULL = 0xffffffffu;
    if (slices[2 * S] > dst_cap) return;  // reported by the bases kernel (K_OUT_OVERFLOW): nothing is written
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *win = sk_smem + (uint32_t)warp * MV_WIN;
    uint64_t *mbar = (uint64_t *)(sk_smem + MV_WARPS * MV_WIN) + warp;
    if (lane == 0) mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t parity = 0;
    // Rows are handed out four at a time (one tile of the warp engine) by a ticket counter, in order: the rows
    // in flight at any moment then span a few tens of megabytes of the input-order stream, so the pieces that
    // share a sector or a line of a sample's run are written within microseconds of each other and meet in L2
    // (with a static split of the rows the dirty lines of a whole wave, > L2, went to DRAM half written).
    // The next ticket's rows are fetched while the current ones are copied.
    constexpr uint32_t RB = 4;
    const uint32_t n_batches = (n_rows + RB - 1) / RB;
    auto take = [&]() -> uint32_t {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        return __shfl_sync(FULL, t, 0);
    };
    auto fetch = [&](uint32_t t) -> ChunkRow {
        ChunkRow row;
        row.base = 0, row.first_group = 0, row.n_groups = 0;
        const uint32_t r = t * RB + (uint32_t)lane;
        if (t < n_batches && (uint32_t)lane < RB && r < n_rows) row = rows[r];
        return row;
    };
    uint32_t t_cur = take();
    ChunkRow row = fetch(t_cur);
    while (t_cur < n_batches) {
        const uint32_t t_next = take();
        const ChunkRow row_next = fetch(t_next);
        uint32_t todo = __ballot_sync(FULL, row.n_groups != 0u);
        while (todo) {
            const int q = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const unsigned long long base = __shfl_sync(FULL, row.base, q);
            const uint32_t fg = __shfl_sync(FULL, row.first_group, q);
            const uint32_t ng = __shfl_sync(FULL, row.n_groups, q);
            unsigned long long run = base;  // source offset of the next piece
            for (uint32_t k0 = 0; k0 < ng; k0 += 32u) {
                const uint32_t k = k0 + (uint32_t)lane;
                Group g;
                g.sample = 0xFFFFu, g.len = 0;
                unsigned long long pd = 0;
                if (k < ng) {
                    g = groups[fg + k];
                    pd = piece_dst[fg + k];
                }
                uint32_t incl = g.len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += y;
                }
                const uint32_t total = __shfl_sync(FULL, incl, 31);
                const bool mine = g.len != 0 && g.sample < S;
                if (total <= MV_ROW && (run & 15ull) == 0ull) {
                    // the row (this slice of it) through the window
                    const uint32_t nbytes = (total + 15u) & ~15u;
                    fence_proxy_async();  // this lane's reads of the previous window before the next TMA write
                    __syncwarp();
                    if (lane == 0 && nbytes) {
                        mbar_expect_tx(mbar, nbytes);
                        bulk_g2s(win, src + run, nbytes, mbar);
                    }
                    if (nbytes) {
                        mbar_wait_parked(mbar, parity);
                        parity ^= 1u;
                    }
                    __syncwarp();
#if SKC_WARP_PIECE
                    {   // the warp copies the row's pieces one after the other: 16 destination-aligned bytes per lane
                        // (two aligned 16-byte reads of the window, shifted into place), the piece's first and last
                        // bytes -- which share their 16 bytes with the neighbouring pieces of the sample's run -- one
                        // per lane.  A piece is a few contiguous store wavefronts, not one per lane and sector.
                        const uint32_t live = __ballot_sync(FULL, mine);
                        const uint32_t my_so = incl - g.len;
                        uint32_t rest = live;
#pragma unroll 2
                        while (rest) {
                            const int j = __ffs((int)rest) - 1;
                            rest &= rest - 1;
                            const unsigned long long d_o = __shfl_sync(FULL, pd, j);
                            const uint32_t so = __shfl_sync(FULL, my_so, j), len = __shfl_sync(FULL, (uint32_t)g.len, j);
                            uint8_t *d = dst + d_o;
                            const uint32_t head = min((16u - ((uint32_t)(uintptr_t)d & 15u)) & 15u, len);
                            const uint32_t body = (len - head) >> 4, tail = (len - head) & 15u;
                            const uint32_t sb = so + head;                 // window offset of the first whole unit
                            const uint32_t wo = (sb & 15u) >> 2, bs = (sb & 3u) * 8u;
                            for (uint32_t u = (uint32_t)lane; u < body; u += 32u) {
                                const uint8_t *sp = win + ((sb + 16u * u) & ~15u);
                                const uint4 A = *(const uint4 *)sp, B = *(const uint4 *)(sp + 16);
                                const uint32_t W[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
                                uint4 o4;
                                switch (wo) {  // uniform over the piece
                                    case 0: o4 = make_uint4(__funnelshift_r(W[0], W[1], bs), __funnelshift_r(W[1], W[2], bs), __funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs)); break;
                                    case 1: o4 = make_uint4(__funnelshift_r(W[1], W[2], bs), __funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs)); break;
                                    case 2: o4 = make_uint4(__funnelshift_r(W[2], W[3], bs), __funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs), __funnelshift_r(W[5], W[6], bs)); break;
                                    default: o4 = make_uint4(__funnelshift_r(W[3], W[4], bs), __funnelshift_r(W[4], W[5], bs), __funnelshift_r(W[5], W[6], bs), __funnelshift_r(W[6], W[7], bs)); break;
                                }
                                *(uint4 *)(d + head + 16u * u) = o4;
                            }
                            // lanes 0..15: the bytes before the first whole unit; lanes 16..31: those behind the last
                            const uint32_t e = (uint32_t)lane & 15u;
                            const bool hi = lane >= 16;
                            if (e < (hi ? tail : head)) {
                                const uint32_t o = hi ? head + 16u * body + e : e;
                                d[o] = win[so + o];
                            }
                        }
                    }
#else
                    if (mine) gcopy(dst + pd, win, incl - g.len, g.len);
#endif
                    __syncwarp();
                } else {
                    const unsigned long long ps = run + incl - g.len;
                    const uint32_t live = __ballot_sync(FULL, mine);
                    const uint32_t n_here = min(32u, ng - k0);
                    for (uint32_t j = 0; j < n_here; j++) {
                        const unsigned long long so = __shfl_sync(FULL, ps, (int)j), d_o = __shfl_sync(FULL, pd, (int)j);
                        const uint32_t len = __shfl_sync(FULL, (uint32_t)g.len, (int)j);
                        if ((live >> j) & 1u) warp_copy_piece(src, dst, so, d_o, len, lane);
                    }
                }
                run += total;
            }
        }
        t_cur = t_next;
        row = row_next;
    }
}

// ------------------------------------------------------------------------------------------------
// launcher: tables -> compacted buffer + slices.  Work area (device): hist u32[nb][S], offs u32[nb][S],
// total u64[S].  Returns the number of launches, < 0 on error.
// ------------------------------------------------------------------------------------------------
uint32_t compact_blocks(uint32_t n_rows) { return (n_rows + CB_ROWS - 1) / CB_ROWS; }
uint64_t compact_work_bytes(uint32_t max_rows, uint32_t S) {
    const uint64_t nb = compact_blocks(max_rows) + 1;
    return nb * S * 4ull * 2ull + (uint64_t)S * 8ull + 256 + 64;  // hist, offs, totals, the move kernel's ticket
}
int launch_compact(const ChunkRow *rows, const Group *groups, uint32_t n_rows, uint32_t S, const uint8_t *src, uint8_t *dst,
                   uint64_t dst_cap, void *work, unsigned long long *slices, unsigned long long *piece_dst, DevStats *st,
                   int sm_count, void *stream_, const char **err) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const uint32_t nb = compact_blocks(n_rows);
    if (S == 0 || S > 4096) {
        *err = "per-sample compaction handles up to 4096 samples";
        return -1;
    }
    uint32_t *hist = (uint32_t *)work;
    uint32_t *offs = hist + (size_t)(nb + 1) * S;
    unsigned long long *total = (unsigned long long *)(offs + (size_t)(nb + 1) * S);
    total = (unsigned long long *)(((uintptr_t)total + 15) & ~(uintptr_t)15);
    unsigned int *ticket = (unsigned int *)(total + S + 2);
    if (nb == 0) {  // nothing was emitted: empty slices
        cudaMemsetAsync(slices, 0, (size_t)(S + 1) * 16, stream);
        return 0;
    }
    sk_compact_hist_kernel<<<nb, 256, S * 4, stream>>>(rows, groups, n_rows, S, hist);
    sk_compact_cols_kernel<<<(S + CL_S - 1) / CL_S, dim3(CL_S, CL_R), 0, stream>>>(hist, offs, nb, S, total);
    sk_compact_bases_kernel<<<1, 1024, 0, stream>>>(total, S, slices, dst_cap, st);
    sk_compact_addr_kernel<<<nb, 32, S * 8, stream>>>(rows, groups, n_rows, S, offs, slices, piece_dst);
    const unsigned want = (n_rows + 4u * MV_WARPS - 1u) / (4u * MV_WARPS);  // MV_WARPS warps x 4 rows per CTA and ticket
    const unsigned grid = std::max(1u, std::min<unsigned>((unsigned)sm_count * 3u, want));
    const int mv_smem = (int)(MV_WARPS * MV_WIN + MV_WARPS * 8 + 16);
    // (a per-device attribute: set on every launch, the process may drive several GPUs)
    cudaFuncSetAttribute(sk_compact_move_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mv_smem);
    cudaMemsetAsync(ticket, 0, 4, stream);
    sk_compact_move_kernel<<<grid, MV_WARPS * 32, mv_smem, stream>>>(rows, groups, n_rows, S, piece_dst, src, dst, slices, dst_cap, ticket);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    return 5;
}

}  // namespace sk
