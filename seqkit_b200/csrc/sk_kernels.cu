// sk_kernels.cu -- the single-pass chunk engine, hand-written for sm_100a.
//
// One persistent kernel template implements every operator of the path:
//   OP_SCAN / OP_TRIM / OP_MASK / OP_ADDBC / OP_DEMUX1 / OP_DEMUX2.
// Input bytes are read from HBM once (TMA bulk copy into shared memory), output bytes are written
// once (TMA bulk store of a shared staging image).  Record framing is by global line index
// (decoupled look-back over per-chunk line counts), exactly like the reference's four read_line
// calls per record (common.rs:106-112).
//
// The kernel is bound by issued instructions and barrier stalls, not by DRAM (profiles/), so the
// per-record logic runs one thread per record (no redundant lanes), independent sub-tasks of a
// record (quality trim | header search + barcode match) run on different warps at the same time,
// the barcode match is two hash probes plus a distance check of the few candidates, and the
// per-sample grouping of a chunk's output is a bit-mask exchange instead of a serial walk.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sk_internal.h"

namespace sk {

extern __shared__ __align__(128) unsigned char sk_smem[];

}  // namespace sk
#include "sk_device.cuh"
namespace sk {


// ------------------------------------------------------------------------------------------------
// the chunk engine
// ------------------------------------------------------------------------------------------------
struct Misc {
    uint64_t mbar;
    uint64_t g0;        // global index of the first line that starts in this chunk
    uint64_t out_base;  // where this chunk's output goes
    uint32_t chunk;
    uint32_t chunk_out;
    uint32_t cta_total, cta_ident, cta_slow;
    uint32_t n_slow;
    uint32_t n_groups;
    uint32_t scratch[40];
};

#ifdef SK_PHASE_TIMING
#define SK_T(i)                                              \
    do {                                                     \
        if (tid == 0) {                                      \
            const long long t_now = clock64();               \
            ph[i] += (unsigned long long)(t_now - t_prev);   \
            t_prev = t_now;                                  \
        }                                                    \
    } while (0)
#else
#define SK_T(i) do { } while (0)
#endif

template <class Cfg, int OP, typename WT>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MIN_CTAS) sk_chunk_kernel(const __grid_constant__ KParams p) {
    constexpr int NT = Cfg::NT, NW = NT / 32, MAXREC = Cfg::MAXREC;
    constexpr bool WIDE = sizeof(WT) == 8;
    constexpr int NWMAX = WIDE ? 16 : 8;
    constexpr bool IS_DEMUX = (OP == OP_DEMUX1 || OP == OP_DEMUX2);
    constexpr bool ORDERED = (OP == OP_TRIM || OP == OP_MASK || OP == OP_ADDBC);
    constexpr bool HAS_OUT = ORDERED || IS_DEMUX;
    // OP_SCAN keeps no per-record plan and no staging image: its line table takes the staging area, so
    // that index reads and barcode files (very short records) do not hit the density limit.
    constexpr int MAXLINES = (OP == OP_SCAN) ? (Cfg::STAGE / 2 - 8) : Cfg::MAXLINES;
    constexpr int PIECE_BYTES = Cfg::PPL * 16;

    const uint32_t S = IS_DEMUX ? p.sheet.S : 0u;
    const SmemLayout &SL = p.sl;
    uint8_t *win = sk_smem + SL.win;
    uint8_t *stage = sk_smem + SL.stage;
    uint16_t *ls = (uint16_t *)(sk_smem + (OP == OP_SCAN ? SL.stage : SL.ls));
    uint32_t *r_aux = (uint32_t *)(sk_smem + SL.rec);  // ADDBC: barcode offset; demux slow layout: group totals
    uint16_t *r_outoff = (uint16_t *)(r_aux + MAXREC);
    uint16_t *r_outlen = r_outoff + MAXREC;
    uint16_t *r_cut0 = r_outlen + MAXREC;
    uint16_t *r_cut1 = r_cut0 + MAXREC;
    uint16_t *r_alen = r_cut1 + MAXREC;
    uint16_t *r_blen = r_alen + MAXREC;
    uint16_t *r_k = r_blen + MAXREC;
    uint16_t *r_body = r_k + MAXREC;
    uint16_t *r_taglen = r_body + MAXREC;
    int16_t *r_sample = (int16_t *)(r_taglen + MAXREC);
    uint8_t *r_mode = (uint8_t *)(r_sample + MAXREC);
    uint8_t *r_flags = r_mode + MAXREC;
    static_assert(REC_BYTES >= 4 + 2 * 10 + 2, "per-record plan fields");
    WT *sh_umask = (WT *)(sk_smem + SL.umask);
    uint8_t *sh_lut = sk_smem + SL.lut;
    uint32_t *ccount = (uint32_t *)(sk_smem + SL.ccount);
    uint32_t *masks = (uint32_t *)(sk_smem + SL.masks);  // per sample: 128 record bits of the current chunk
    uint32_t *gbase_sm = (uint32_t *)(sk_smem + SL.gtot);
    uint32_t *hcls = (uint32_t *)(sk_smem + SL.hcls);
    uint16_t *slow_list = (uint16_t *)(sk_smem + SL.slow);
    Misc *M = (Misc *)(sk_smem + SL.misc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevStats *st = p.stats;
    const bool use_hidx = (OP == OP_DEMUX1) && p.sheet.hidx.n_classes != 0 && p.n_index == 0;

    // ---- one-time per CTA: mbarrier, sheet -> shared memory
    if (tid == 0) {
        mbar_init(&M->mbar, 1);
        M->cta_total = 0;
        M->cta_ident = 0;
        M->cta_slow = 0;
    }
    if (IS_DEMUX) {
        const WT *gu = (const WT *)p.sheet.umask;
        for (uint32_t i = tid; i < S; i += NT) {
            sh_umask[i] = gu[i];
            ccount[i] = 0;
        }
        for (uint32_t i = tid; i < 4 * S; i += NT) masks[i] = 0;
        for (uint32_t i = tid; i < 256; i += NT) sh_lut[i] = p.sheet.lut[i];
        if (OP == OP_DEMUX1) {
            const uint32_t nh = p.sheet.hidx.n_classes * HIDX_CLS_ROWS * p.sheet.hidx.nwp;
            for (uint32_t i = tid; i < nh; i += NT) hcls[i] = p.sheet.hidx.cls[i];
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

#ifdef SK_PHASE_TIMING
    unsigned long long ph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
#endif
    uint32_t parity = 0, flip = 0;
    bool store_pending = false;  // thread 0: a TMA store of the staging image may still be reading it
    if (tid == 0) {
        M->chunk = atomicAdd(&st->ticket, 1u);
        M->n_slow = 0;
    }
    __syncthreads();
    for (;;) {
        // ---- P0 ticket (taken by thread 0 at the end of the previous chunk, published by its last barrier)
        const uint32_t c = M->chunk;
        if (c >= p.n_chunks) break;
        SK_T(0);  // ticket

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
            fence_proxy_async();  // order earlier generic-proxy accesses to `win` before the async write
            mbar_expect_tx(&M->mbar, bulk);
            bulk_g2s(win, p.in + w0, bulk, &M->mbar);
        }
        if (bulk != (uint32_t)Cfg::WIN_MAX && tid < 16) {  // tail bytes; zero padding so the last partial piece reads defined data
            const uint32_t o = bulk + tid;
            if (o < (uint32_t)Cfg::WIN_MAX) win[o] = (o < wlen) ? p.in[w0 + o] : (uint8_t)0;
        }
        if (bulk) {
            mbar_wait(&M->mbar, parity);
            parity ^= 1;
        }
        if (bulk != (uint32_t)Cfg::WIN_MAX) __syncthreads();  // the tail bytes (last windows of the stream only)
        SK_T(1);  // window load

        // ---- P2 newline scan: PPL*16 contiguous bytes per thread, kept as PPL piece maps
        // A '\n' at window offset q starts a line at q+1.  Lines that start before the chunk belong to
        // the previous chunk (q+1 >= cs); a '\n' that is the last byte of the buffer starts nothing.
        const uint32_t ls_lo = cs ? cs - 1 : 0;
        const uint32_t ls_hi = at_end ? (wlen ? wlen - 1 : 0) : wlen;
        const uint32_t o0 = (uint32_t)tid * PIECE_BYTES;
        uint32_t ymap[Cfg::PPL];
        uint32_t cnt_all = 0, cnt_chunk = 0, hib = 0;
        if (o0 >= ls_lo && o0 + PIECE_BYTES <= ls_hi && o0 + PIECE_BYTES <= wlen &&
            (o0 + PIECE_BYTES <= ce - 1 || o0 >= ce - 1)) {
            // interior thread: no range boundary inside its bytes
#pragma unroll
            for (int q = 0; q < Cfg::PPL; q++) {
                const uint4 v = *(const uint4 *)(win + o0 + q * 16);
                const uint32_t y = nl_map(v);
                hib |= (v.x | v.y | v.z | v.w);
                cnt_all += __popc(y);
                ymap[q] = y;
            }
            cnt_chunk = o0 >= ce - 1 ? 0u : cnt_all;
        } else {
#pragma unroll
            for (int q = 0; q < Cfg::PPL; q++) {
                const uint32_t o = o0 + q * 16;
                uint32_t y = 0;
                if (o < wlen) {
                    const uint4 v = *(const uint4 *)(win + o);
                    y = nl_map(v);
                    hib |= (v.x | v.y | v.z | v.w);  // bytes past wlen in the last piece are zero
                    if (o < ls_lo || o + 16 > ls_hi) y = map_clip(y, o, ls_lo, ls_hi);
                    const uint32_t n = __popc(y);
                    cnt_all += n;
                    if (o + 16 <= ce - 1) cnt_chunk += n;
                    else if (o < ce - 1) cnt_chunk += __popc(map_clip(y, o, 0, ce - 1));
                }
                ymap[q] = y;
            }
        }
        if (__any_sync(0xffffffffu, (hib & 0x80808080u) != 0) && lane == 0) atomicOr(&st->flags, F_NON_ASCII);

        uint32_t tot;
        const uint32_t pre = block_excl_scan<NT>((cnt_chunk << 16) | cnt_all, M->scratch, flip, tot);
        const uint32_t extra = (c0 == 0) ? 1u : 0u;      // the line that starts at byte 0
        const uint32_t nls = (tot & 0xFFFFu) + extra;    // line starts in [cs, wlen)
        const uint32_t nls_chunk = (tot >> 16) + extra;  // ... of which inside the chunk

        SK_T(2);  // scan + block scan
        // ---- P3 look-back for the global line index (warp 0) + line-start table (everybody)
        if (warp == 0) {
            const uint64_t excl = lookback(p.tile_lines, c, nls_chunk, lane);
            if (lane == 0) {
                M->g0 = excl;
                if (c == p.n_chunks - 1) st->n_lines = excl + nls_chunk;
            }
        }
        {
            uint32_t idx = (pre & 0xFFFFu) + extra;
            if (tid == 0 && extra) ls[0] = 0;
#pragma unroll
            for (int q = 0; q < Cfg::PPL; q++) {
                const uint32_t y = ymap[q];
                if (y) {
                    const uint32_t o = o0 + q * 16;
                    const uint32_t y2 = y & (y - 1);
                    const uint32_t i1 = __ffs(y) - 1;
                    const uint32_t k1 = 4u * (i1 & 7u) + (i1 >> 3);  // piece offset of a map bit
                    if (y2 == 0) {  // one newline in the piece
                        if (idx < (uint32_t)MAXLINES) ls[idx] = (uint16_t)(o + k1 + 1u);
                        idx++;
                    } else if ((y2 & (y2 - 1)) == 0) {  // two ("\n+\n" puts two in one piece for most records)
                        const uint32_t i2 = __ffs(y2) - 1;
                        const uint32_t k2 = 4u * (i2 & 7u) + (i2 >> 3);
                        const uint32_t ka = k1 < k2 ? k1 : k2, kb = k1 < k2 ? k2 : k1;
                        if (idx < (uint32_t)MAXLINES) ls[idx] = (uint16_t)(o + ka + 1u);
                        if (idx + 1 < (uint32_t)MAXLINES) ls[idx + 1] = (uint16_t)(o + kb + 1u);
                        idx += 2;
                    } else {
                        for (uint32_t k = 0; k < 16; k++)
                            if ((y >> map_bit(k)) & 1u) {
                                if (idx < (uint32_t)MAXLINES) ls[idx] = (uint16_t)(o + k + 1u);
                                idx++;
                            }
                    }
                }
            }
            // sentinels: a line index past the last line start reads as "end of window" (EOF records)
            if (tid >= NT - 8) {
                const uint32_t k = nls + (uint32_t)(tid - (NT - 8));
                if (k < (uint32_t)MAXLINES + 8u) ls[k] = (uint16_t)wlen;
            }
        }
        __syncthreads();
        SK_T(3);  // line table + look-back (lines)
#define LB(x) ((uint32_t)ls[(x)])

        // ---- P4 which records does this chunk own?
        const uint64_t g0 = M->g0;
        const uint32_t lpr = p.lpr;
        const uint32_t j0 = (lpr - (uint32_t)(g0 & (lpr - 1))) & (lpr - 1);  // lpr is 2 or 4
        const uint64_t rec0 = (g0 + j0) / lpr;
        uint32_t nrec = j0 < nls_chunk ? (nls_chunk - 1 - j0) / lpr + 1 : 0;
        if (rec0 >= p.rec_limit) nrec = 0;
        else if ((uint64_t)nrec > p.rec_limit - rec0) nrec = (uint32_t)(p.rec_limit - rec0);
        if (nrec) {
            unsigned chunk_err = 0;
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
                if (need >= (uint32_t)MAXLINES || (OP != OP_SCAN && nrec > (uint32_t)MAXREC)) chunk_err = K_TOO_DENSE;
            }
            // per-record tables hold sk_limits.max_records entries (OP_SCAN guards its own table, trim / mask have none)
            bool too_many = false;
            if (!chunk_err && nrec) {
                if (IS_DEMUX) too_many = rec0 + nrec > p.max_records;
                if (OP == OP_ADDBC && p.ext_stats[0] && p.ext_stats[0]->n_records) {
                    const unsigned long long nb = p.ext_stats[0]->n_records, hi = rec0 + nrec - 1u;
                    too_many = (hi < nb ? hi : nb - 1ull) >= p.max_records;
                }
            }
            if (too_many) {
                if (tid == 0) report_err(st, p.max_records, K_TOO_MANY);
                nrec = 0;
            } else if (chunk_err) {
                if (tid == 0) report_err(st, rec0, chunk_err);
                nrec = 0;
            }
        }
        if (tid == 0 && nrec) {
            atomicAdd(&st->n_records, (unsigned long long)nrec);
            atomicMax(&st->consumed, (unsigned long long)(w0 + LB(j0 + nrec * lpr)));
        }

        // ---- P5 plan: the reference's per-record logic, one thread per record.  In the fused
        // trim+demultiplex operators the two independent halves of a record's plan (quality trim |
        // header search and barcode match) run on the two halves of the CTA at the same time.
        const bool fused = IS_DEMUX && p.fused_trim >= 0;
        const uint32_t tl = fused ? (uint32_t)tid % (NT / 2) : (uint32_t)tid;  // record slot of this thread
        const uint32_t tstride = fused ? NT / 2 : NT;
        const bool do_trim = fused && tid < NT / 2;
        const bool do_main = !fused || tid >= NT / 2;
        uint32_t my_total = 0, my_ident = 0;

        if (do_trim) {
            for (uint32_t r = tl; r < nrec; r += tstride) {
                const uint32_t j = j0 + r * lpr;
                const uint32_t L0 = LB(j), L1 = LB(j + 1), L2 = LB(j + 2), L3 = LB(j + 3), L4 = LB(j + 4);
                uint8_t mode = B_FAIL;
                uint32_t kk = 0, body = 0;
                if (L1 > L0 && win[L1 - 1] == '\n') {
                    if (!plan_trim_body(win, L1, L2, L3, L4, p.fused_trim, mode, kk, body)) mode = B_FAIL;
                }
                r_k[r] = (uint16_t)kk;
                r_body[r] = (uint16_t)(body > 0xFFFFu ? 0xFFFFu : body);
                r_mode[r] = mode;
            }
        }
        if (do_main) {
            for (uint32_t r = tl; r < nrec; r += tstride) {
                const uint32_t j = j0 + r * lpr;
                const uint64_t rec = rec0 + r;
                const uint32_t L0 = LB(j), L1 = LB(j + 1), L2 = LB(j + 2);
                const uint32_t L3 = lpr == 4 ? LB(j + 3) : L2, L4 = lpr == 4 ? LB(j + 4) : L2;
                const uint32_t Lend = lpr == 4 ? L4 : L2;
                uint8_t mode = B_NONE, flags = 0;
                uint32_t kk = 0, outlen = 0, alen = 0, cut0 = 0, cut1 = 0, blen = 0, taglen = 0, ext = 0;
                int sample = -1;

                if (OP == OP_SCAN) {
                    // (seq_off, seq_len after trim_end, flags) of an index read / barcode record
                    uint32_t sl = L2 - L1;
                    while (sl > 0 && is_ws(win[L1 + sl - 1])) sl--;
                    uint16_t fl = 0;
                    if (L1 > L0 && win[L0] == '@') fl |= RR_L0_AT;
                    if (L1 > L0 && win[L0] == '>') fl |= RR_L0_GT;
                    if (lpr == 4 && L3 > L2 && win[L2] == '+') fl |= RR_L2_PLUS;
                    if (sl > 0xFFFFu) {
                        fl |= RR_LONG;
                        sl = 0xFFFFu;
                    }
                    if (p.head_char && !(L1 > L0 && win[L0] == (uint8_t)p.head_char)) report_err(st, rec, K_MIXED);
                    if (rec < p.scan_cap) {
                        RecRef rr;
                        rr.seq_off = (uint32_t)(w0 + L1);
                        rr.seq_len = (uint16_t)sl;
                        rr.flags = fl;
                        p.scan_out[rec] = rr;
                    }
                    continue;
                } else if (OP == OP_TRIM) {
                    if (win[L0] != '@') {  // fasta_trim_by_quality.rs:20-22
                        report_err(st, rec, K_BAD_HEADER);
                    } else {
                        uint32_t body;
                        if (!plan_trim_body(win, L1, L2, L3, L4, (int)p.min_baseq, mode, kk, body)) {
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
                    bool bc_long = false;
                    const uint64_t nb = p.ext_stats[0] ? p.ext_stats[0]->n_records : 0;
                    if (nb) {
                        const uint64_t bi = rec < nb ? rec : nb - 1;
                        const RecRef rr = p.ext_tab[0][bi];
                        bl = rr.seq_len;
                        bo = rr.seq_off;
                        bc_long = (rr.flags & RR_LONG) != 0;  // a barcode line of 64 KiB or more
                    }
                    taglen = 4 + bl;  // " BC:" + barcode
                    ext = bo;
                    cut1 = alen;
                    if (bc_long) {
                        report_err(st, rec, K_TOO_LONG);
                        taglen = 0;
                    } else if (h != (uint8_t)p.head_char) {
                        // the reference prints the BC'd header and then stops (:33 before :41-43); the host
                        // reproduces that line, the kernel only reports where.
                        report_err(st, rec, (h == '@' || h == '>') ? K_MIXED : K_BAD_FASTX_LINE);
                        taglen = 0;
                    } else {
                        mode = B_VERBATIM;
                        outlen = alen + taglen + 1 + (Lend - L1);
                    }
                } else if (OP == OP_DEMUX1) {
                    // fasta_demultiplex.rs:117-194: validate, locate the barcode, match, decide.
                    flags = RF_DEAD;
                    if (win[L0] != '@') {  // :118-120
                        report_err(st, rec, K_BAD_HEADER);
                    } else if (fused && !(L1 > L0 && win[L1 - 1] == '\n')) {
                        report_err(st, rec, K_TRUNC_FUSED);
                    } else if (p.n_index) {
                        flags = RF_SLOW;  // --index route: the barcode lives in other streams (:126-136)
                    } else {
                        uint32_t stp;
                        if (!bc_find(win, sh_lut, L0, L1, stp)) {  // :138-141
                            report_err(st, rec, K_NO_BC);
                        } else {
                            cut0 = stp - L0;
                            const uint32_t bs = stp + 4;
                            const uint32_t e = bc_run_end(win, sh_lut, bs + 1, L1);  // greedy class run (:38)
                            cut1 = e - L0;
                            if (e - bs != p.sheet.L) {  // :148-150
                                report_err(st, rec, K_BC_LEN);
                            } else if (!use_hidx) {
                                flags = RF_SLOW;
                            } else {
                                flags = 0;
                                uint32_t lowest, best, last;
                                hidx_match<NWMAX>(win, bs, p.sheet.hidx, hcls, lowest, best, last);
                                my_total++;                // :169
                                if (lowest <= 1u) {        // :172
                                    if (best == last) {    // :173-178
                                        sample = (int)best;
                                        my_ident++;
                                        atomicAdd(&ccount[best], 1u);
                                    } else {  // :184-188
                                        sample = -2;
                                        const uint32_t ei = atomicAdd(&st->n_events, 1u);
                                        if (ei < p.events_cap) {
                                            Event ev;
                                            ev.record = (uint32_t)rec;
                                            ev.bc_off = (uint32_t)(w0 + bs);
                                            ev.bc_off2 = 0xFFFFFFFFu;
                                            ev.best = (int16_t)best;
                                            ev.last = (int16_t)last;
                                            ev.mismatches = lowest;
                                            p.events[ei] = ev;
                                        } else {
                                            atomicOr(&st->flags, F_EVENTS_OVERFLOW);
                                        }
                                    }
                                }
                            }
                        }
                    }
                    if (flags & RF_SLOW) slow_list[atomicAdd(&M->n_slow, 1u)] = (uint16_t)r;
                } else if (OP == OP_DEMUX2) {
                    // fasta_demultiplex.rs:215-237: mate 2 of an assigned pair
                    sample = rec < p.r1_stats->n_records ? (int)p.assign[rec] : -1;
                    if (sample >= 0 && p.out) {
                        uint32_t c0h = L1, c1h = L1;
                        if (!p.n_index) {  // :219-227
                            uint32_t a;
                            if (bc_find(win, sh_lut, L0, L1, a)) {
                                c0h = a;
                                c1h = bc_run_end(win, sh_lut, a + 5, L1);
                            }
                        }
                        header_pieces(win, L0, L1, c0h, c1h, alen, blen);  // :229
                        cut1 = c1h - L0;
                        const uint32_t ul = popcw<WT>(sh_umask[sample]);
                        taglen = ul ? 5 + ul : 0;
                        flags = 0;
                        if (fused && !(L1 > L0 && win[L1 - 1] == '\n')) {
                            report_err(st, rec, K_TRUNC_FUSED);
                            flags = RF_DEAD;
                        }
                    } else {
                        flags = RF_DEAD;
                    }
                }
                r_outlen[r] = (uint16_t)(outlen > 0xFFFFu ? 0xFFFFu : outlen);
                if (OP == OP_ADDBC) r_aux[r] = ext;
                r_cut0[r] = (uint16_t)cut0;
                r_cut1[r] = (uint16_t)cut1;
                r_alen[r] = (uint16_t)alen;
                r_blen[r] = (uint16_t)blen;
                r_taglen[r] = (uint16_t)taglen;
                r_sample[r] = (int16_t)sample;
                r_flags[r] = flags;
                if (!fused) {
                    r_k[r] = (uint16_t)kk;
                    r_mode[r] = mode;
                }
            }
        }
        __syncthreads();
        SK_T(4);  // plan tasks

        if (OP == OP_DEMUX1 && M->n_slow) {
            // ---- P5b brute-force matcher, one warp per queued record (fasta_demultiplex.rs:126-194): used
            // for the --index route and for sheets the pigeonhole index cannot represent.  Lanes encode
            // the observed barcode into bit planes with ballots, then split the samples.
            const WT *planes = (const WT *)p.sheet.planes;
            const uint32_t n_slow = M->n_slow;
            for (uint32_t si = warp; si < n_slow; si += NW) {
                const uint32_t r = slow_list[si];
                const uint32_t j = j0 + r * lpr;
                const uint64_t rec = rec0 + r;
                const uint32_t L0 = LB(j);
                const uint32_t Lb = p.sheet.L;
                uint32_t bs = 0, bclen = 0, sep = 0;
                RecRef ir0{0, 0, 0}, ir1{0, 0, 0};
                bool ok = true;
                if (p.n_index) {
                    for (uint32_t q = 0; q < p.n_index && ok; q++) {
                        if (rec >= p.ext_stats[q]->n_records) {
                            ok = false;
                            break;
                        }
                        const RecRef t = p.ext_tab[q][rec];
                        if (!(t.flags & RR_L0_AT) || !(t.flags & RR_L2_PLUS)) ok = false;  // :130,:134
                        if (q == 0) ir0 = t;
                        else ir1 = t;
                    }
                    if (!ok) {
                        if (lane == 0) report_err(st, rec, K_INDEX_ASSERT);
                    } else {
                        bclen = ir0.seq_len;
                        if (p.n_index == 2) {
                            sep = bclen ? 1u : 0u;  // '+' only if the barcode so far is non-empty (:128)
                            bclen += sep + ir1.seq_len;
                        }
                    }
                    if (ok && bclen != Lb) {  // :148-150
                        if (lane == 0) report_err(st, rec, K_BC_LEN);
                        ok = false;
                    }
                } else {
                    bs = L0 + r_cut0[r] + 4;  // class run and length were checked by the planning thread
                }
                int sample = -1;
                if (ok) {
                    auto obs = [&](uint32_t q) -> uint8_t {
                        if (!p.n_index) return win[bs + q];
                        if (q < ir0.seq_len) return p.ext_data[0][(uint64_t)ir0.seq_off + q];
                        if (q < ir0.seq_len + sep) return (uint8_t)'+';
                        return p.ext_data[1][(uint64_t)ir1.seq_off + (q - ir0.seq_len - sep)];
                    };
                    WT o0b = 0, o1b = 0, o2b = 0;
                    for (uint32_t base = 0; base < Lb; base += 32) {
                        const uint32_t q = base + lane;
                        const uint32_t code = q < Lb ? (uint32_t)sh_lut[obs(q)] : 0u;
                        o0b |= (WT)__ballot_sync(0xffffffffu, code & 1u) << base;
                        o1b |= (WT)__ballot_sync(0xffffffffu, code & 2u) << base;
                        o2b |= (WT)__ballot_sync(0xffffffffu, code & 4u) << base;
                    }
                    // each lane scans samples lane, lane+32, ...; then merge (lowest, first, last)
                    uint32_t lowest = 0xFFFFFFFFu, best = 0xFFFFFFFFu, last = 0;
                    for (uint32_t s2 = lane; s2 < S; s2 += 32) {
                        const WT p0 = planes[4 * s2], p1 = planes[4 * s2 + 1], p2 = planes[4 * s2 + 2],
                                 care = planes[4 * s2 + 3];
                        const uint32_t d = popcw<WT>(((o0b ^ p0) | (o1b ^ p1) | (o2b ^ p2)) & care);
                        if (d < lowest) {
                            lowest = d;
                            best = s2;
                            last = s2;
                        } else if (d == lowest) {
                            last = s2;
                        }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        const uint32_t l2 = __shfl_xor_sync(0xffffffffu, lowest, o);
                        const uint32_t b2 = __shfl_xor_sync(0xffffffffu, best, o);
                        const uint32_t a2 = __shfl_xor_sync(0xffffffffu, last, o);
                        if (l2 < lowest) {
                            lowest = l2;
                            best = b2;
                            last = a2;
                        } else if (l2 == lowest) {
                            best = b2 < best ? b2 : best;
                            last = a2 > last ? a2 : last;
                        }
                    }
                    if (lane == 0) {
                        atomicAdd(&M->cta_total, 1u);  // :169
                        atomicAdd(&M->cta_slow, 1u);
                        if (lowest <= 1u) {            // :172
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
                                    ev.bc_off = p.n_index ? ir0.seq_off : (uint32_t)(w0 + bs);
                                    ev.bc_off2 = p.n_index == 2 ? ir1.seq_off : 0xFFFFFFFFu;
                                    ev.best = (int16_t)best;
                                    ev.last = (int16_t)last;
                                    ev.mismatches = lowest;
                                    p.events[ei] = ev;
                                } else {
                                    atomicOr(&st->flags, F_EVENTS_OVERFLOW);
                                }
                            }
                        }
                    }
                    sample = __shfl_sync(0xffffffffu, sample, 0);
                    if (sample >= 0 && p.n_index) {
                        // --index route: the UMI bytes come from other streams; park them in the side table
                        // (the header route gathers them from the window when the record is assembled)
                        const WT um = sh_umask[sample];
                        for (uint32_t base = 0; base < Lb; base += 32) {
                            const uint32_t q = base + lane;
                            if (q < Lb && ((um >> q) & 1u)) {
                                const uint32_t rank = popcw<WT>(um & (((WT)1 << q) - 1));
                                p.umi[rec * p.sheet.Umax + rank] = obs(q);
                            }
                        }
                    }
                }
                if (lane == 0) {
                    r_sample[r] = (int16_t)sample;
                    r_flags[r] = ok ? (uint8_t)RF_SLOW : (uint8_t)RF_DEAD;
                }
            }
            __syncthreads();
        }
        SK_T(5);  // slow matcher

        // ---- P5c finish the plan (one thread per record): kept header pieces, tag, body, output length
        if (IS_DEMUX) {
            for (uint32_t r = tid; r < nrec; r += NT) {
                const int sample = r_sample[r];
                uint32_t outlen = 0;
                if (sample >= 0 && !(r_flags[r] & RF_DEAD)) {
                    const uint32_t j = j0 + r * lpr;
                    const uint64_t rec = rec0 + r;
                    const uint32_t L0 = LB(j), L1 = LB(j + 1), L4 = LB(j + 4);
                    uint32_t alen = r_alen[r], blen = r_blen[r], taglen = r_taglen[r];
                    if (OP == OP_DEMUX1) {
                        const uint32_t c0h = p.n_index ? L1 : L0 + r_cut0[r], c1h = p.n_index ? L1 : L0 + r_cut1[r];
                        header_pieces(win, L0, L1, c0h, c1h, alen, blen);  // drain (:145) + trim_end (:206)
                        const uint32_t ul = popcw<WT>(sh_umask[sample]);
                        taglen = ul ? 5 + ul : 0;  // " UMI:" + umi (:207)
                        r_alen[r] = (uint16_t)alen;
                        r_blen[r] = (uint16_t)blen;
                        r_cut1[r] = (uint16_t)(c1h - L0);
                        r_taglen[r] = (uint16_t)taglen;
                    }
                    uint32_t body = L4 - L1;  // three lines verbatim (:209-212)
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
                        report_err(st, rec, K_SEQ_SHORT);
                        if (OP == OP_DEMUX1) r_sample[r] = -1;
                        mode = B_NONE;
                    } else if (!p.out) {
                        mode = B_NONE;  // dry run: count only (:77-78,:179)
                    } else {
                        outlen = alen + blen + taglen + 1 + body;
                    }
                    r_mode[r] = mode;
                }
                r_outlen[r] = (uint16_t)(outlen > 0xFFFFu ? 0xFFFFu : outlen);
            }
            if (OP == OP_DEMUX1) {
                // counters of this thread's records (fasta_demultiplex.rs:169,177), one atomic per warp
                const uint32_t wt = __reduce_add_sync(0xffffffffu, my_total), wi = __reduce_add_sync(0xffffffffu, my_ident);
                if (lane == 0 && wt) atomicAdd(&M->cta_total, wt);
                if (lane == 0 && wi) atomicAdd(&M->cta_ident, wi);
            }
            __syncthreads();  // plans of other threads' records are read below
        }

        if (HAS_OUT) {
            // ---- P6 layout of the chunk's output
            uint32_t chunk_out = 0;
            SK_T(6);  // plan C
            if (tid == 0 && store_pending) {  // the previous chunk's TMA store must have read the staging image
                bulk_wait_read0();
                store_pending = false;
            }
            if (ORDERED) {
                uint32_t run = 0;
                for (uint32_t rb = 0; rb < nrec; rb += NT) {  // (one pass unless the chunk is very dense)
                    const uint32_t r = rb + tid;
                    const uint32_t mine = r < nrec ? r_outlen[r] : 0;
                    uint32_t t2;
                    const uint32_t off = block_excl_scan<NT>(mine, M->scratch, flip, t2);
                    if (r < nrec) r_outoff[r] = (uint16_t)(run + off > 0xFFFFu ? 0xFFFFu : run + off);
                    run += t2;
                }
                chunk_out = run;
            } else if (nrec <= (uint32_t)LAYOUT_FAST_MAXREC && nrec <= (uint32_t)NT) {
                // Sample-major inside the chunk, input order inside a sample.  Every record sets its bit
                // in its sample's 128-bit mask; the lowest record of a sample (the leader) owns the
                // group, members add up the lengths of the peers before them.
                const uint32_t r = tid;
                int sm = -1;
                uint32_t len = 0;
                if (r < nrec) {
                    sm = r_sample[r];
                    len = r_outlen[r];
                    if (sm < 0 || !len) sm = -1, len = 0;
                }
                if (sm >= 0) atomicOr(&masks[4 * sm + (r >> 5)], 1u << (r & 31u));
                __syncthreads();
                uint32_t within = 0, gt = 0, lead = r;
                bool leader = false;
                if (sm >= 0) {
                    const uint4 mv = *(const uint4 *)&masks[4 * sm];
                    const uint32_t wi = r >> 5, bit = 1u << (r & 31u);
                    uint32_t lo[4], hi[4];
                    const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        lo[w] = (uint32_t)w < wi ? mw[w] : ((uint32_t)w == wi ? mw[w] & (bit - 1u) : 0u);
                        hi[w] = (uint32_t)w > wi ? mw[w] : ((uint32_t)w == wi ? mw[w] & ~(bit - 1u) & ~bit : 0u);
                    }
                    leader = (lo[0] | lo[1] | lo[2] | lo[3]) == 0u;
                    if (!leader) {
                        bool first = true;
#pragma unroll
                        for (int w = 0; w < 4; w++) {
                            uint32_t x = lo[w];
                            while (x) {
                                const uint32_t q = 32u * w + (uint32_t)__ffs(x) - 1u;
                                x &= x - 1;
                                if (first) lead = q, first = false;
                                within += r_outlen[q];
                            }
                        }
                    } else {
                        gt = len;
#pragma unroll
                        for (int w = 0; w < 4; w++) {
                            uint32_t x = hi[w];
                            while (x) {
                                const uint32_t q = 32u * w + (uint32_t)__ffs(x) - 1u;
                                x &= x - 1;
                                gt += r_outlen[q];
                            }
                        }
                    }
                }
                // exclusive scan over the leaders (in record order) of group bytes and group count
                uint32_t t2;
                const uint32_t sc = block_excl_scan<NT>(gt | (leader ? 1u << 24 : 0u), M->scratch, flip, t2);
                chunk_out = t2 & 0xFFFFFFu;
                const uint32_t n_groups = t2 >> 24;
                if (leader) {
                    gbase_sm[r] = sc & 0xFFFFFFu;
                    *(uint4 *)&masks[4 * sm] = make_uint4(0, 0, 0, 0);  // every member has read it (barrier in the scan)
                    if (p.out) {
                        Group g;
                        g.sample = (uint16_t)sm;
                        g.len = (uint16_t)gt;
                        p.groups[rec0 + (sc >> 24)] = g;
                    }
                }
                if (tid == 0) M->n_groups = n_groups;
                __syncthreads();
                if (sm >= 0) r_outoff[r] = (uint16_t)(gbase_sm[lead] + within);
            } else {
                // Very dense chunk: thread 0 walks the records (masks[4*s] = group id + 1, masks[4*s+1] =
                // bytes of the group so far, r_aux[g] = group total, r_body[r] = group of record r).
                if (tid == 0) {
                    uint32_t ng = 0;
                    for (uint32_t r = 0; r < nrec; r++) {
                        const int sm = r_sample[r];
                        const uint32_t len = r_outlen[r];
                        if (sm < 0 || !len) continue;
                        uint32_t g = masks[4 * sm];
                        if (!g) {
                            g = ++ng;
                            masks[4 * sm] = g;
                            masks[4 * sm + 1] = 0;
                            r_aux[g - 1] = 0;
                        }
                        r_outoff[r] = (uint16_t)masks[4 * sm + 1];
                        masks[4 * sm + 1] += len;
                        r_aux[g - 1] += len;
                        r_body[r] = (uint16_t)(g - 1);
                    }
                    // group bases in order of first appearance; reset the per-sample words
                    uint32_t run = 0, gi = 0;
                    for (uint32_t r = 0; r < nrec && gi < ng; r++) {
                        const int sm = r_sample[r];
                        if (sm < 0 || !r_outlen[r] || masks[4 * sm] != gi + 1) continue;
                        const uint32_t tot2 = r_aux[gi];
                        if (p.out) {
                            Group g;
                            g.sample = (uint16_t)sm;
                            g.len = (uint16_t)tot2;
                            p.groups[rec0 + gi] = g;
                        }
                        r_aux[gi] = run;
                        run += tot2;
                        gi++;
                    }
                    for (uint32_t r = 0; r < nrec; r++) {
                        const int sm = r_sample[r];
                        if (sm >= 0) masks[4 * sm] = 0, masks[4 * sm + 1] = 0;
                    }
                    M->chunk_out = run;
                    M->n_groups = ng;
                }
                __syncthreads();
                chunk_out = M->chunk_out;
                for (uint32_t r = tid; r < nrec; r += NT) {
                    const int sm = r_sample[r];
                    if (sm >= 0 && r_outlen[r]) r_outoff[r] = (uint16_t)(r_outoff[r] + r_aux[r_body[r]]);
                }
            }

            SK_T(7);  // layout
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
                    ChunkRow row;
                    row.base = base;
                    row.first_group = (uint32_t)rec0;
                    row.n_groups = M->n_groups;
                    p.rows[c] = row;
                    if (chunk_out) atomicAdd(&st->out_bytes, (unsigned long long)chunk_out);
                }
                M->out_base = base;
            }
            __syncthreads();
            SK_T(8);  // reserve (look-back on output bytes / atomic)
            const uint64_t out_base = M->out_base;
            bool writable = p.out != nullptr && chunk_out > 0;
            if (writable && out_base + ((chunk_out + 15u) & ~15u) > p.out_cap) {
                if (tid == 0) report_err(st, rec0, K_OUT_OVERFLOW);
                writable = false;
            }
            if (chunk_out > 0xFFFFu) {  // offsets are 16-bit
                if (tid == 0) report_err(st, rec0, K_OUT_OVERFLOW);
                writable = false;
            }

            bool staged_now = false;
            uint32_t shift = 0;
            if (writable) {
                // ---- P8 assemble.  Fast path: build the chunk's output image in shared memory, aligned
                // like its global destination; 4 lanes per record, each copying whole pieces word-wise.
                shift = (uint32_t)(out_base & 15u);
                const bool staged = shift + chunk_out <= (uint32_t)Cfg::STAGE;
                if (staged) {
                    staged_now = true;
                    uint8_t *sb = stage + shift;
                    const uint32_t q4 = tid & 3;
                    for (uint32_t r0 = 0; r0 < nrec; r0 += NT / 4) {
                        const uint32_t r = r0 + (tid >> 2);
                        // every lane ends up with at most one word-copy job; literals are written directly
                        uint8_t *jd = nullptr;
                        const uint8_t *js = nullptr;
                        uint32_t jl = 0;
                        const uint32_t outlen = r < nrec ? r_outlen[r] : 0;
                        if (outlen) {
                            const uint32_t j = j0 + r * lpr;
                            const uint32_t L0 = LB(j), L1 = LB(j + 1);
                            uint8_t *d0 = sb + r_outoff[r];
                            const uint32_t kk = r_k[r];
                            const uint8_t mode = r_mode[r];
                            uint32_t hlen;  // bytes of the (rewritten) header line incl. '\n'
                            if (OP == OP_TRIM || OP == OP_MASK) {
                                hlen = L1 - L0;
                                if (q4 == 0) { jd = d0; js = win + L0; jl = hlen; }
                            } else {
                                const uint32_t alen = r_alen[r], blen = r_blen[r], taglen = r_taglen[r];
                                hlen = alen + blen + taglen + 1;
                                if (q4 == 0) { jd = d0; js = win + L0; jl = alen; }
                                if (q4 == 3) {  // the (usually empty) piece after the cut, the tag and the newline
                                    uint8_t *d = d0 + alen;
                                    const uint8_t *sB = win + L0 + r_cut1[r];
                                    for (uint32_t i = 0; i < blen; i++) d[i] = sB[i];
                                    d += blen;
                                    if (taglen) {
                                        if (OP == OP_ADDBC) {
                                            d[0] = ' '; d[1] = 'B'; d[2] = 'C'; d[3] = ':';
                                            bcopy(d + 4, p.ext_data[0] + r_aux[r], taglen - 4);
                                        } else {
                                            d[0] = ' '; d[1] = 'U'; d[2] = 'M'; d[3] = 'I'; d[4] = ':';
                                            uint8_t *gu = p.umi + (rec0 + r) * p.sheet.Umax;
                                            if (OP == OP_DEMUX1 && !p.n_index) {
                                                // UMI = observed chars where the sheet barcode has 'U' (:200-203);
                                                // also parked in the side table for mate 2
                                                const uint8_t *ob = win + L0 + r_cut0[r] + 4;
                                                WT m = sh_umask[r_sample[r]];
                                                uint32_t t = 0;
                                                while (m) {
                                                    const uint32_t q = ffsw<WT>(m);
                                                    m &= m - 1;
                                                    const uint8_t ch = ob[q];
                                                    d[5 + t] = ch;
                                                    gu[t] = ch;
                                                    t++;
                                                }
                                            } else {
                                                bcopy(d + 5, gu, taglen - 5);
                                            }
                                        }
                                        d += taglen;
                                    }
                                    d[0] = '\n';
                                }
                            }
                            uint8_t *db = d0 + hlen;  // body destination
                            if (mode == B_VERBATIM) {
                                const uint32_t Lend = LB(j + lpr);
                                const uint32_t blen2 = Lend - L1, half = (blen2 / 2 + 3) & ~3u;
                                const uint32_t h1 = half < blen2 ? half : blen2;
                                if (q4 == 1) { jd = db; js = win + L1; jl = h1; }
                                if (q4 == 2) { jd = db + h1; js = win + L1 + h1; jl = blen2 - h1; }
                            } else if (mode == B_TRIM) {
                                const uint32_t L3 = LB(j + 3);
                                if (q4 == 1) { jd = db; js = win + L1; jl = kk; }
                                if (q4 == 2) { jd = db + kk + 3; js = win + L3; jl = kk; }
                                if (q4 == 3) {
                                    uint8_t *d = db + kk;
                                    d[0] = '\n'; d[1] = '+'; d[2] = '\n';
                                    d[3 + kk] = '\n';
                                }
                            } else if (mode == B_GARBAGE) {
                                if (q4 == 3) {
                                    uint8_t *d = db;
                                    d[0] = 'N'; d[1] = '\n'; d[2] = '+'; d[3] = '\n'; d[4] = '!'; d[5] = '\n';
                                }
                            } else if (mode == B_MASK) {
                                const uint32_t L3 = LB(j + 3);
                                if (q4 == 1) {
                                    const uint32_t minq = p.min_baseq;
                                    for (uint32_t i = 0; i < kk; i++) {
                                        const uint8_t q = (uint8_t)(win[L3 + i] - 33u);  // fasta_mask_by_quality.rs:42
                                        db[i] = q < minq ? (uint8_t)'N' : win[L1 + i];
                                    }
                                }
                                if (q4 == 2) { jd = db + kk + 3; js = win + L3; jl = kk; }
                                if (q4 == 3) {
                                    uint8_t *d = db + kk;
                                    d[0] = '\n'; d[1] = '+'; d[2] = '\n';
                                    d[3 + kk] = '\n';
                                }
                            }
                        }
                        tcopy(jd, js, jl);  // one call site: all lanes copy their piece together
                    }
                    fence_proxy_async();  // staging writes -> visible to the TMA store issued after the barrier
                    SK_T(9);  // assemble
                } else {
                    // Oversized output (pathological growth): one thread per record, bytes straight to global.
                    for (uint32_t r = tid; r < nrec; r += NT) {
                        if (!r_outlen[r]) continue;
                        const uint32_t j = j0 + r * lpr;
                        const uint32_t L0 = LB(j), L1 = LB(j + 1), L3 = LB(j + 3);
                        const uint32_t kk = r_k[r];
                        const uint8_t mode = r_mode[r];
                        uint8_t *d = p.out + out_base + r_outoff[r];
                        if (OP == OP_TRIM || OP == OP_MASK) {
                            bcopy(d, win + L0, L1 - L0);
                            d += L1 - L0;
                        } else {
                            const uint32_t alen = r_alen[r], blen = r_blen[r], taglen = r_taglen[r];
                            bcopy(d, win + L0, alen);
                            d += alen;
                            bcopy(d, win + L0 + r_cut1[r], blen);
                            d += blen;
                            if (taglen) {
                                if (OP == OP_ADDBC) {
                                    bcopy(d, (const uint8_t *)" BC:", 4);
                                    bcopy(d + 4, p.ext_data[0] + r_aux[r], taglen - 4);
                                } else {
                                    bcopy(d, (const uint8_t *)" UMI:", 5);
                                    uint8_t *gu = p.umi + (rec0 + r) * p.sheet.Umax;
                                    if (OP == OP_DEMUX1 && !p.n_index) {
                                        const uint8_t *ob = win + L0 + r_cut0[r] + 4;
                                        WT m = sh_umask[r_sample[r]];
                                        uint32_t t = 0;
                                        while (m) {
                                            const uint32_t q = ffsw<WT>(m);
                                            m &= m - 1;
                                            gu[t++] = ob[q];
                                        }
                                    }
                                    bcopy(d + 5, gu, taglen - 5);
                                }
                                d += taglen;
                            }
                            *d++ = '\n';
                        }
                        if (mode == B_VERBATIM) {
                            bcopy(d, win + L1, LB(j + lpr) - L1);
                        } else if (mode == B_TRIM) {
                            bcopy(d, win + L1, kk);
                            bcopy(d + kk, (const uint8_t *)"\n+\n", 3);
                            bcopy(d + kk + 3, win + L3, kk);
                            d[2 * kk + 3] = '\n';
                        } else if (mode == B_GARBAGE) {
                            bcopy(d, (const uint8_t *)"N\n+\n!\n", 6);
                        } else if (mode == B_MASK) {
                            for (uint32_t i = 0; i < kk; i++) {
                                const uint8_t q = (uint8_t)(win[L3 + i] - 33u);
                                d[i] = q < p.min_baseq ? (uint8_t)'N' : win[L1 + i];
                            }
                            bcopy(d + kk, (const uint8_t *)"\n+\n", 3);
                            bcopy(d + kk + 3, win + L3, kk);
                            d[2 * kk + 3] = '\n';
                        }
                    }
                }
            }
            // UMI side table for mate 2 of assigned records whose output was not assembled here
            if (OP == OP_DEMUX1 && !p.n_index && !writable && p.sheet.Umax) {
                for (uint32_t r = tid; r < nrec; r += NT) {
                    const int sm = r_sample[r];
                    if (sm < 0 || (r_flags[r] & RF_DEAD)) continue;
                    const uint32_t L0 = LB(j0 + r * lpr);
                    const uint8_t *ob = win + L0 + r_cut0[r] + 4;
                    uint8_t *gu = p.umi + (rec0 + r) * p.sheet.Umax;
                    WT m = sh_umask[sm];
                    uint32_t t = 0;
                    while (m) {
                        const uint32_t q = ffsw<WT>(m);
                        m &= m - 1;
                        gu[t++] = ob[q];
                    }
                }
            }

            // ---- P10 per-record side tables, next ticket
            if (OP == OP_DEMUX1)
                for (uint32_t r = tid; r < nrec; r += NT) p.assign[rec0 + r] = r_sample[r];
            if (tid == 0) {
                // Take the next ticket only now: a ticket held while another chunk is still being processed
                // stalls the look-back of every later chunk.
                M->chunk = atomicAdd(&st->ticket, 1u);
                M->n_slow = 0;
            }
            __syncthreads();  // window, staging and record arrays are reused by the next chunk

            // ---- P9 store: one TMA bulk copy of the staged image (16-byte units); the ragged first and
            // last bytes of an ordered chunk belong to 16-byte units shared with its neighbours.
            if (staged_now) {
                const uint32_t span = shift + chunk_out;  // image occupies stage[shift, span)
                uint8_t *g16 = p.out + (out_base - shift);
                if (ORDERED) {
                    const uint32_t a = shift ? 16u : 0u;   // first whole unit
                    const uint32_t b2 = span & ~15u;       // end of the last whole unit
                    if (b2 > a) {
                        if (tid == 0) {
                            bulk_s2g(g16 + a, stage + a, b2 - a);
                            bulk_commit();
                            store_pending = true;
                        }
                        if (tid < 32) {
                            const uint32_t t = (uint32_t)tid;
                            if (t < 16u) {
                                if (t >= shift && t < a && t < span) g16[t] = stage[t];
                            } else {
                                const uint32_t o = b2 + (t - 16u);
                                if (o < span) g16[o] = stage[o];
                            }
                        }
                        // the byte stores above read the staging image with generic loads; they finish
                        // before the next chunk's assembly because of the barriers in between
                    } else {
                        for (uint32_t o = shift + tid; o < span; o += NT) g16[o] = stage[o];
                    }
                } else if (tid == 0) {
                    bulk_s2g(g16, stage, (span + 15u) & ~15u);  // demux chunks own whole 16-byte units
                    bulk_commit();
                    store_pending = true;
                }
            }
            SK_T(10);  // store + tables
        } else {
            if (tid == 0) {
                M->chunk = atomicAdd(&st->ticket, 1u);
                M->n_slow = 0;
            }
            __syncthreads();
        }
    }
    if (tid == 0 && store_pending) bulk_wait_read0();  // shared memory must outlive the last TMA store
#ifdef SK_PHASE_TIMING
    if (tid == 0)
        for (int i = 0; i < 16; i++)
            if (ph[i]) atomicAdd(&st->phase_cycles[i], ph[i]);
#endif

    // ---- flush per-CTA counters (fasta_demultiplex.rs:108-109,169,177-178)
    if (OP == OP_DEMUX1) {
        __syncthreads();
        for (uint32_t s = tid; s < S; s += NT)
            if (ccount[s]) atomicAdd(&p.counts[s], (unsigned long long)ccount[s]);
        if (tid == 0) {
            if (M->cta_total) atomicAdd(&p.counts[S], (unsigned long long)M->cta_total);
            if (M->cta_ident) atomicAdd(&p.counts[S + 1], (unsigned long long)M->cta_ident);
            if (M->cta_slow) atomicAdd(&st->n_slow, M->cta_slow);
        }
    }
#undef LB
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
int chunk_kernel_smem_bytes(int cfg, uint32_t S, uint32_t wide, uint32_t n_classes, uint32_t nwp) {
    return cfg == CfgB::ID ? (int)smem_layout<CfgB>(S, wide, n_classes, nwp).total
                           : (int)smem_layout<CfgA>(S, wide, n_classes, nwp).total;
}

template <class Cfg, int OP, typename WT>
static int launch_one(const KParams &p_in, int sm_count, cudaStream_t stream, const char **err) {
    auto kfn = sk_chunk_kernel<Cfg, OP, WT>;
    KParams p = p_in;
    {
        const bool demux = (OP == OP_DEMUX1 || OP == OP_DEMUX2);
        const bool d1 = OP == OP_DEMUX1;
        p.sl = smem_layout<Cfg>(demux ? p.sheet.S : 0u, sizeof(WT) == 8 ? 1u : 0u, d1 ? p.sheet.hidx.n_classes : 0u,
                                d1 ? p.sheet.hidx.nwp : 0u);
    }
    const int smem = (int)p.sl.total;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Cfg::NT, smem);
    if (e != cudaSuccess || per_sm < 1) {
        *err = e != cudaSuccess ? cudaGetErrorString(e) : "kernel does not fit on an SM";
        return -1;
    }
    long long grid = (long long)sm_count * per_sm;  // persistent CTAs, chunks handed out by ticket
    if (grid > (long long)p.n_chunks) grid = p.n_chunks;
    if (grid < 1) return 0;
    kfn<<<(unsigned)grid, Cfg::NT, smem, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return -1;
    }
    return 1;
}

template <class Cfg>
static int launch_cfg(int op, const KParams &p, int sm_count, cudaStream_t stream, const char **err) {
    const bool wide = p.sheet.wide != 0;
    switch (op) {
        case OP_SCAN: return launch_one<Cfg, OP_SCAN, uint32_t>(p, sm_count, stream, err);
        case OP_TRIM: return launch_one<Cfg, OP_TRIM, uint32_t>(p, sm_count, stream, err);
        case OP_MASK: return launch_one<Cfg, OP_MASK, uint32_t>(p, sm_count, stream, err);
        case OP_ADDBC: return launch_one<Cfg, OP_ADDBC, uint32_t>(p, sm_count, stream, err);
        case OP_DEMUX1:
            return wide ? launch_one<Cfg, OP_DEMUX1, uint64_t>(p, sm_count, stream, err)
                        : launch_one<Cfg, OP_DEMUX1, uint32_t>(p, sm_count, stream, err);
        case OP_DEMUX2:
            return wide ? launch_one<Cfg, OP_DEMUX2, uint64_t>(p, sm_count, stream, err)
                        : launch_one<Cfg, OP_DEMUX2, uint32_t>(p, sm_count, stream, err);
    }
    *err = "unknown operator";
    return -1;
}

int launch_chunk_kernel(int cfg, int op, const KParams &p, int sm_count, void *stream_, const char **err) {
    cudaStream_t stream = (cudaStream_t)stream_;
    return cfg == CfgB::ID ? launch_cfg<CfgB>(op, p, sm_count, stream, err)
                           : launch_cfg<CfgA>(op, p, sm_count, stream, err);
}

}  // namespace sk
