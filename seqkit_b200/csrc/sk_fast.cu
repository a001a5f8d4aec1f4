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

}  // namespace sk
#include "sk_lean.cuh"
namespace sk {


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
