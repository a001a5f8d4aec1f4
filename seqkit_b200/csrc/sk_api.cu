// sk_api.cu -- host side of the C ABI in include/seqkit_b200.h: contexts, slots, sample-sheet
// packing, operator launch sequences, result collection.  No CPU implementation of any operator
// lives here: every operator enqueues the chunk-engine kernels of sk_kernels.cu.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/seqkit_b200.h"
#include "sk_internal.h"

using namespace sk;

static_assert(sizeof(sk_event) == sizeof(Event), "sk_event layout");

namespace sk {
int launch_synth(void *ctx_stream, uint8_t *dst, uint64_t cap, const sk_synth_spec &spec, const uint8_t *sheet_raw,
                 uint32_t S, uint32_t L, uint64_t *offsets_tmp, uint64_t *n_out, const char **err);
}

static std::string g_create_error;

struct Slot {
    cudaStream_t stream = nullptr;
    uint8_t *in[SK_N_INPUTS] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t in_cap[SK_N_INPUTS] = {0, 0, 0, 0};
    uint64_t in_len[SK_N_INPUTS] = {0, 0, 0, 0};
    uint8_t *out[2] = {nullptr, nullptr};
    uint64_t out_cap = 0;
    uint64_t *tile_lines[SK_N_INPUTS] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t *tile_out = nullptr;
    DevStats *stats = nullptr;    // [SK_N_INPUTS] device
    DevStats *stats_h = nullptr;  // pinned mirror
    int16_t *assign = nullptr;
    uint8_t *umi = nullptr;
    uint64_t umi_cap = 0;
    uint16_t *lens[2] = {nullptr, nullptr};
    uint64_t *chunk_base[2] = {nullptr, nullptr};
    unsigned long long *counts = nullptr;
    Event *events = nullptr;
    RecRef *scan_tab[2] = {nullptr, nullptr};
    uint64_t *synth_tmp = nullptr;
    // description of the last operator, for sk_wait
    int last_op = -1;
    bool paired = false;
    uint32_t n_chunks[SK_N_INPUTS] = {0, 0, 0, 0};
    uint32_t launches = 0;
    bool pass_ran[SK_N_INPUTS] = {false, false, false, false};
    cudaEvent_t ev[SK_N_INPUTS][2] = {};
};

struct sk_ctx {
    int device = 0;
    int sm_count = 0;
    sk_limits lim{};
    uint32_t max_chunks = 0;
    std::vector<Slot> slots;
    std::string err;
    bool profiling = false;
    // sample sheet
    bool have_sheet = false;
    uint32_t S = 0, L = 0, Umax = 0, wide = 0;
    std::vector<uint8_t> sheet_raw;
    uint32_t *d_planes = nullptr, *d_umask = nullptr;
    uint8_t *d_lut = nullptr, *d_sheet_raw = nullptr;
    // exact-match index (FastIdx)
    uint32_t f_classes = 0, f_nw = 0, f_tsize = 0;
    uint32_t *d_fcls = nullptr, *d_skeys = nullptr;
    unsigned long long *d_ftab = nullptr;
    int cfg = 0;  // chunk-engine geometry: 0 = CfgBig (32 KiB chunks), 1 = CfgSmall (16 KiB chunks)
};

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return SK_E_CUDA;                                                            \
        }                                                                                \
    } while (0)

static uint32_t chunks_of(const sk_ctx *ctx, uint64_t n) {
    const uint64_t ch = (uint64_t)cfg_chunk_bytes(ctx->cfg);
    return (uint32_t)((n + ch - 1) / ch);
}

extern "C" int sk_abi_version(void) { return SK_ABI_VERSION; }

extern "C" const char *sk_last_error(const sk_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static void free_slot(Slot &s) {
    for (int i = 0; i < SK_N_INPUTS; i++) {
        cudaFree(s.in[i]);
        cudaFree(s.tile_lines[i]);
    }
    for (int i = 0; i < 2; i++) {
        cudaFree(s.out[i]);
        cudaFree(s.lens[i]);
        cudaFree(s.chunk_base[i]);
        cudaFree(s.scan_tab[i]);
    }
    cudaFree(s.tile_out);
    cudaFree(s.stats);
    cudaFreeHost(s.stats_h);
    cudaFree(s.assign);
    cudaFree(s.umi);
    cudaFree(s.counts);
    cudaFree(s.events);
    cudaFree(s.synth_tmp);
    for (int i = 0; i < SK_N_INPUTS; i++)
        for (int k = 0; k < 2; k++)
            if (s.ev[i][k]) cudaEventDestroy(s.ev[i][k]);
    if (s.stream) cudaStreamDestroy(s.stream);
}

extern "C" void sk_ctx_destroy(sk_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto &s : ctx->slots) free_slot(s);
    cudaFree(ctx->d_planes);
    cudaFree(ctx->d_umask);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_sheet_raw);
    cudaFree(ctx->d_fcls);
    cudaFree(ctx->d_skeys);
    cudaFree(ctx->d_ftab);
    delete ctx;
}

extern "C" int sk_ctx_create(int device, const sk_limits *lim, sk_ctx **out) {
    if (!lim || !out) return SK_E_INVALID;
    *out = nullptr;
    if (lim->max_stream_bytes == 0 || lim->max_stream_bytes >= (1ull << 32) - (1u << 20) || lim->n_slots == 0 ||
        lim->n_slots > 16 || lim->max_records == 0 || lim->max_records >= (1ull << 32)) {
        g_create_error = "sk_ctx_create: limits out of range (streams must be < 4 GiB, 1..16 slots)";
        return SK_E_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        g_create_error = std::string("sk_ctx_create: no usable CUDA device (") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
                         "); seqkit_b200 has no CPU fallback";
        return SK_E_CUDA;
    }
    sk_ctx *ctx = new sk_ctx();
    ctx->device = device;
    ctx->lim = *lim;
    ctx->cfg = lim->reserved == 1 ? 1 : lim->reserved == 2 ? 0 : 1;  // reserved: 0 default, 1 small, 2 big chunks
    if (const char *e = getenv("SK_CFG")) ctx->cfg = atoi(e) ? 1 : 0;
    auto fail = [&](int code) {
        g_create_error = ctx->err;
        sk_ctx_destroy(ctx);
        return code;
    };
#define CKC(call)                                                              \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);     \
            return fail(e_ == cudaErrorMemoryAllocation ? SK_E_NOMEM : SK_E_CUDA); \
        }                                                                      \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        ctx->err = "sk_ctx_create: device is not sm_100 (Blackwell) class; this library ships sm_100a code only";
        return fail(SK_E_CUDA);
    }
    ctx->sm_count = prop.multiProcessorCount;
    const uint64_t B = (lim->max_stream_bytes + 15) & ~15ull;
    const uint64_t R = lim->max_records;
    ctx->max_chunks = chunks_of(ctx, B) + 1;
    const uint64_t out_cap = B + R * 72 + (uint64_t)ctx->max_chunks * 16 + 4096;
    const uint32_t Smax = lim->max_samples;
    ctx->slots.resize(lim->n_slots);
    for (auto &s : ctx->slots) {
        CKC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        for (int i = 0; i < SK_N_INPUTS; i++)
            for (int k = 0; k < 2; k++) CKC(cudaEventCreate(&s.ev[i][k]));
        const int nin = lim->aux_streams ? SK_N_INPUTS : 2;
        for (int i = 0; i < nin; i++) {
            CKC(cudaMalloc(&s.in[i], B + 64));
            s.in_cap[i] = B;
            CKC(cudaMalloc(&s.tile_lines[i], (uint64_t)ctx->max_chunks * 8));
        }
        for (int i = 0; i < 2; i++) CKC(cudaMalloc(&s.out[i], out_cap));
        s.out_cap = out_cap;
        CKC(cudaMalloc(&s.tile_out, (uint64_t)ctx->max_chunks * 8));
        CKC(cudaMalloc(&s.stats, sizeof(DevStats) * SK_N_INPUTS));
        CKC(cudaMallocHost(&s.stats_h, sizeof(DevStats) * SK_N_INPUTS));
        memset(s.stats_h, 0, sizeof(DevStats) * SK_N_INPUTS);
        CKC(cudaMalloc(&s.synth_tmp, (R + 1) * 8));
        if (Smax) {
            CKC(cudaMalloc(&s.assign, R * 2));
            for (int i = 0; i < 2; i++) {
                CKC(cudaMalloc(&s.lens[i], (uint64_t)ctx->max_chunks * Smax * 2));
                CKC(cudaMalloc(&s.chunk_base[i], (uint64_t)ctx->max_chunks * 8));
            }
            CKC(cudaMalloc(&s.counts, (uint64_t)(Smax + 2) * 8));
            CKC(cudaMalloc(&s.events, R * sizeof(Event)));
        }
        if (lim->aux_streams)
            for (int i = 0; i < 2; i++) CKC(cudaMalloc(&s.scan_tab[i], R * sizeof(RecRef)));
    }
#undef CKC
    *out = ctx;
    return SK_OK;
}

static Slot *get_slot(sk_ctx *ctx, uint32_t slot) {
    if (!ctx || slot >= ctx->slots.size()) return nullptr;
    cudaSetDevice(ctx->device);
    return &ctx->slots[slot];
}

extern "C" void *sk_slot_stream(sk_ctx *ctx, uint32_t slot) {
    Slot *s = get_slot(ctx, slot);
    return s ? (void *)s->stream : nullptr;
}
extern "C" uint32_t sk_max_chunks(sk_ctx *ctx) { return ctx ? ctx->max_chunks : 0; }
extern "C" int sk_debug_phase_cycles(sk_ctx *ctx, uint32_t slot, uint32_t which, uint64_t out[16]) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= SK_N_INPUTS || !out) return SK_E_INVALID;
    CK(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < 16; i++) out[i] = s->stats_h[which].phase_cycles[i];
    return SK_OK;
}
extern "C" int sk_set_profiling(sk_ctx *ctx, int on) {
    if (!ctx) return SK_E_INVALID;
    ctx->profiling = on != 0;
    return SK_OK;
}
extern "C" void *sk_pinned_alloc(sk_ctx *ctx, uint64_t bytes) {
    if (!ctx || !bytes) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        ctx->err = "cudaMallocHost failed";
        return nullptr;
    }
    return p;
}
extern "C" void sk_pinned_free(sk_ctx *ctx, void *p) {
    if (ctx && p) cudaFreeHost(p);
}
extern "C" uint64_t sk_out_capacity(sk_ctx *ctx) { return (ctx && !ctx->slots.empty()) ? ctx->slots[0].out_cap : 0; }
extern "C" void *sk_slot_in(sk_ctx *ctx, uint32_t slot, uint32_t which) {
    Slot *s = get_slot(ctx, slot);
    return (s && which < SK_N_INPUTS) ? s->in[which] : nullptr;
}
extern "C" uint64_t sk_slot_in_capacity(sk_ctx *ctx, uint32_t slot, uint32_t which) {
    Slot *s = get_slot(ctx, slot);
    return (s && which < SK_N_INPUTS && s->in[which]) ? s->in_cap[which] : 0;
}
extern "C" int sk_set_input_len(sk_ctx *ctx, uint32_t slot, uint32_t which, uint64_t n) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= SK_N_INPUTS) return SK_E_INVALID;
    if (n && (!s->in[which] || n > s->in_cap[which])) {
        ctx->err = "input larger than the slot capacity";
        return SK_E_TOO_LARGE;
    }
    s->in_len[which] = n;
    return SK_OK;
}
extern "C" int sk_upload(sk_ctx *ctx, uint32_t slot, uint32_t which, const void *host, uint64_t n) {
    int rc = sk_set_input_len(ctx, slot, which, n);
    if (rc != SK_OK) return rc;
    Slot *s = get_slot(ctx, slot);
    if (n) CK(cudaMemcpyAsync(s->in[which], host, n, cudaMemcpyHostToDevice, s->stream));
    return SK_OK;
}
extern "C" int sk_download_in(sk_ctx *ctx, uint32_t slot, uint32_t which, void *host, uint64_t n) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= SK_N_INPUTS || n > s->in_len[which]) return SK_E_INVALID;
    if (n) CK(cudaMemcpyAsync(host, s->in[which], n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SK_OK;
}

// ------------------------------------------------------------------------------------------------
// sample sheet -> bit planes (replaces the byte loop of barcode_diff, fasta_demultiplex.rs:269-277)
// ------------------------------------------------------------------------------------------------
extern "C" int sk_set_sheet(sk_ctx *ctx, const uint8_t *barcodes, uint32_t S, uint32_t L) {
    if (!ctx || (!barcodes && S && L)) return SK_E_INVALID;
    cudaSetDevice(ctx->device);
    if (S > ctx->lim.max_samples || S > 32767) {
        ctx->err = "sample sheet larger than sk_limits.max_samples";
        return SK_E_TOO_LARGE;
    }
    if (L > 64) {
        ctx->err = "barcodes longer than 64 characters are not supported by the bit-plane matcher";
        return SK_E_UNSUPPORTED;
    }
    const uint32_t wide = L > 32 ? 1u : 0u;
    // literal alphabet of the sheet: every byte that is not a wildcard ('N'/'U', :273)
    uint8_t lut[256];
    memset(lut, 0, sizeof lut);
    uint32_t ncode = 0;
    for (uint64_t i = 0; i < (uint64_t)S * L; i++) {
        const uint8_t b = barcodes[i];
        if (b == 'N' || b == 'U' || lut[b]) continue;
        if (ncode == 7) {
            ctx->err = "sample sheet uses more than 7 distinct literal characters";
            return SK_E_UNSUPPORTED;
        }
        lut[b] = (uint8_t)(++ncode);
    }
    for (const char *c = "ACGTNacgtn+"; *c; c++) lut[(uint8_t)*c] |= 8u;  // regex class of fasta_demultiplex.rs:38
    const uint32_t wpe = wide ? 2u : 1u;  // u32 words per plane element
    std::vector<uint32_t> planes((size_t)S * 4 * wpe, 0u), umask((size_t)S * wpe, 0u);
    uint32_t Umax = 0;
    for (uint32_t s = 0; s < S; s++) {
        uint64_t pl[4] = {0, 0, 0, 0}, um = 0;
        for (uint32_t q = 0; q < L; q++) {
            const uint8_t b = barcodes[(uint64_t)s * L + q];
            if (b == 'U') um |= 1ull << q;
            if (b == 'N' || b == 'U') continue;
            const uint32_t code = lut[b] & 7u;
            for (int k = 0; k < 3; k++)
                if ((code >> k) & 1u) pl[k] |= 1ull << q;
            pl[3] |= 1ull << q;  // care
        }
        for (int k = 0; k < 4; k++) {
            planes[((size_t)s * 4 + k) * wpe] = (uint32_t)pl[k];
            if (wide) planes[((size_t)s * 4 + k) * wpe + 1] = (uint32_t)(pl[k] >> 32);
        }
        umask[(size_t)s * wpe] = (uint32_t)um;
        if (wide) umask[(size_t)s * wpe + 1] = (uint32_t)(um >> 32);
        Umax = std::max<uint32_t>(Umax, (uint32_t)__builtin_popcountll(um));
    }

    // ---- exact-match index: classes of identical care mask, one open-addressing table per class
    const uint32_t nw = (L + 3) / 4;
    std::vector<std::vector<uint8_t>> cls_care;
    std::vector<uint32_t> cls_of(S, 0);
    for (uint32_t s = 0; s < S; s++) {
        std::vector<uint8_t> care(L);
        for (uint32_t q = 0; q < L; q++) {
            const uint8_t b = barcodes[(uint64_t)s * L + q];
            care[q] = (b != 'N' && b != 'U') ? 0xFF : 0x00;
        }
        uint32_t c = 0;
        for (; c < cls_care.size(); c++)
            if (cls_care[c] == care) break;
        if (c == cls_care.size()) cls_care.push_back(care);
        cls_of[s] = c;
    }
    uint32_t f_classes = (S && L && cls_care.size() <= 2) ? (uint32_t)cls_care.size() : 0;
    uint32_t tsize = 16;
    while (3 * tsize < 4 * S) tsize <<= 1;  // load factor <= 0.75
    if (f_classes && chunk_kernel_smem_bytes(ctx->cfg, S, wide, f_classes, tsize) > 227 * 1024) f_classes = 0;
    if (chunk_kernel_smem_bytes(ctx->cfg, S, wide, f_classes, tsize) > 227 * 1024) {
        ctx->err = "sample sheet does not fit in shared memory";
        return SK_E_UNSUPPORTED;
    }
    std::vector<uint32_t> skeys((size_t)S * std::max(nw, 1u), 0u), fcls((size_t)std::max(f_classes, 1u) * FAST_CLS_WORDS, 0u);
    std::vector<unsigned long long> ftab((size_t)std::max(f_classes, 1u) * tsize, 0xFFFFull << 32);
    for (uint32_t s = 0; s < S; s++)
        for (uint32_t q = 0; q < L; q++) {
            const uint8_t b = barcodes[(uint64_t)s * L + q] & cls_care[cls_of[s]][q];
            skeys[(size_t)s * nw + q / 4] |= (uint32_t)b << (8 * (q % 4));
        }
    uint64_t rng = 0x9E3779B97F4A7C15ull ^ ((uint64_t)S << 32) ^ L;
    auto next = [&]() {
        rng ^= rng << 13;
        rng ^= rng >> 7;
        rng ^= rng << 17;
        return (uint32_t)(rng >> 16);
    };
    for (uint32_t c = 0; c < f_classes; c++) {
        uint32_t *cw = &fcls[(size_t)c * FAST_CLS_WORDS];
        for (uint32_t q = 0; q < L; q++) cw[q / 4] |= (uint32_t)cls_care[c][q] << (8 * (q % 4));
        for (uint32_t w = 0; w < (uint32_t)FAST_NWMAX; w++) {
            cw[FAST_NWMAX + w] = next() | 1u;
            cw[2 * FAST_NWMAX + w] = next() | 1u;
        }
        unsigned long long *tab = &ftab[(size_t)c * tsize];
        for (uint32_t s = 0; s < S; s++) {
            if (cls_of[s] != c) continue;
            const uint32_t *key = &skeys[(size_t)s * nw];
            uint32_t h1 = 0, h2 = 0;
            for (uint32_t w = 0; w < nw; w++) {
                h1 += key[w] * cw[FAST_NWMAX + w];
                h2 += key[w] * cw[2 * FAST_NWMAX + w];
            }
            h1 ^= h1 >> 15;
            uint32_t slot = h1 & (tsize - 1);
            for (;;) {
                const unsigned long long e = tab[slot];
                const uint32_t f = (uint32_t)(e >> 32) & 0xFFFFu;
                if (f == 0xFFFFu) {
                    tab[slot] = (unsigned long long)h2 | ((unsigned long long)s << 32) | ((unsigned long long)s << 48);
                    break;
                }
                if (memcmp(&skeys[(size_t)f * nw], key, (size_t)nw * 4) == 0) {  // duplicate barcode: widen [first,last]
                    tab[slot] = (e & 0x0000FFFFFFFFFFFFull) | ((unsigned long long)s << 48);
                    break;
                }
                slot = (slot + 1) & (tsize - 1);
            }
        }
    }

    cudaFree(ctx->d_planes);
    cudaFree(ctx->d_umask);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_sheet_raw);
    cudaFree(ctx->d_fcls);
    cudaFree(ctx->d_skeys);
    cudaFree(ctx->d_ftab);
    ctx->d_planes = ctx->d_umask = ctx->d_fcls = ctx->d_skeys = nullptr;
    ctx->d_lut = ctx->d_sheet_raw = nullptr;
    ctx->d_ftab = nullptr;
    CK(cudaMalloc(&ctx->d_planes, std::max<size_t>(planes.size() * 4, 16)));
    CK(cudaMalloc(&ctx->d_umask, std::max<size_t>(umask.size() * 4, 16)));
    CK(cudaMalloc(&ctx->d_lut, 256));
    CK(cudaMalloc(&ctx->d_sheet_raw, std::max<size_t>((size_t)S * L, 16)));
    CK(cudaMalloc(&ctx->d_fcls, fcls.size() * 4));
    CK(cudaMalloc(&ctx->d_skeys, std::max<size_t>(skeys.size() * 4, 16)));
    CK(cudaMalloc(&ctx->d_ftab, ftab.size() * 8));
    if (S) {
        CK(cudaMemcpy(ctx->d_planes, planes.data(), planes.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->d_umask, umask.data(), umask.size() * 4, cudaMemcpyHostToDevice));
        if (L) CK(cudaMemcpy(ctx->d_sheet_raw, barcodes, (size_t)S * L, cudaMemcpyHostToDevice));
        if (L) CK(cudaMemcpy(ctx->d_skeys, skeys.data(), skeys.size() * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(ctx->d_lut, lut, 256, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_fcls, fcls.data(), fcls.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_ftab, ftab.data(), ftab.size() * 8, cudaMemcpyHostToDevice));
    ctx->sheet_raw.assign(barcodes, barcodes + (size_t)S * L);
    ctx->S = S;
    ctx->L = L;
    ctx->Umax = Umax;
    ctx->wide = wide;
    ctx->f_classes = f_classes;
    ctx->f_nw = nw;
    ctx->f_tsize = tsize;
    ctx->have_sheet = true;
    // the UMI side table depends on the sheet
    for (auto &s : ctx->slots) {
        const uint64_t need = ctx->lim.max_records * (uint64_t)Umax;
        if (need > s.umi_cap) {
            cudaFree(s.umi);
            s.umi = nullptr;
            s.umi_cap = 0;
            CK(cudaMalloc(&s.umi, need));
            s.umi_cap = need;
        }
    }
    return SK_OK;
}

// ------------------------------------------------------------------------------------------------
// operator launch sequences
// ------------------------------------------------------------------------------------------------
static int begin_op(sk_ctx *ctx, Slot *s, int op) {
    s->last_op = op;
    s->launches = 0;
    for (int i = 0; i < SK_N_INPUTS; i++) {
        s->pass_ran[i] = false;
        s->n_chunks[i] = 0;
    }
    CK(cudaMemsetAsync(s->stats, 0, sizeof(DevStats) * SK_N_INPUTS, s->stream));
    return SK_OK;
}

static void base_params(sk_ctx *ctx, Slot *s, int which, KParams &p) {
    memset(&p, 0, sizeof p);
    p.in = s->in[which];
    p.n = s->in_len[which];
    p.n_chunks = chunks_of(ctx, p.n);
    p.lpr = 4;
    p.rec_limit = ~0ull;
    p.final_batch = 1;
    p.fused_trim = -1;
    p.tile_lines = s->tile_lines[which];
    p.tile_out = s->tile_out;
    p.stats = s->stats + which;
    s->n_chunks[which] = p.n_chunks;
}

static int run_pass(sk_ctx *ctx, Slot *s, int which, int op, const KParams &p, bool ordered_out) {
    if (p.n_chunks == 0) return SK_OK;
    CK(cudaMemsetAsync(p.tile_lines, 0, (uint64_t)p.n_chunks * 8, s->stream));
    if (ordered_out) CK(cudaMemsetAsync(p.tile_out, 0, (uint64_t)p.n_chunks * 8, s->stream));
    const char *err = nullptr;
    if (ctx->profiling) CK(cudaEventRecord(s->ev[which][0], s->stream));
    int rc = launch_chunk_kernel(ctx->cfg, op, p, ctx->sm_count, s->stream, &err);
    if (rc < 0) {
        ctx->err = std::string("kernel launch failed: ") + (err ? err : "?");
        return SK_E_CUDA;
    }
    if (ctx->profiling) CK(cudaEventRecord(s->ev[which][1], s->stream));
    s->launches += (uint32_t)rc;
    s->pass_ran[which] = true;
    return SK_OK;
}

static int end_op(sk_ctx *ctx, Slot *s) {
    CK(cudaMemcpyAsync(s->stats_h, s->stats, sizeof(DevStats) * SK_N_INPUTS, cudaMemcpyDeviceToHost, s->stream));
    return SK_OK;
}

static int stream_op(sk_ctx *ctx, uint32_t slot, int op, uint32_t min_baseq, uint64_t rec_limit) {
    Slot *s = get_slot(ctx, slot);
    if (!s || min_baseq > 255) return SK_E_INVALID;
    int rc = begin_op(ctx, s, op);
    if (rc) return rc;
    KParams p;
    base_params(ctx, s, SK_IN_R1, p);
    p.min_baseq = min_baseq;
    p.rec_limit = rec_limit ? rec_limit : ~0ull;
    p.out = s->out[0];
    p.out_cap = s->out_cap;
    rc = run_pass(ctx, s, SK_IN_R1, op, p, true);
    if (rc) return rc;
    return end_op(ctx, s);
}

extern "C" int sk_trim_by_quality(sk_ctx *ctx, uint32_t slot, uint32_t min_baseq, uint64_t rec_limit) {
    return stream_op(ctx, slot, OP_TRIM, min_baseq, rec_limit);
}
extern "C" int sk_mask_by_quality(sk_ctx *ctx, uint32_t slot, uint32_t min_baseq, uint64_t rec_limit) {
    return stream_op(ctx, slot, OP_MASK, min_baseq, rec_limit);
}

// First byte of an input stream decides the framing ('@': 4 lines, '>': 2 lines), as the per-record
// test of fasta_add_barcode.rs:21-27,35-43 does for a uniform file.
static int peek_first_byte(sk_ctx *ctx, Slot *s, int which, int *c) {
    *c = -1;
    if (!s->in_len[which]) return SK_OK;
    uint8_t b = 0;
    CK(cudaMemcpyAsync(&b, s->in[which], 1, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *c = b;
    return SK_OK;
}

extern "C" int sk_add_barcode(sk_ctx *ctx, uint32_t slot, uint64_t rec_limit) {
    Slot *s = get_slot(ctx, slot);
    if (!s) return SK_E_INVALID;
    if (!s->in[SK_IN_AUX1]) {
        ctx->err = "sk_add_barcode needs sk_limits.aux_streams = 1";
        return SK_E_INVALID;
    }
    int c_reads, c_bc;
    int rc = peek_first_byte(ctx, s, SK_IN_R1, &c_reads);
    if (rc) return rc;
    rc = peek_first_byte(ctx, s, SK_IN_AUX1, &c_bc);
    if (rc) return rc;
    rc = begin_op(ctx, s, OP_ADDBC);
    if (rc) return rc;
    // pass 1: where is the sequence line of every barcode record?
    KParams q;
    base_params(ctx, s, SK_IN_AUX1, q);
    const bool bc_fastx = (c_bc == '@' || c_bc == '>');
    q.lpr = c_bc == '>' ? 2 : 4;
    q.head_char = bc_fastx ? (uint32_t)c_bc : 0;
    q.scan_out = s->scan_tab[0];
    q.scan_cap = ctx->lim.max_records;
    // A barcode file whose first line is neither '@' nor '>' never yields a barcode (:20-27): every
    // read gets an empty one.  (A later '@' line in such a file is reported as mixed format.)
    if (!bc_fastx) q.rec_limit = 0, q.head_char = 0xFFFF;
    rc = run_pass(ctx, s, SK_IN_AUX1, OP_SCAN, q, false);
    if (rc) return rc;
    // pass 2: the reads
    KParams p;
    base_params(ctx, s, SK_IN_R1, p);
    p.lpr = c_reads == '>' ? 2 : 4;
    p.head_char = c_reads == '>' ? '>' : '@';
    p.rec_limit = rec_limit ? rec_limit : ~0ull;
    p.out = s->out[0];
    p.out_cap = s->out_cap;
    p.ext_tab[0] = s->scan_tab[0];
    p.ext_data[0] = s->in[SK_IN_AUX1];
    p.ext_stats[0] = s->stats + SK_IN_AUX1;
    rc = run_pass(ctx, s, SK_IN_R1, OP_ADDBC, p, true);
    if (rc) return rc;
    return end_op(ctx, s);
}

extern "C" int sk_demultiplex(sk_ctx *ctx, uint32_t slot, const sk_demux_opts *o) {
    Slot *s = get_slot(ctx, slot);
    if (!s || !o) return SK_E_INVALID;
    if (!ctx->have_sheet || !s->assign) {
        ctx->err = "sk_demultiplex: call sk_set_sheet first (and create the context with max_samples > 0)";
        return SK_E_NO_SHEET;
    }
    if ((o->use_index & 3u) && !s->in[SK_IN_AUX1]) {
        ctx->err = "index reads need sk_limits.aux_streams = 1";
        return SK_E_INVALID;
    }
    if (o->fused_trim_min_baseq > 255) return SK_E_INVALID;
    int rc = begin_op(ctx, s, OP_DEMUX1);
    if (rc) return rc;
    const uint32_t S = ctx->S;
    s->paired = s->in_len[SK_IN_R2] > 0;
    CK(cudaMemsetAsync(s->counts, 0, (uint64_t)(S + 2) * 8, s->stream));
    const uint64_t limit = o->rec_limit ? o->rec_limit : ~0ull;

    // index reads (--index1 / --index2): OP_SCAN passes produce (seq_off, seq_len, flags) per record
    int idx_stream[2];
    uint32_t n_index = 0;
    if (o->use_index & 1u) idx_stream[n_index++] = SK_IN_AUX1;
    if (o->use_index & 2u) idx_stream[n_index++] = SK_IN_AUX2;
    for (uint32_t q = 0; q < n_index; q++) {
        KParams k;
        base_params(ctx, s, idx_stream[q], k);
        k.scan_out = s->scan_tab[q];
        k.scan_cap = ctx->lim.max_records;
        k.rec_limit = limit;  // so that consumed[] marks where record `limit` starts (error replay)
        rc = run_pass(ctx, s, idx_stream[q], OP_SCAN, k, false);
        if (rc) return rc;
    }
    auto demux_params = [&](int which, int mate, KParams &p) {
        base_params(ctx, s, which, p);
        p.rec_limit = limit;
        p.fused_trim = o->fused_trim_min_baseq;
        p.out = o->no_output ? nullptr : s->out[mate];
        p.out_cap = s->out_cap;
        p.sheet.planes = ctx->d_planes;
        p.sheet.umask = ctx->d_umask;
        p.sheet.lut = ctx->d_lut;
        p.sheet.S = S;
        p.sheet.L = ctx->L;
        p.sheet.Umax = ctx->Umax;
        p.sheet.wide = ctx->wide;
        p.sheet.fast.n_classes = ctx->f_classes;
        p.sheet.fast.nw = ctx->f_nw;
        p.sheet.fast.tsize = ctx->f_tsize;
        p.sheet.fast.cls = ctx->d_fcls;
        p.sheet.fast.table = ctx->d_ftab;
        p.sheet.fast.skeys = ctx->d_skeys;
        p.assign = s->assign;
        p.umi = s->umi;
        p.lens = s->lens[mate];
        p.chunk_base = s->chunk_base[mate];
        p.counts = s->counts;
        p.events = s->events;
        p.events_cap = (uint32_t)std::min<uint64_t>(ctx->lim.max_records, 0xFFFFFFFFull);
        p.n_index = n_index;
        p.r1_stats = s->stats + SK_IN_R1;
        for (uint32_t q = 0; q < n_index; q++) {
            p.ext_tab[q] = s->scan_tab[q];
            p.ext_data[q] = s->in[idx_stream[q]];
            p.ext_stats[q] = s->stats + idx_stream[q];
        }
    };
    KParams p1;
    demux_params(SK_IN_R1, 0, p1);
    rc = run_pass(ctx, s, SK_IN_R1, OP_DEMUX1, p1, false);
    if (rc) return rc;
    if (s->paired) {
        KParams p2;
        demux_params(SK_IN_R2, 1, p2);
        rc = run_pass(ctx, s, SK_IN_R2, OP_DEMUX2, p2, false);
        if (rc) return rc;
    }
    return end_op(ctx, s);
}

extern "C" int sk_wait(sk_ctx *ctx, uint32_t slot, sk_result *res) {
    Slot *s = get_slot(ctx, slot);
    if (!s) return SK_E_INVALID;
    CK(cudaStreamSynchronize(s->stream));
    if (!res) return SK_OK;
    memset(res, 0, sizeof *res);
    if (s->last_op < 0) return SK_OK;
    const DevStats *h = s->stats_h;
    unsigned long long best_key = ~0ull;
    uint32_t flags = 0;
    for (int i = 0; i < SK_N_INPUTS; i++) {
        res->n_lines[i] = h[i].n_lines;
        res->consumed[i] = h[i].consumed;
        if (h[i].err_key) best_key = std::min(best_key, ~h[i].err_key);
        flags |= h[i].flags;
    }
    res->n_records = h[SK_IN_R1].n_records;
    res->gpu_launches = s->launches;
    if (ctx->profiling)
        for (int i = 0; i < SK_N_INPUTS; i++)
            if (s->pass_ran[i]) cudaEventElapsedTime(&res->pass_ms[i], s->ev[i][0], s->ev[i][1]);
    if (s->last_op == OP_DEMUX1) {
        res->out_bytes[0] = h[SK_IN_R1].out_bytes;
        res->out_extent[0] = h[SK_IN_R1].out_cursor;
        res->out_bytes[1] = h[SK_IN_R2].out_bytes;
        res->out_extent[1] = h[SK_IN_R2].out_cursor;
        res->n_chunks[0] = s->n_chunks[SK_IN_R1];
        res->n_chunks[1] = s->n_chunks[SK_IN_R2];
        res->n_events = std::min<uint32_t>(h[SK_IN_R1].n_events, (uint32_t)ctx->lim.max_records);
        if (s->paired && h[SK_IN_R2].n_records < h[SK_IN_R1].n_records) flags |= F_MATE_COUNT;
        for (int i = SK_IN_AUX1; i <= SK_IN_AUX2; i++)
            if (s->pass_ran[i] && h[i].n_records < h[SK_IN_R1].n_records) flags |= F_MATE_COUNT;
        // counters live on the device; fetch total / identified for convenience
        unsigned long long ti[2] = {0, 0};
        CK(cudaMemcpy(ti, s->counts + ctx->S, 16, cudaMemcpyDeviceToHost));
        res->total_reads = ti[0];
        res->identified_reads = ti[1];
    } else {
        res->out_bytes[0] = h[SK_IN_R1].out_bytes;
        res->out_extent[0] = h[SK_IN_R1].out_extent;
    }
    res->flags = flags & 0xFFu;
    if (best_key != ~0ull) {
        res->status = (int32_t)(best_key & 0xFFu);
        res->err_record = best_key >> 8;
    }
    if (flags & F_NON_ASCII) {  // refuse non-ASCII batches outright
        res->status = SK_DATA_NON_ASCII;
        res->err_record = 0;
    }
    return SK_OK;
}

// ------------------------------------------------------------------------------------------------
// outputs
// ------------------------------------------------------------------------------------------------
extern "C" const void *sk_out_dev(sk_ctx *ctx, uint32_t slot, uint32_t which) {
    Slot *s = get_slot(ctx, slot);
    return (s && which < 2) ? s->out[which] : nullptr;
}
extern "C" int sk_download_out(sk_ctx *ctx, uint32_t slot, uint32_t which, void *host, uint64_t n) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= 2 || n > s->out_cap) return SK_E_INVALID;
    if (n) CK(cudaMemcpyAsync(host, s->out[which], n, cudaMemcpyDeviceToHost, s->stream));
    return SK_OK;
}
extern "C" int sk_download_demux_tables(sk_ctx *ctx, uint32_t slot, uint32_t which, uint64_t *chunk_base,
                                        uint16_t *lens) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= 2 || !s->lens[which]) return SK_E_INVALID;
    const uint32_t nc = s->n_chunks[which == 0 ? SK_IN_R1 : SK_IN_R2];
    if (nc) {
        CK(cudaMemcpyAsync(chunk_base, s->chunk_base[which], (uint64_t)nc * 8, cudaMemcpyDeviceToHost, s->stream));
        if (ctx->S)
            CK(cudaMemcpyAsync(lens, s->lens[which], (uint64_t)nc * ctx->S * 2, cudaMemcpyDeviceToHost, s->stream));
    }
    return SK_OK;
}
extern "C" int sk_download_counts(sk_ctx *ctx, uint32_t slot, uint64_t *counts) {
    Slot *s = get_slot(ctx, slot);
    if (!s || !s->counts) return SK_E_INVALID;
    CK(cudaMemcpyAsync(counts, s->counts, (uint64_t)(ctx->S + 2) * 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SK_OK;
}
extern "C" const void *sk_counts_dev(sk_ctx *ctx, uint32_t slot) {
    Slot *s = get_slot(ctx, slot);
    return s ? s->counts : nullptr;
}
extern "C" int sk_download_events(sk_ctx *ctx, uint32_t slot, sk_event *events, uint32_t cap) {
    Slot *s = get_slot(ctx, slot);
    if (!s || !s->events) return SK_E_INVALID;
    CK(cudaStreamSynchronize(s->stream));
    uint32_t n = std::min<uint32_t>(s->stats_h[SK_IN_R1].n_events, (uint32_t)ctx->lim.max_records);
    n = std::min(n, cap);
    if (n) CK(cudaMemcpy(events, s->events, (uint64_t)n * sizeof(Event), cudaMemcpyDeviceToHost));
    std::sort(events, events + n, [](const sk_event &a, const sk_event &b) { return a.record < b.record; });
    return (int)n;
}
extern "C" int sk_download_assign(sk_ctx *ctx, uint32_t slot, int16_t *assign, uint64_t n_records) {
    Slot *s = get_slot(ctx, slot);
    if (!s || !s->assign || n_records > ctx->lim.max_records) return SK_E_INVALID;
    if (n_records) CK(cudaMemcpyAsync(assign, s->assign, n_records * 2, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SK_OK;
}

extern "C" uint64_t sk_demux_gather(const uint8_t *out_host, const uint64_t *chunk_base, const uint16_t *lens,
                                    uint32_t n_chunks, uint32_t S, uint32_t smp, uint8_t *dst, uint64_t dst_cap) {
    uint64_t total = 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint16_t *row = lens + (uint64_t)c * S;
        const uint32_t len = row[smp];
        if (!len) continue;
        uint64_t off = chunk_base[c];
        for (uint32_t t = 0; t < smp; t++) off += row[t];
        if (dst && total + len <= dst_cap) memcpy(dst + total, out_host + off, len);
        total += len;
    }
    return total;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: the one collective on the path.  NCCL is resolved at run time so the
// library also loads where libnccl is absent.
// ------------------------------------------------------------------------------------------------
static void *nccl_sym(sk_ctx *ctx, const char *name) {
    static void *h = nullptr;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    void *f = h ? dlsym(h, name) : nullptr;
    if (!f) ctx->err = std::string("libnccl.so.2 / ") + name + " not found";
    return f;
}
extern "C" int sk_allreduce_counts(sk_ctx *ctx, uint32_t slot, void *nccl_comm) {
    Slot *s = get_slot(ctx, slot);
    if (!s || !s->counts || !nccl_comm) return SK_E_INVALID;
    typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    allreduce_fn fn = (allreduce_fn)nccl_sym(ctx, "ncclAllReduce");
    if (!fn) return SK_E_UNSUPPORTED;
    // ncclUint64 = 5, ncclSum = 0
    int rc = fn(s->counts, s->counts, (size_t)ctx->S + 2, 5, 0, nccl_comm, s->stream);
    if (rc != 0) {
        ctx->err = "ncclAllReduce failed with code " + std::to_string(rc);
        return SK_E_CUDA;
    }
    return SK_OK;
}
extern "C" int sk_nccl_unique_id(sk_ctx *ctx, void *id128) {
    if (!ctx || !id128) return SK_E_INVALID;
    typedef int (*fn_t)(void *);
    fn_t fn = (fn_t)nccl_sym(ctx, "ncclGetUniqueId");
    if (!fn) return SK_E_UNSUPPORTED;
    return fn(id128) == 0 ? SK_OK : SK_E_CUDA;
}
extern "C" int sk_nccl_comm_init(sk_ctx *ctx, const void *id128, int nranks, int rank, void **comm) {
    if (!ctx || !id128 || !comm) return SK_E_INVALID;
    cudaSetDevice(ctx->device);
    struct Id { char b[128]; } id;
    memcpy(&id, id128, 128);
    typedef int (*fn_t)(void **, int, Id, int);
    fn_t fn = (fn_t)nccl_sym(ctx, "ncclCommInitRank");
    if (!fn) return SK_E_UNSUPPORTED;
    int rc = fn(comm, nranks, id, rank);
    if (rc != 0) {
        ctx->err = "ncclCommInitRank failed with code " + std::to_string(rc);
        return SK_E_CUDA;
    }
    return SK_OK;
}
extern "C" int sk_nccl_comm_destroy(sk_ctx *ctx, void *comm) {
    if (!ctx || !comm) return SK_E_INVALID;
    typedef int (*fn_t)(void *);
    fn_t fn = (fn_t)nccl_sym(ctx, "ncclCommDestroy");
    if (!fn) return SK_E_UNSUPPORTED;
    return fn(comm) == 0 ? SK_OK : SK_E_CUDA;
}

// ------------------------------------------------------------------------------------------------
// synthetic workloads
// ------------------------------------------------------------------------------------------------
extern "C" int sk_synth_fastq(sk_ctx *ctx, uint32_t slot, uint32_t which, const sk_synth_spec *spec, uint64_t *n) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= SK_N_INPUTS || !spec || !s->in[which]) return SK_E_INVALID;
    if (spec->n_pairs > ctx->lim.max_records) return SK_E_TOO_LARGE;
    if (spec->with_bc && (!ctx->have_sheet || !ctx->S)) return SK_E_NO_SHEET;
    uint64_t bytes = 0;
    const char *err = nullptr;
    int rc = launch_synth(s->stream, s->in[which], s->in_cap[which], *spec, ctx->d_sheet_raw, ctx->S, ctx->L,
                          s->synth_tmp, &bytes, &err);
    if (rc != 0) {
        ctx->err = std::string("synthetic generator: ") + (err ? err : "?");
        return rc;
    }
    s->in_len[which] = bytes;
    if (n) *n = bytes;
    return SK_OK;
}
