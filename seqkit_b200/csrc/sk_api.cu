// sk_api.cu -- host side of the C ABI in include/seqkit_b200.h: contexts, slots, sample-sheet
// packing, operator launch sequences, result collection.  No CPU implementation of any operator
// lives here: every operator enqueues the chunk-engine kernels of sk_kernels.cu.
#include <cuda_runtime.h>
#include <cmath>
#include <dlfcn.h>
#include <sched.h>
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
static_assert(sizeof(sk_group) == sizeof(Group) && sizeof(sk_chunk_row) == sizeof(ChunkRow), "demux table layout");

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
    unsigned long long *ti_h = nullptr;  // pinned: total_reads, identified_reads of the last demultiplex
    int16_t *assign = nullptr;
    uint8_t *umi = nullptr;
    uint64_t umi_cap = 0;
    Group *groups[2] = {nullptr, nullptr};
    ChunkRow *rows[2] = {nullptr, nullptr};
    unsigned long long *counts = nullptr;
    Event *events = nullptr;
    RecRef *scan_tab[2] = {nullptr, nullptr};
    uint4 *bc_inline = nullptr;  // add barcode: 32 bytes of barcode per record (sk_lineops.cu:sk_recref_kernel)
    uint64_t *synth_tmp = nullptr;
    // per-sample compaction (sk_compact.cu): mate 1 -> cbuf, mate 2 -> out[0] (free once mate 1 is compacted)
    uint8_t *cbuf = nullptr;
    void *cwork = nullptr;
    unsigned long long *slices = nullptr;     // [2][(Smax + 1) * 2]
    unsigned long long *piece_dst = nullptr;  // [max_records]
    bool want_compact = false, compacted = false;
    // line engine (sk_lineops.cu; sk_limits.reserved bit 9)
    void *lwork = nullptr;
    void *stats_tab = nullptr;
    uint32_t h_cap = 0;
    uint32_t line_op = 0;
    // description of the last operator, for sk_wait
    int last_op = -1;
    bool paired = false;
    uint32_t n_chunks[SK_N_INPUTS] = {0, 0, 0, 0};
    uint32_t launches = 0;
    bool pass_ran[SK_N_INPUTS] = {false, false, false, false};
    cudaEvent_t ev[SK_N_INPUTS][2] = {};
    // the request behind last_op, so that sk_wait can re-run it on the general engine
    bool used_fast = false;
    bool reran_general = false;
    bool ran_line = false;  // trim / mask by quality ended up on the line engine (sk_lineops.cu)
    bool no_inplace = false;  // mask: the in-place layout was refused by the data (F_NEED_ORDERED), ordered form from now on
    uint32_t req_min_baseq = 0;
    uint64_t req_rec_limit = 0;
    sk_demux_opts req_opts{};
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
    bool u_uniform = false;            // every sample has its 'U' at the same positions
    unsigned long long u_mask = 0;
    std::vector<uint8_t> sheet_raw;
    uint32_t *d_planes = nullptr, *d_umask = nullptr;
    uint8_t *d_lut = nullptr, *d_sheet_raw = nullptr;
    unsigned long long *d_totals = nullptr;  // [Smax + 2] counters summed over the batches of a run (sk_counts_accumulate)
    // pigeonhole index (HalfIdx)
    uint32_t h_classes = 0, h_nw = 0, h_nwp = 0, h_tsize = 0;
    uint32_t *d_hcls = nullptr, *d_skeys = nullptr;
    uint2 *d_htab = nullptr;
    uint16_t *d_hcand = nullptr;
    uint2 *d_ftab = nullptr;      // FastIdx
    uint16_t *d_fnext = nullptr;
    bool fast_sheet = false;  // the sheet's FastIdx is usable
    bool warp = true;   // warp engine (sk_warp.cu) for header-route demultiplex; SK_NO_WARP=1 disables
    // Warp engine for the two ordered operators as well: bit 0 = trim, bit 1 = mask (SK_WARP_STREAM=0: the general
    // engine for both).  Their tiles write in input order, so a tile needs the output sizes of all tiles before
    // it.  Mask knows its sizes right after the line table: second look-back on output bytes, 2.9 ms per 8 M
    // reads of 150 bp (the retired lean engine: 5.4 ms).  Trim knows them only after its plan, and waiting for the slowest
    // plan among a thousand predecessors costs more than a second pass (5.5 ms): with trim_gather (default,
    // SK_TRIM_GATHER=0 switches it off) the tiles write wherever the output cursor puts them (the spare output
    // buffer), a one-block scan turns their lengths into destinations and a gather kernel writes the stream in
    // input order: 3.0 ms (the retired lean engine: 5.3 ms).
    uint32_t warp_stream = 3;
    bool trim_gather = true;
    // mask by quality of a regular file keeps every record's length: tiles write at their input offsets, no second
    // look-back (SK_MASK_INPLACE=0 switches it off); data of any other shape raises F_NEED_ORDERED and runs again
    bool mask_inplace = true;
    uint32_t tile_lanes = GeoW::TILE_LANES;  // warp-engine tile = tile_lanes x 400 B; SK_TILE_LANES
    bool tile_auto = true;   // tile_lanes follows the record size of the data (no SK_TILE_LANES override)
    double rec_est = 0.0;    // bytes per record: peeked from the first batch, then measured by every operator
    bool fast = true;   // warp engine for trim / mask / header-route demultiplex; SK_NO_FAST=1: everything on the general engine
    int cfg = 0;  // chunk-engine geometry: 0 = CfgA (16 KiB chunks, 4 warps), 1 = CfgB (32 KiB chunks, 8 warps)
};

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return SK_E_CUDA;                                                            \
        }                                                                                \
    } while (0)

// engine of a pass: 0 = general (sk_kernels.cu), 2 = warp (sk_warp.cu); 1 was the lean engine (sk_fast.cu, retired in round 2)
enum { ENG_GENERAL = 0, ENG_WARP = 2 };
static uint32_t chunks_of(const sk_ctx *ctx, uint64_t n, int eng = ENG_GENERAL) {
    const uint64_t ch = eng == ENG_WARP ? (uint64_t)ctx->tile_lanes * GeoW::LANE_BYTES : (uint64_t)cfg_chunk_bytes(ctx->cfg);
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
        cudaFree(s.groups[i]);
        cudaFree(s.rows[i]);
        cudaFree(s.scan_tab[i]);
        if (i == 0) cudaFree(s.bc_inline);
    }
    cudaFree(s.tile_out);
    cudaFree(s.stats);
    cudaFreeHost(s.stats_h);
    cudaFreeHost(s.ti_h);
    cudaFree(s.assign);
    cudaFree(s.umi);
    cudaFree(s.counts);
    cudaFree(s.events);
    cudaFree(s.synth_tmp);
    cudaFree(s.lwork);
    cudaFree(s.stats_tab);
    cudaFree(s.cbuf);
    cudaFree(s.cwork);
    cudaFree(s.slices);
    cudaFree(s.piece_dst);
    for (int i = 0; i < SK_N_INPUTS; i++)
        for (int k = 0; k < 2; k++)
            if (s.ev[i][k]) cudaEventDestroy(s.ev[i][k]);
    if (s.stream) cudaStreamDestroy(s.stream);
}

extern "C" void sk_ctx_destroy(sk_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto &s : ctx->slots) free_slot(s);
    cudaFree(ctx->d_totals);
    cudaFree(ctx->d_planes);
    cudaFree(ctx->d_umask);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_sheet_raw);
    cudaFree(ctx->d_hcls);
    cudaFree(ctx->d_skeys);
    cudaFree(ctx->d_htab);
    cudaFree(ctx->d_hcand);
    cudaFree(ctx->d_ftab);
    cudaFree(ctx->d_fnext);
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
    ctx->cfg = (lim->reserved & 0xFFu) == 2 ? 1 : 0;  // reserved: 0/1 = 16 KiB chunks (default), 2 = 32 KiB chunks
    if (const char *e = getenv("SK_CFG")) ctx->cfg = atoi(e) ? 1 : 0;
    if (const char *e = getenv("SK_NO_FAST")) ctx->fast = atoi(e) == 0;
    if (const char *e = getenv("SK_NO_WARP")) ctx->warp = atoi(e) == 0;
    if (const char *e = getenv("SK_WARP_STREAM")) ctx->warp_stream = atoi(e) ? 3u : 0u;
    if (const char *e = getenv("SK_TRIM_GATHER")) ctx->trim_gather = atoi(e) != 0;
    if (const char *e = getenv("SK_MASK_INPLACE")) ctx->mask_inplace = atoi(e) != 0;
    if (const char *e = getenv("SK_TILE_LANES")) {
        ctx->tile_lanes = (uint32_t)std::min(30, std::max(8, atoi(e)));
        ctx->tile_auto = false;
    }
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
    // slice-table rows: one per chunk, or GeoW::ROUNDS per tile of the warp engine (smallest tile: 8 lanes)
    ctx->max_chunks = chunks_of(ctx, B, ENG_GENERAL) + 1;
    ctx->max_chunks = std::max(ctx->max_chunks, (uint32_t)(B / (8 * GeoW::LANE_BYTES) + 1) * GeoW::ROUNDS);
    const uint32_t Smax = lim->max_samples;
    // every chunk / round owns whole 32-byte sectors; a compacted buffer starts every sample on a 128-byte line
    const uint64_t out_cap = B + R * 72 + (uint64_t)ctx->max_chunks * 32 + 4096 + (uint64_t)Smax * 128;
    if (Smax) {
        CKC(cudaMalloc(&ctx->d_totals, (uint64_t)(Smax + 2) * 8));
        CKC(cudaMemset(ctx->d_totals, 0, (uint64_t)(Smax + 2) * 8));
    }
    ctx->slots.resize(lim->n_slots);
    for (auto &s : ctx->slots) {
        CKC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        for (int i = 0; i < SK_N_INPUTS; i++)
            for (int k = 0; k < 2; k++) CKC(cudaEventCreate(&s.ev[i][k]));
        const int nin = lim->aux_streams ? SK_N_INPUTS : 2;
        for (int i = 0; i < nin; i++) {
            CKC(cudaMalloc(&s.in[i], B + 64));
            s.in_cap[i] = B;
            CKC(cudaMalloc(&s.tile_lines[i], ((uint64_t)ctx->max_chunks + 8) * 8));  // look-back reads whole 4-entry blocks
        }
        for (int i = 0; i < 2; i++) CKC(cudaMalloc(&s.out[i], out_cap + 64));
        s.out_cap = out_cap;
        CKC(cudaMalloc(&s.tile_out, ((uint64_t)ctx->max_chunks + 8) * 8));
        CKC(cudaMalloc(&s.stats, sizeof(DevStats) * SK_N_INPUTS));
        CKC(cudaMallocHost(&s.stats_h, sizeof(DevStats) * SK_N_INPUTS));
        memset(s.stats_h, 0, sizeof(DevStats) * SK_N_INPUTS);
        CKC(cudaMallocHost(&s.ti_h, 16));
        s.ti_h[0] = s.ti_h[1] = 0;
        CKC(cudaMalloc(&s.synth_tmp, (R + 1) * 8));
        if (Smax) {
            CKC(cudaMalloc(&s.assign, R * 2));
            for (int i = 0; i < 2; i++) {
                CKC(cudaMalloc(&s.groups[i], R * sizeof(Group)));
                CKC(cudaMalloc(&s.rows[i], (uint64_t)ctx->max_chunks * sizeof(ChunkRow)));
            }
            CKC(cudaMalloc(&s.counts, (uint64_t)(Smax + 2) * 8));
            CKC(cudaMalloc(&s.events, R * sizeof(Event)));
            if (Smax <= 4096 && !(lim->reserved & 0x100u)) {  // (reserved bit 8: no compaction buffers)
                CKC(cudaMalloc(&s.cbuf, out_cap + 64));
                CKC(cudaMalloc(&s.cwork, compact_work_bytes(ctx->max_chunks, Smax)));
                CKC(cudaMalloc(&s.slices, (uint64_t)(Smax + 1) * 16 * 2));
                CKC(cudaMalloc(&s.piece_dst, R * 8));
            }
        }
        if (lim->aux_streams)
            for (int i = 0; i < 2; i++) CKC(cudaMalloc(&s.scan_tab[i], R * sizeof(RecRef)));
            CKC(cudaMalloc(&s.bc_inline, R * 32));
        // the line engine's work area: the line operators (reserved bit 9), and trim / mask by quality as the last resort
        CKC(cudaMalloc(&s.lwork, lineops_work_bytes(B, R)));
        if (lim->reserved & 0x200u) {  // statistics table
            uint32_t cap = 1024;
            while ((uint64_t)cap < 2 * R && cap < (1u << 30)) cap <<= 1;
            s.h_cap = cap;
            CKC(cudaMalloc(&s.stats_tab, (uint64_t)cap * 36 + 64));
        }
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
// CPUs of the NUMA node a device hangs off (sysfs: the PCI device's numa_node, the node's cpulist).
static bool device_node_cpus(int device, cpu_set_t *set, int *node_out) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return false;
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    char path[256];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    char list[4096] = {0};
    const bool ok = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int n = 0;
    for (char *q = list; *q;) {  // "0-31,64-95"
        char *e;
        const long a = strtol(q, &e, 10);
        if (e == q) break;
        long b = a;
        if (*e == '-') b = strtol(e + 1, &e, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) {
            CPU_SET((int)c, set);
            n++;
        }
        q = (*e == ',') ? e + 1 : e;
        if (*e != ',') break;
    }
    if (node_out) *node_out = node;
    return n > 0;
}
extern "C" int sk_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
extern "C" int sk_bind_thread_to_device(int device) {
    cpu_set_t set;
    int node = -1;
    if (!device_node_cpus(device, &set, &node)) return -1;
    // keep to the CPUs this process may use at all (containers hand out subsets)
    cpu_set_t cur, both;
    if (sched_getaffinity(0, sizeof cur, &cur) == 0) {
        CPU_AND(&both, &cur, &set);
        if (CPU_COUNT(&both) == 0) return -1;
        set = both;
    }
    return sched_setaffinity(0, sizeof set, &set) == 0 ? node : -1;
}
// Pinned, portable (usable from every device's streams) and allocated while the calling thread sits on the
// NUMA node of the context's device, so that first touch puts the pages next to the GPU's PCIe root.
extern "C" void *sk_pinned_alloc(sk_ctx *ctx, uint64_t bytes) {
    if (!ctx || !bytes) return nullptr;
    cudaSetDevice(ctx->device);
    cpu_set_t old, node;
    const bool have_old = sched_getaffinity(0, sizeof old, &old) == 0;
    bool moved = false;
    if (have_old && !getenv("SK_NO_NUMA") && device_node_cpus(ctx->device, &node, nullptr)) {
        cpu_set_t both;
        CPU_AND(&both, &old, &node);
        if (CPU_COUNT(&both) > 0) moved = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    void *p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (moved) sched_setaffinity(0, sizeof old, &old);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaHostAlloc failed: ") + cudaGetErrorString(e);
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
    unsigned long long u_first = 0;
    bool u_same = true;
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
        if (s == 0) u_first = um;
        else if (um != u_first) u_same = false;
    }

    // ---- pigeonhole index (HalfIdx, sk_internal.h): classes of identical care mask; per class the cared
    // positions are split into two halves and every sample is filed under the bytes of each half.
    const uint32_t nw = (L + 3) / 4, nwp = std::max(4u, (nw + 3u) & ~3u);
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
    uint32_t h_classes = (S && L && cls_care.size() <= 4) ? (uint32_t)cls_care.size() : 0;
    std::vector<std::vector<uint32_t>> halfpos[2];  // [half][class] -> positions
    for (uint32_t c = 0; c < h_classes; c++) {
        std::vector<uint32_t> pos;
        for (uint32_t q = 0; q < L; q++)
            if (cls_care[c][q]) pos.push_back(q);
        if (pos.size() < 2) {  // one mismatch could hide anywhere: no half is guaranteed to match
            h_classes = 0;
            break;
        }
        const size_t na = (pos.size() + 1) / 2;
        halfpos[0].push_back(std::vector<uint32_t>(pos.begin(), pos.begin() + na));
        halfpos[1].push_back(std::vector<uint32_t>(pos.begin() + na, pos.end()));
    }
    uint32_t tsize = 16;
    while (2 * tsize < 3 * S) tsize <<= 1;  // load factor <= 2/3
    if (chunk_kernel_smem_bytes(ctx->cfg, S, wide, h_classes, nwp) > 227 * 1024) {
        ctx->err = "sample sheet does not fit in shared memory";
        return SK_E_UNSUPPORTED;
    }
    std::vector<uint32_t> skeys((size_t)std::max(S, 1u) * nwp, 0u);
    for (uint32_t s = 0; s < S; s++)
        for (uint32_t q = 0; q < L; q++) {
            const uint8_t b = barcodes[(uint64_t)s * L + q] & cls_care[cls_of[s]][q];
            skeys[(size_t)s * nwp + q / 4] |= (uint32_t)b << (8 * (q % 4));
        }
    std::vector<uint32_t> hcls((size_t)std::max(h_classes, 1u) * HIDX_CLS_ROWS * nwp, 0u);
    std::vector<uint2> htab((size_t)std::max(h_classes, 1u) * 2 * tsize, make_uint2(0u, 0u));
    std::vector<uint16_t> hcand;
    uint64_t rng = 0x9E3779B97F4A7C15ull ^ ((uint64_t)S << 32) ^ L;
    auto next = [&]() {
        rng ^= rng << 13;
        rng ^= rng >> 7;
        rng ^= rng << 17;
        return (uint32_t)(rng >> 16);
    };
    for (uint32_t c = 0; c < h_classes; c++) {
        uint32_t *row = &hcls[(size_t)c * HIDX_CLS_ROWS * nwp];
        for (uint32_t q = 0; q < L; q++) row[q / 4] |= (uint32_t)cls_care[c][q] << (8 * (q % 4));
        for (uint32_t h = 0; h < 2; h++) {
            uint32_t *hm = row + (1 + 3 * h) * nwp;
            for (uint32_t q : halfpos[h][c]) hm[q / 4] |= 0xFFu << (8 * (q % 4));
            for (uint32_t w = 0; w < nwp; w++) {
                hm[nwp + w] = next() | 1u;
                hm[2 * nwp + w] = next() | 1u;
            }
            // group the samples of the class by their bytes on this half
            std::vector<uint32_t> members;
            for (uint32_t s = 0; s < S; s++)
                if (cls_of[s] == c) members.push_back(s);
            auto halfkey = [&](uint32_t s, uint32_t w) { return skeys[(size_t)s * nwp + w] & hm[w]; };
            std::vector<char> done(members.size(), 0);
            uint2 *tab = &htab[(size_t)(c * 2 + h) * tsize];
            for (size_t i = 0; i < members.size(); i++) {
                if (done[i]) continue;
                const uint32_t s0 = members[i];
                const uint32_t start = (uint32_t)hcand.size();
                for (size_t k = i; k < members.size(); k++) {
                    if (done[k]) continue;
                    bool same = true;
                    for (uint32_t w = 0; w < nw && same; w++) same = halfkey(members[k], w) == halfkey(s0, w);
                    if (same) {
                        done[k] = 1;
                        hcand.push_back((uint16_t)members[k]);
                    }
                }
                const uint32_t count = (uint32_t)hcand.size() - start;
                uint32_t h1 = 0, h2 = 0;
                for (uint32_t w = 0; w < nw; w++) {
                    h1 += halfkey(s0, w) * hm[nwp + w];
                    h2 += halfkey(s0, w) * hm[2 * nwp + w];
                }
                h1 ^= h1 >> 15;
                uint32_t slot = h1 & (tsize - 1);
                while (tab[slot].y >> 16) slot = (slot + 1) & (tsize - 1);
                tab[slot] = make_uint2(h2, start | (count << 16));
            }
        }
    }
    if (hcand.size() > 0xFFFFu) h_classes = 0;  // list offsets are 16-bit
    if (hcand.empty()) hcand.push_back(0);
    // the same index in the form the warp engine probes (FastIdx): same slots, tag = the slot hash before mixing, first
    // sample in the slot, the rest chained.  Tags must be unique inside a table so that a probe can stop
    // at the first tag match; otherwise the warp engine is not used for this sheet.
    std::vector<uint2> ftab(htab.size(), make_uint2(0u, 0u));
    std::vector<uint16_t> fnext((size_t)std::max(h_classes, 1u) * 2 * std::max(S, 1u), (uint16_t)0xFFFFu);
    bool fast_ok = h_classes != 0;
    for (uint32_t c = 0; c < h_classes && fast_ok; c++)
        for (uint32_t h = 0; h < 2 && fast_ok; h++) {
            const uint32_t t = c * 2 + h;
            const uint32_t *hm = &hcls[(size_t)c * HIDX_CLS_ROWS * nwp] + (1 + 3 * h) * nwp;
            std::vector<uint32_t> tags;
            for (uint32_t sl = 0; sl < tsize; sl++) {
                const uint2 e = htab[(size_t)t * tsize + sl];
                const uint32_t count = e.y >> 16, start = e.y & 0xFFFFu;
                if (!count) continue;
                const uint32_t s0 = hcand[start];
                uint32_t h1 = 0;
                for (uint32_t w = 0; w < nw; w++) h1 += (skeys[(size_t)s0 * nwp + w] & hm[w]) * hm[nwp + w];
                tags.push_back(h1);
                ftab[(size_t)t * tsize + sl] = make_uint2(h1, (s0 + 1u) | (count > 1 ? 1u << 16 : 0u));
                for (uint32_t i = 0; i + 1 < count; i++) fnext[(size_t)t * S + hcand[start + i]] = hcand[start + i + 1];
            }
            std::sort(tags.begin(), tags.end());
            if (std::adjacent_find(tags.begin(), tags.end()) != tags.end()) fast_ok = false;
        }
    ctx->fast_sheet = fast_ok;

    cudaFree(ctx->d_planes);
    cudaFree(ctx->d_umask);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_sheet_raw);
    cudaFree(ctx->d_hcls);
    cudaFree(ctx->d_skeys);
    cudaFree(ctx->d_htab);
    cudaFree(ctx->d_hcand);
    cudaFree(ctx->d_ftab);
    cudaFree(ctx->d_fnext);
    ctx->d_ftab = nullptr;
    ctx->d_fnext = nullptr;
    ctx->d_planes = ctx->d_umask = ctx->d_hcls = ctx->d_skeys = nullptr;
    ctx->d_lut = ctx->d_sheet_raw = nullptr;
    ctx->d_htab = nullptr;
    ctx->d_hcand = nullptr;
    CK(cudaMalloc(&ctx->d_planes, std::max<size_t>(planes.size() * 4, 16)));
    CK(cudaMalloc(&ctx->d_umask, std::max<size_t>(umask.size() * 4, 16)));
    CK(cudaMalloc(&ctx->d_lut, 256));
    CK(cudaMalloc(&ctx->d_sheet_raw, std::max<size_t>((size_t)S * L, 16)));
    CK(cudaMalloc(&ctx->d_hcls, hcls.size() * 4));
    CK(cudaMalloc(&ctx->d_skeys, std::max<size_t>(skeys.size() * 4, 16)));
    CK(cudaMalloc(&ctx->d_htab, htab.size() * 8));
    CK(cudaMalloc(&ctx->d_hcand, hcand.size() * 2));
    CK(cudaMalloc(&ctx->d_ftab, ftab.size() * 8));
    CK(cudaMalloc(&ctx->d_fnext, fnext.size() * 2));
    CK(cudaMemcpy(ctx->d_ftab, ftab.data(), ftab.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_fnext, fnext.data(), fnext.size() * 2, cudaMemcpyHostToDevice));
    if (S) {
        CK(cudaMemcpy(ctx->d_planes, planes.data(), planes.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->d_umask, umask.data(), umask.size() * 4, cudaMemcpyHostToDevice));
        if (L) CK(cudaMemcpy(ctx->d_sheet_raw, barcodes, (size_t)S * L, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->d_skeys, skeys.data(), skeys.size() * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(ctx->d_lut, lut, 256, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_hcls, hcls.data(), hcls.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_htab, htab.data(), htab.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_hcand, hcand.data(), hcand.size() * 2, cudaMemcpyHostToDevice));
    ctx->sheet_raw.assign(barcodes, barcodes + (size_t)S * L);
    ctx->S = S;
    ctx->L = L;
    ctx->Umax = Umax;
    ctx->u_uniform = u_same;
    ctx->u_mask = u_first;
    ctx->wide = wide;
    ctx->h_classes = h_classes;
    ctx->h_nw = nw;
    ctx->h_nwp = nwp;
    ctx->h_tsize = tsize;
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
    s->used_fast = false;
    s->launches = 0;
    s->want_compact = s->compacted = false;
    s->ran_line = false;
    for (int i = 0; i < SK_N_INPUTS; i++) {
        s->pass_ran[i] = false;
        s->n_chunks[i] = 0;
    }
    CK(cudaMemsetAsync(s->stats, 0, sizeof(DevStats) * SK_N_INPUTS, s->stream));
    return SK_OK;
}

static void base_params(sk_ctx *ctx, Slot *s, int which, KParams &p, int eng = ENG_GENERAL) {
    memset(&p, 0, sizeof p);
    p.in = s->in[which];
    p.n = s->in_len[which];
    p.n_chunks = chunks_of(ctx, p.n, eng);
    p.tile_lanes = ctx->tile_lanes;
    p.lpr = 4;
    p.max_records = ctx->lim.max_records;
    p.rec_limit = ~0ull;
    p.final_batch = 1;
    p.fused_trim = -1;
    p.tile_lines = s->tile_lines[which];
    p.tile_out = s->tile_out;
    p.stats = s->stats + which;
    s->n_chunks[which] = eng == ENG_WARP ? p.n_chunks * GeoW::ROUNDS : p.n_chunks;  // slice-table rows
}

static int run_pass(sk_ctx *ctx, Slot *s, int which, int op, const KParams &p, bool ordered_out, int eng = ENG_GENERAL) {
    if (p.n_chunks == 0) return SK_OK;
    // look-back words (the warp engine keeps 8 + 2 bytes per tile, sk_warp.cu:wlb_agg)
    CK(cudaMemsetAsync(p.tile_lines, 0, eng == ENG_WARP ? (uint64_t)p.n_chunks * 10 + 64 : (uint64_t)p.n_chunks * 8, s->stream));
    if (ordered_out)  // (an unordered trim keeps 20 bytes per tile there: base, length, destination)
        CK(cudaMemsetAsync(p.tile_out, 0, eng == ENG_WARP ? (uint64_t)p.n_chunks * (p.unordered ? 20 : 10) + 64 : (uint64_t)p.n_chunks * 8, s->stream));
    const char *err = nullptr;
    if (ctx->profiling) CK(cudaEventRecord(s->ev[which][0], s->stream));
    int rc = eng == ENG_WARP ? launch_warp_kernel(op, p, ctx->sm_count, s->stream, &err)
                             : launch_chunk_kernel(ctx->cfg, op, p, ctx->sm_count, s->stream, &err);
    if (rc < 0) {
        ctx->err = std::string("kernel launch failed: ") + (err ? err : "?");
        return SK_E_CUDA;
    }
    if (rc > 0 && eng == ENG_WARP && p.unordered) {  // trim: scan of the tiles' lengths, gather into input order
        const int rc2 = launch_tile_gather(p, ctx->sm_count, s->stream, &err);
        if (rc2 < 0) {
            ctx->err = std::string("kernel launch failed: ") + (err ? err : "?");
            return SK_E_CUDA;
        }
        rc += rc2;
    }
    if (ctx->profiling) CK(cudaEventRecord(s->ev[which][1], s->stream));
    s->launches += (uint32_t)rc;
    s->pass_ran[which] = true;
    return SK_OK;
}

static int end_op(sk_ctx *ctx, Slot *s) {
    CK(cudaMemcpyAsync(s->stats_h, s->stats, sizeof(DevStats) * SK_N_INPUTS, cudaMemcpyDeviceToHost, s->stream));
    if (s->last_op == OP_DEMUX1 && s->counts)  // total / identified ride along (no blocking copy in sk_wait)
        CK(cudaMemcpyAsync(s->ti_h, s->counts + ctx->S, 16, cudaMemcpyDeviceToHost, s->stream));
    return SK_OK;
}

static int choose_tile_lanes(sk_ctx *ctx, Slot *s);
static int stream_op_enqueue(sk_ctx *ctx, Slot *s, int op, uint32_t min_baseq, uint64_t rec_limit, bool fast) {
    int rc = (op == OP_TRIM || op == OP_MASK) ? choose_tile_lanes(ctx, s) : SK_OK;
    if (rc) return rc;
    rc = begin_op(ctx, s, op);
    if (rc) return rc;
    KParams p;
    const bool want_warp = ctx->warp && (ctx->warp_stream & (op == OP_TRIM ? 1u : 2u)) != 0;
    int eng = fast && want_warp ? ENG_WARP : ENG_GENERAL;
    auto fill = [&](int e) {
        base_params(ctx, s, SK_IN_R1, p, e);
        p.min_baseq = min_baseq;
        p.rec_limit = rec_limit ? rec_limit : ~0ull;
        p.out = s->out[0];
        p.out_cap = s->out_cap;
    };
    fill(eng);
    if (eng == ENG_WARP && !warp_supported(op, p)) {
        eng = ENG_GENERAL;
        fill(eng);
    }
    if (eng == ENG_WARP && op == OP_TRIM && ctx->trim_gather) {  // unordered tiles into the spare output buffer, then gather
        p.unordered = 1;
        p.final_out = s->out[0];
        p.out = s->out[1];
    }
    if (eng == ENG_WARP && op == OP_MASK && ctx->mask_inplace && !s->no_inplace && p.rec_limit == ~0ull) p.inplace = 1;
    fast = eng != ENG_GENERAL;
    s->used_fast = fast;
    s->req_min_baseq = min_baseq;
    s->req_rec_limit = rec_limit;
    rc = run_pass(ctx, s, SK_IN_R1, op, p, true, eng);
    if (rc) return rc;
    return end_op(ctx, s);
}
static int stream_op(sk_ctx *ctx, uint32_t slot, int op, uint32_t min_baseq, uint64_t rec_limit) {
    Slot *s = get_slot(ctx, slot);
    if (!s || min_baseq > 255) return SK_E_INVALID;
    s->reran_general = false;
    s->no_inplace = false;
    return stream_op_enqueue(ctx, s, op, min_baseq, rec_limit, ctx->fast);
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

// Pass 1: (seq_off, seq_len, flags) of every barcode record, from the stream's global line table (sk_lineops.cu);
// pass 2: the reads -- on the warp engine when they are FASTQ (first with every record at its input offset plus a
// multiple of the first barcode's length, sk_wait repeats it in the ordered form when a record does not fit that),
// else on the general engine.
static int addbc_enqueue(sk_ctx *ctx, Slot *s, uint64_t rec_limit, bool fast) {
    int c_reads, c_bc;
    int rc = peek_first_byte(ctx, s, SK_IN_R1, &c_reads);
    if (rc) return rc;
    rc = peek_first_byte(ctx, s, SK_IN_AUX1, &c_bc);
    if (rc) return rc;
    rc = (fast && ctx->warp && c_reads == '@') ? choose_tile_lanes(ctx, s) : SK_OK;
    if (rc) return rc;
    rc = begin_op(ctx, s, OP_ADDBC);
    if (rc) return rc;
    s->req_rec_limit = rec_limit;
    // pass 1: where is the sequence line of every barcode record?
    const bool bc_fastx = (c_bc == '@' || c_bc == '>');
    // A barcode file whose first line is neither '@' nor '>' never yields a barcode (:20-27): every
    // read gets an empty one.  (A later '@' line in such a file is reported as mixed format.)
    if (s->lwork) {
        const char *err = nullptr;
        if (ctx->profiling) CK(cudaEventRecord(s->ev[SK_IN_AUX1][0], s->stream));
        const int n = launch_scan_table(s->in[SK_IN_AUX1], s->in_len[SK_IN_AUX1], c_bc == '>' ? 2u : 4u, bc_fastx ? (uint32_t)c_bc : 0xFFFFu,
                                        bc_fastx ? ~0ull : 0ull, 1u, s->scan_tab[0], s->bc_inline, ctx->lim.max_records, s->lwork, 1,
                                        ctx->lim.max_stream_bytes, ctx->lim.max_records, s->stats + SK_IN_AUX1, ctx->sm_count, s->stream, &err);
        if (n < 0) {
            ctx->err = std::string("record table launch failed: ") + (err ? err : "?");
            return SK_E_CUDA;
        }
        if (ctx->profiling) CK(cudaEventRecord(s->ev[SK_IN_AUX1][1], s->stream));
        s->launches += (uint32_t)n;
        s->pass_ran[SK_IN_AUX1] = true;
    } else {
        KParams q;
        base_params(ctx, s, SK_IN_AUX1, q);
        q.lpr = c_bc == '>' ? 2 : 4;
        q.head_char = bc_fastx ? (uint32_t)c_bc : 0;
        q.scan_out = s->scan_tab[0];
        q.scan_cap = ctx->lim.max_records;
        if (!bc_fastx) q.rec_limit = 0, q.head_char = 0xFFFF;
        rc = run_pass(ctx, s, SK_IN_AUX1, OP_SCAN, q, false);
        if (rc) return rc;
    }
    // pass 2: the reads
    KParams p;
    int eng = (fast && ctx->warp && c_reads == '@') ? ENG_WARP : ENG_GENERAL;
    auto fill = [&](int e) {
        base_params(ctx, s, SK_IN_R1, p, e);
        p.lpr = c_reads == '>' ? 2 : 4;
        p.head_char = c_reads == '>' ? '>' : '@';
        p.rec_limit = rec_limit ? rec_limit : ~0ull;
        p.out = s->out[0];
        p.out_cap = s->out_cap;
        p.ext_tab[0] = s->scan_tab[0];
        p.ext_data[0] = s->in[SK_IN_AUX1];
        p.ext_stats[0] = s->stats + SK_IN_AUX1;
        p.bc_inline = s->lwork ? s->bc_inline : nullptr;
    };
    fill(eng);
    if (eng == ENG_WARP && !warp_supported(OP_ADDBC, p)) {
        eng = ENG_GENERAL;
        fill(eng);
    }
    if (eng == ENG_WARP && !s->no_inplace && p.rec_limit == ~0ull) p.inplace = 1;
    s->used_fast = eng != ENG_GENERAL;
    rc = run_pass(ctx, s, SK_IN_R1, OP_ADDBC, p, true, eng);
    if (rc) return rc;
    return end_op(ctx, s);
}
// Add barcode on the line engine: reads of any length and density, UTF-8 lines ('@' and '>' reads alike).
static int line_addbc_enqueue(sk_ctx *ctx, Slot *s) {
    int c_reads, c_bc;
    int rc = peek_first_byte(ctx, s, SK_IN_R1, &c_reads);
    if (rc) return rc;
    rc = peek_first_byte(ctx, s, SK_IN_AUX1, &c_bc);
    if (rc) return rc;
    const uint64_t rec_limit = s->req_rec_limit;
    rc = begin_op(ctx, s, OP_ADDBC);
    if (rc) return rc;
    s->req_rec_limit = rec_limit;
    const bool bc_fastx = (c_bc == '@' || c_bc == '>');
    const char *err = nullptr;
    int n = launch_scan_table(s->in[SK_IN_AUX1], s->in_len[SK_IN_AUX1], c_bc == '>' ? 2u : 4u, bc_fastx ? (uint32_t)c_bc : 0xFFFFu,
                              bc_fastx ? ~0ull : 0ull, 1u, s->scan_tab[0], nullptr, ctx->lim.max_records, s->lwork, 1,
                              ctx->lim.max_stream_bytes, ctx->lim.max_records, s->stats + SK_IN_AUX1, ctx->sm_count, s->stream, &err);
    if (n >= 0) {
        s->launches += (uint32_t)n;
        s->pass_ran[SK_IN_AUX1] = true;
        n = launch_lineop(8 /* LOP_ADDBC */, s->in[SK_IN_R1], s->in_len[SK_IN_R1], s->in[SK_IN_AUX1], s->in_len[SK_IN_AUX1],
                          c_reads == '>' ? 2u : 4u, c_reads == '>' ? (uint32_t)'>' : (uint32_t)'@', 0, 0, rec_limit, s->out[0], s->out[1],
                          s->out_cap, s->lwork, ctx->lim.max_stream_bytes, ctx->lim.max_records, nullptr, 0, s->stats + SK_IN_R1,
                          ctx->sm_count, s->stream, &err, s->scan_tab[0], s->stats + SK_IN_AUX1);
    }
    if (n < 0) {
        ctx->err = std::string("line engine launch failed: ") + (err ? err : "?");
        return SK_E_CUDA;
    }
    s->launches += (uint32_t)n;
    s->pass_ran[SK_IN_R1] = true;
    s->ran_line = true;
    return end_op(ctx, s);
}
extern "C" int sk_add_barcode(sk_ctx *ctx, uint32_t slot, uint64_t rec_limit) {
    Slot *s = get_slot(ctx, slot);
    if (!s) return SK_E_INVALID;
    if (!s->in[SK_IN_AUX1]) {
        ctx->err = "sk_add_barcode needs sk_limits.aux_streams = 1";
        return SK_E_INVALID;
    }
    s->reran_general = false;
    s->no_inplace = false;
    return addbc_enqueue(ctx, s, rec_limit, ctx->fast);
}

// The line engine's operators (sk_lineops.cu): SK_IN_R1 (+ SK_IN_R2 for interleave) -> output stream 0 (+ 1 for
// deinterleave).  Framing by the stream's first byte: '@' = 4 lines per record, '>' = 2.
extern "C" int sk_line_op(sk_ctx *ctx, uint32_t slot, uint32_t op, uint32_t x, uint32_t y, uint64_t rec_limit) {
    Slot *s = get_slot(ctx, slot);
    if (!s || op > 5) return SK_E_INVALID;
    if (op == 2 && !s->stats_tab) {
        ctx->err = "SK_LOP_STATS needs a context created with sk_limits.reserved bit 9 (0x200)";
        return SK_E_INVALID;
    }
    int c = -1;
    int rc = peek_first_byte(ctx, s, SK_IN_R1, &c);
    if (rc) return rc;
    rc = begin_op(ctx, s, OP_LINE);
    if (rc) return rc;
    s->line_op = op;
    const uint32_t lpr = c == '>' ? 2u : 4u, head = c == '>' ? (uint32_t)'>' : (uint32_t)'@';
    const char *err = nullptr;
    const int n = launch_lineop((int)op, s->in[SK_IN_R1], s->in_len[SK_IN_R1], s->in[SK_IN_R2], s->in_len[SK_IN_R2], lpr, head, x, y,
                                rec_limit, s->out[0], s->out[1], s->out_cap, s->lwork, ctx->lim.max_stream_bytes, ctx->lim.max_records,
                                s->stats_tab, s->h_cap, s->stats + SK_IN_R1, ctx->sm_count, s->stream, &err);
    if (n < 0) {
        ctx->err = std::string("line operator launch failed: ") + (err ? err : "?");
        return SK_E_CUDA;
    }
    s->launches += (uint32_t)n;
    s->pass_ran[SK_IN_R1] = true;
    return end_op(ctx, s);
}
// Statistics: the distinct barcodes of the last sk_line_op(SK_LOP_STATS) -- (offset, length) of one occurrence
// (the earliest) in SK_IN_R1 and the number of records that carry it -- in no particular order.
extern "C" int sk_download_stats(sk_ctx *ctx, uint32_t slot, sk_stat_entry *entries, uint32_t cap, uint32_t *n) {
    Slot *s = get_slot(ctx, slot);
    if (!s || !s->stats_tab || !n) return SK_E_INVALID;
    CK(cudaStreamSynchronize(s->stream));
    const uint64_t hc = s->h_cap;
    const uint8_t *base = (const uint8_t *)s->stats_tab;
    const unsigned long long *list = (const unsigned long long *)(base + hc * 20);
    const uint32_t *n_list = (const uint32_t *)(base + hc * 36);
    uint32_t cnt = 0;
    CK(cudaMemcpy(&cnt, n_list, 4, cudaMemcpyDeviceToHost));
    *n = cnt;
    if (!entries || cap < cnt) return cnt ? SK_E_TOO_LARGE : SK_OK;
    std::vector<unsigned long long> tmp((size_t)cnt * 2);
    if (cnt) CK(cudaMemcpy(tmp.data(), list, (size_t)cnt * 16, cudaMemcpyDeviceToHost));
    for (uint32_t k = 0; k < cnt; k++) {
        entries[k].off = (uint32_t)(tmp[2 * (size_t)k] >> 32);
        entries[k].len = (uint32_t)tmp[2 * (size_t)k];
        entries[k].count = tmp[2 * (size_t)k + 1];
    }
    return SK_OK;
}

static int demux_enqueue(sk_ctx *ctx, Slot *s, const sk_demux_opts *o, bool fast);
extern "C" int sk_demultiplex(sk_ctx *ctx, uint32_t slot, const sk_demux_opts *o) {
    Slot *s = get_slot(ctx, slot);
    if (!s || !o) return SK_E_INVALID;
    s->reran_general = false;
    return demux_enqueue(ctx, s, o, ctx->fast);
}
// The warp engine wants about 31 records in a tile (32 lanes, one round): tile = the largest whole number
// of 400-byte lanes that holds at most 31.1 records of the size the data has, 29 lanes (11 600 B, 2x150 bp
// reads) at most -- three lanes stay the overhang.  The size comes from the previous operator's outcome
// (consumed bytes / records); before the first one, from the newlines of the first 64 KiB of mate 1.
static int choose_tile_lanes(sk_ctx *ctx, Slot *s) {
    if (!ctx->tile_auto) return SK_OK;
    if (ctx->rec_est <= 0.0 && s->in_len[SK_IN_R1]) {
        const size_t n = (size_t)std::min<uint64_t>(s->in_len[SK_IN_R1], 64u << 10);
        std::vector<uint8_t> head(n);
        CK(cudaMemcpyAsync(head.data(), s->in[SK_IN_R1], n, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        size_t lines = 0, last = 0;
        for (size_t i = 0; i < n; i++)
            if (head[i] == '\n') {
                lines++;
                if (lines % 4 == 0) last = i + 1;
            }
        if (lines >= 8) ctx->rec_est = (double)last / (double)(lines / 4);
    }
    if (ctx->rec_est > 0.0) {
        const double lanes = 31.1 * ctx->rec_est / (double)GeoW::LANE_BYTES;
        ctx->tile_lanes = (uint32_t)std::min<double>((double)GeoW::TILE_LANES, std::max(8.0, std::floor(lanes)));
    }
    return SK_OK;
}

static int demux_enqueue(sk_ctx *ctx, Slot *s, const sk_demux_opts *o, bool fast) {
    if (!ctx->have_sheet || !s->assign) {
        ctx->err = "sk_demultiplex: call sk_set_sheet first (and create the context with max_samples > 0)";
        return SK_E_NO_SHEET;
    }
    if ((o->use_index & 3u) && !s->in[SK_IN_AUX1]) {
        ctx->err = "index reads need sk_limits.aux_streams = 1";
        return SK_E_INVALID;
    }
    if (o->fused_trim_min_baseq > 255) return SK_E_INVALID;
    int rc = choose_tile_lanes(ctx, s);
    if (rc) return rc;
    rc = begin_op(ctx, s, OP_DEMUX1);
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
    // the warp engine does the header route on sheets its index can represent
    if (n_index || !ctx->h_classes || !ctx->fast_sheet || getenv("SK_NO_HIDX")) fast = false;
    s->used_fast = fast;
    {
        const sk_demux_opts keep = *o;  // o may alias s->req_opts (re-run)
        s->req_opts = keep;
    }
    for (uint32_t q = 0; q < n_index; q++) {
        if (s->lwork) {  // from the stream's line table: index reads are short, hundreds of records per 16 KiB
            const char *err = nullptr;
            const int w = idx_stream[q];
            if (ctx->profiling) CK(cudaEventRecord(s->ev[w][0], s->stream));
            const int n = launch_scan_table(s->in[w], s->in_len[w], 4u, 0u, limit, 1u, s->scan_tab[q], nullptr, ctx->lim.max_records, s->lwork, (int)q,
                                            ctx->lim.max_stream_bytes, ctx->lim.max_records, s->stats + w, ctx->sm_count, s->stream, &err);
            if (n < 0) {
                ctx->err = std::string("record table launch failed: ") + (err ? err : "?");
                return SK_E_CUDA;
            }
            if (ctx->profiling) CK(cudaEventRecord(s->ev[w][1], s->stream));
            s->launches += (uint32_t)n;
            s->pass_ran[w] = true;
            continue;
        }
        KParams k;
        base_params(ctx, s, idx_stream[q], k);
        k.scan_out = s->scan_tab[q];
        k.scan_cap = ctx->lim.max_records;
        k.rec_limit = limit;  // so that consumed[] marks where record `limit` starts (error replay)
        rc = run_pass(ctx, s, idx_stream[q], OP_SCAN, k, false);
        if (rc) return rc;
    }
    int eng = fast && ctx->warp ? ENG_WARP : ENG_GENERAL;
    if (eng == ENG_GENERAL) s->used_fast = false;
    auto demux_params = [&](int which, int mate, KParams &p) {
        base_params(ctx, s, which, p, eng);
        p.sheet.fidx.table = ctx->d_ftab;
        p.sheet.fidx.next = ctx->d_fnext;
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
        p.sheet.u_uniform = ctx->u_uniform ? 1u : 0u;
        p.sheet.u_mask = ctx->u_mask;
        p.sheet.wide = ctx->wide;
        p.sheet.hidx.n_classes = getenv("SK_NO_HIDX") ? 0u : ctx->h_classes;
        p.sheet.hidx.nw = ctx->h_nw;
        p.sheet.hidx.nwp = ctx->h_nwp;
        p.sheet.hidx.tsize = ctx->h_tsize;
        p.sheet.hidx.cls = ctx->d_hcls;
        p.sheet.hidx.table = ctx->d_htab;
        p.sheet.hidx.cand = ctx->d_hcand;
        p.sheet.hidx.skeys = ctx->d_skeys;
        p.assign = s->assign;
        p.umi = s->umi;
        p.groups = s->groups[mate];
        p.rows = s->rows[mate];
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
    if (eng == ENG_WARP && !warp_supported(OP_DEMUX1, p1)) {
        eng = ENG_GENERAL;
        s->used_fast = false;
        demux_params(SK_IN_R1, 0, p1);
    }
    rc = run_pass(ctx, s, SK_IN_R1, OP_DEMUX1, p1, false, eng);
    if (rc) return rc;
    if (s->paired) {
        KParams p2;
        demux_params(SK_IN_R2, 1, p2);
        rc = run_pass(ctx, s, SK_IN_R2, OP_DEMUX2, p2, false, eng);
        if (rc) return rc;
    }
    return end_op(ctx, s);
}

// Header-route demultiplex on the line engine (sk_lineops.cu): what neither chunk engine takes -- records longer than the
// general engine's overhang, more records in a chunk than it has slots for, UTF-8 in header and '+' lines.  Same tables and
// buffers as demux_enqueue leaves.
static int line_demux_enqueue(sk_ctx *ctx, Slot *s) {
    const sk_demux_opts o = s->req_opts;
    const uint32_t S = ctx->S;
    int rc = begin_op(ctx, s, OP_DEMUX1);
    if (rc) return rc;
    CK(cudaMemsetAsync(s->counts, 0, (uint64_t)(S + 2) * 8, s->stream));
    SheetDev sh;
    memset(&sh, 0, sizeof sh);
    sh.planes = ctx->d_planes;
    sh.umask = ctx->d_umask;
    sh.lut = ctx->d_lut;
    sh.S = S;
    sh.L = ctx->L;
    sh.Umax = ctx->Umax;
    sh.wide = ctx->wide;
    // --index1 / --index2: record tables of the index streams (line tables 0 and 1 are free until the mates take them)
    int idx_stream[2];
    uint32_t n_index = 0;
    if (o.use_index & 1u) idx_stream[n_index++] = SK_IN_AUX1;
    if (o.use_index & 2u) idx_stream[n_index++] = SK_IN_AUX2;
    const RecRef *ext_tab[2] = {nullptr, nullptr};
    const uint8_t *ext_data[2] = {nullptr, nullptr};
    const DevStats *ext_stats[2] = {nullptr, nullptr};
    for (uint32_t q = 0; q < n_index; q++) {
        const int w = idx_stream[q];
        const char *err = nullptr;
        const int n = launch_scan_table(s->in[w], s->in_len[w], 4u, 0u, o.rec_limit ? o.rec_limit : ~0ull, 1u, s->scan_tab[q], nullptr,
                                        ctx->lim.max_records, s->lwork, (int)q, ctx->lim.max_stream_bytes, ctx->lim.max_records,
                                        s->stats + w, ctx->sm_count, s->stream, &err);
        if (n < 0) {
            ctx->err = std::string("record table launch failed: ") + (err ? err : "?");
            return SK_E_CUDA;
        }
        s->launches += (uint32_t)n;
        s->pass_ran[w] = true;
        ext_tab[q] = s->scan_tab[q];
        ext_data[q] = s->in[w];
        ext_stats[q] = s->stats + w;
    }
    const int nm = s->paired ? 2 : 1;
    for (int mate = 0; mate < nm; mate++) {
        const int which = mate == 0 ? SK_IN_R1 : SK_IN_R2;
        const char *err = nullptr;
        uint32_t n_rows = 0;
        const int n = launch_line_demux(mate, s->in[which], s->in_len[which], o.rec_limit, o.fused_trim_min_baseq, sh, s->assign, s->umi,
                                        s->groups[mate], s->rows[mate], ctx->max_chunks, s->counts, s->events,
                                        (uint32_t)std::min<uint64_t>(ctx->lim.max_records, 0xFFFFFFFFull), s->stats + SK_IN_R1,
                                        o.no_output ? nullptr : s->out[mate], s->out_cap, s->lwork, ctx->lim.max_stream_bytes,
                                        ctx->lim.max_records, s->stats + which, ctx->sm_count, s->stream, &n_rows, &err, n_index,
                                        ext_tab, ext_data, ext_stats);
        if (n < 0) {
            ctx->err = std::string("line engine launch failed: ") + (err ? err : "?");
            return SK_E_CUDA;
        }
        s->n_chunks[which] = n_rows;
        s->launches += (uint32_t)n;
        s->pass_ran[which] = true;
    }
    s->ran_line = true;
    return end_op(ctx, s);
}

// Per-sample compaction of the last demultiplex of the slot (sk_compact.cu), both mates, on the slot's stream.
static int compact_enqueue(sk_ctx *ctx, Slot *s);
static int compact_enqueue(sk_ctx *ctx, Slot *s) {
    const uint32_t S = ctx->S;
    const int nm = s->paired ? 2 : 1;
    // (a mate that is absent or empty has empty slices, not those of an earlier batch)
    CK(cudaMemsetAsync(s->slices, 0, (uint64_t)(ctx->lim.max_samples + 1) * 16 * 2, s->stream));
    for (int m = 0; m < nm; m++) {
        const int which = m == 0 ? SK_IN_R1 : SK_IN_R2;
        const char *err = nullptr;
        const int rc = launch_compact(s->rows[m], s->groups[m], s->n_chunks[which], S, s->out[m], m == 0 ? s->cbuf : s->out[0],
                                      s->out_cap, s->cwork, s->slices + (size_t)m * (ctx->lim.max_samples + 1) * 2, s->piece_dst,
                                      s->stats + which, ctx->sm_count, s->stream, &err);
        if (rc < 0) {
            ctx->err = std::string("per-sample compaction: ") + (err ? err : "?");
            return SK_E_UNSUPPORTED;
        }
        s->launches += (uint32_t)rc;
    }
    s->compacted = true;
    return end_op(ctx, s);
}
extern "C" int sk_demux_compact(sk_ctx *ctx, uint32_t slot) {
    Slot *s = get_slot(ctx, slot);
    if (!s || s->last_op != OP_DEMUX1) return SK_E_INVALID;
    if (!s->cbuf) {
        ctx->err = "sk_demux_compact: the context holds no compaction buffers (more than 4096 samples, or switched off)";
        return SK_E_UNSUPPORTED;
    }
    if (s->req_opts.no_output) return SK_OK;  // dry run: nothing was written
    s->want_compact = true;
    return compact_enqueue(ctx, s);
}
extern "C" const void *sk_compact_dev(sk_ctx *ctx, uint32_t slot, uint32_t which) {
    Slot *s = get_slot(ctx, slot);
    return (s && which < 2 && s->cbuf) ? (which == 0 ? s->cbuf : s->out[0]) : nullptr;
}
extern "C" int sk_download_compact(sk_ctx *ctx, uint32_t slot, uint32_t which, void *host, uint64_t n) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= 2 || !s->cbuf || n > s->out_cap) return SK_E_INVALID;
    if (n) CK(cudaMemcpyAsync(host, which == 0 ? s->cbuf : s->out[0], n, cudaMemcpyDeviceToHost, s->stream));
    return SK_OK;
}
extern "C" int sk_download_slices(sk_ctx *ctx, uint32_t slot, uint32_t which, sk_slice *slices) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= 2 || !s->slices || !slices) return SK_E_INVALID;
    CK(cudaMemcpyAsync(slices, s->slices + (size_t)which * (ctx->lim.max_samples + 1) * 2, (uint64_t)(ctx->S + 1) * 16,
                       cudaMemcpyDeviceToHost, s->stream));
    return SK_OK;
}

extern "C" int sk_wait(sk_ctx *ctx, uint32_t slot, sk_result *res) {
    Slot *s = get_slot(ctx, slot);
    if (!s) return SK_E_INVALID;
    CK(cudaStreamSynchronize(s->stream));
    if (s->used_fast && s->last_op >= 0) {
        // The warp engine met something outside its limits (long record, dense tile, oversized
        // output): run the operator again on the general engine.
        unsigned fl = 0;
        for (int i = 0; i < SK_N_INPUTS; i++) fl |= s->stats_h[i].flags;
        if (!(fl & F_NEED_GENERAL) && (fl & F_NEED_ORDERED) && (s->last_op == OP_MASK || s->last_op == OP_ADDBC) && !s->no_inplace) {
            // mask met a record that changes its length (or fails), add barcode one that does not grow like the
            // first: same engine, ordered output
            const uint32_t first_launches = s->launches;
            s->no_inplace = true;
            int rc = s->last_op == OP_ADDBC ? addbc_enqueue(ctx, s, s->req_rec_limit, ctx->fast)
                                            : stream_op_enqueue(ctx, s, s->last_op, s->req_min_baseq, s->req_rec_limit, ctx->fast);
            if (rc) return rc;
            CK(cudaStreamSynchronize(s->stream));
            s->launches += first_launches;
            fl = 0;
            for (int i = 0; i < SK_N_INPUTS; i++) fl |= s->stats_h[i].flags;
        }
        if (fl & F_NEED_GENERAL) {
            const uint32_t fast_launches = s->launches;
            const sk_demux_opts o = s->req_opts;
            const bool again = s->want_compact;  // (begin_op clears it)
            int rc = s->last_op == OP_DEMUX1  ? demux_enqueue(ctx, s, &o, false)
                     : s->last_op == OP_ADDBC ? addbc_enqueue(ctx, s, s->req_rec_limit, false)
                                              : stream_op_enqueue(ctx, s, s->last_op, s->req_min_baseq, s->req_rec_limit, false);
            if (rc) return rc;
            if (again) {
                s->want_compact = true;
                rc = compact_enqueue(ctx, s);
                if (rc) return rc;
            }
            CK(cudaStreamSynchronize(s->stream));
            s->launches += fast_launches;
            s->reran_general = true;
        }
    }
    if ((s->last_op == OP_TRIM || s->last_op == OP_MASK) && !s->ran_line && s->lwork) {
        // Last resort for trim / mask by quality: what neither chunk engine takes -- a record longer than the general
        // engine's overhang, more records in a chunk than it has slots for, bytes >= 0x80 (UTF-8 in header or '+' lines
        // is data like any other; in bases or qualities the batch is still refused) -- runs on the line engine, which
        // frames records by a global line table and has no such limits.
        unsigned fl = 0;
        unsigned long long key = ~0ull;
        for (int i = 0; i < SK_N_INPUTS; i++) {
            fl |= s->stats_h[i].flags;
            if (s->stats_h[i].err_key) key = std::min(key, ~s->stats_h[i].err_key);
        }
        const unsigned kind = key == ~0ull ? 0u : (unsigned)(key & 0xFFu);
        if ((fl & F_NON_ASCII) || kind == K_TOO_LONG || kind == K_TOO_DENSE) {
            const uint32_t before = s->launches;
            CK(cudaMemsetAsync(s->stats, 0, sizeof(DevStats) * SK_N_INPUTS, s->stream));
            const char *err = nullptr;
            const int n = launch_lineop(s->last_op == OP_TRIM ? 6 : 7 /* LOP_TRIMQ / LOP_MASKQ */, s->in[SK_IN_R1], s->in_len[SK_IN_R1],
                                        nullptr, 0, 4, '@', s->req_min_baseq, 0, s->req_rec_limit, s->out[0], s->out[1], s->out_cap,
                                        s->lwork, ctx->lim.max_stream_bytes, ctx->lim.max_records, nullptr, 0, s->stats + SK_IN_R1,
                                        ctx->sm_count, s->stream, &err);
            if (n < 0) {
                ctx->err = std::string("line engine launch failed: ") + (err ? err : "?");
                return SK_E_CUDA;
            }
            s->launches = before + (uint32_t)n;
            s->ran_line = true;
            int rc = end_op(ctx, s);
            if (rc) return rc;
            CK(cudaStreamSynchronize(s->stream));
        }
    }
    if (s->last_op == OP_ADDBC && !s->ran_line && s->lwork) {  // and for add barcode
        unsigned fl = 0;
        unsigned long long key = ~0ull;
        for (int i = 0; i < SK_N_INPUTS; i++) {
            fl |= s->stats_h[i].flags;
            if (s->stats_h[i].err_key) key = std::min(key, ~s->stats_h[i].err_key);
        }
        const unsigned kind = key == ~0ull ? 0u : (unsigned)(key & 0xFFu);
        if ((fl & F_NON_ASCII) || kind == K_TOO_LONG || kind == K_TOO_DENSE) {
            const uint32_t before = s->launches;
            int rc = line_addbc_enqueue(ctx, s);
            if (rc) return rc;
            CK(cudaStreamSynchronize(s->stream));
            s->launches += before;
        }
    }
    if (s->last_op == OP_DEMUX1 && !s->ran_line && s->lwork) {
        // The same last resort for header-route demultiplex (sk_result.reserved bit 4).
        unsigned fl = 0;
        unsigned long long key = ~0ull;
        for (int i = 0; i < SK_N_INPUTS; i++) {
            fl |= s->stats_h[i].flags;
            if (s->stats_h[i].err_key) key = std::min(key, ~s->stats_h[i].err_key);
        }
        const unsigned kind = key == ~0ull ? 0u : (unsigned)(key & 0xFFu);
        if ((fl & F_NON_ASCII) || kind == K_TOO_LONG || kind == K_TOO_DENSE) {
            const uint32_t before = s->launches;
            const bool again = s->want_compact || s->compacted;
            int rc = line_demux_enqueue(ctx, s);
            if (rc) return rc;
            if (again) {
                s->want_compact = true;
                rc = compact_enqueue(ctx, s);
                if (rc) return rc;
            }
            CK(cudaStreamSynchronize(s->stream));
            s->launches += before;
        }
    }
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
    if (h[SK_IN_R1].n_records >= 64 && h[SK_IN_R1].consumed)  // record size of this data, for the next tile choice
        ctx->rec_est = (double)h[SK_IN_R1].consumed / (double)h[SK_IN_R1].n_records;
    res->gpu_launches = s->launches;
    res->reserved = (s->used_fast ? 1u : 0u) | (s->reran_general ? 2u : 0u) | (s->no_inplace ? 4u : 0u) | (s->compacted ? 8u : 0u) | (s->ran_line ? 16u : 0u);  // diagnostic: bit0 warp engine, bit1 re-run on the general engine, bit2 mask re-run in its ordered form
    if (ctx->profiling)
        for (int i = 0; i < SK_N_INPUTS; i++)
            if (s->pass_ran[i]) cudaEventElapsedTime(&res->pass_ms[i], s->ev[i][0], s->ev[i][1]);
    if (s->last_op == OP_DEMUX1) {
        res->out_bytes[0] = h[SK_IN_R1].out_bytes;
        res->out_extent[0] = h[SK_IN_R1].out_cursor;
        res->out_bytes[1] = h[SK_IN_R2].out_bytes;
        res->out_extent[1] = h[SK_IN_R2].out_cursor;
        if (s->compacted) {  // extent of the compacted buffers (sk_download_compact)
            res->out_extent[0] = h[SK_IN_R1].compact_extent;
            res->out_extent[1] = s->paired ? h[SK_IN_R2].compact_extent : 0;
        }
        res->n_chunks[0] = s->n_chunks[SK_IN_R1];
        res->n_chunks[1] = s->n_chunks[SK_IN_R2];
        res->n_events = std::min<uint32_t>(h[SK_IN_R1].n_events, (uint32_t)ctx->lim.max_records);
        if (s->paired && h[SK_IN_R2].n_records < h[SK_IN_R1].n_records) flags |= F_MATE_COUNT;
        for (int i = SK_IN_AUX1; i <= SK_IN_AUX2; i++)
            if (s->pass_ran[i] && h[i].n_records < h[SK_IN_R1].n_records) flags |= F_MATE_COUNT;
        // counters live on the device; total / identified came over with the outcome block (end_op)
        res->total_reads = s->ti_h[0];
        res->identified_reads = s->ti_h[1];
    } else {
        res->out_bytes[0] = h[SK_IN_R1].out_bytes;
        res->out_extent[0] = h[SK_IN_R1].out_extent;
        if (s->last_op == OP_LINE) {  // deinterleave: the second output stream
            res->out_bytes[1] = h[SK_IN_R1].compact_extent;
            res->out_extent[1] = h[SK_IN_R1].compact_extent;
        }
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
extern "C" int sk_download_demux_tables(sk_ctx *ctx, uint32_t slot, uint32_t which, sk_chunk_row *rows, sk_group *groups,
                                        uint64_t n_records) {
    Slot *s = get_slot(ctx, slot);
    if (!s || which >= 2 || !s->groups[which] || n_records > ctx->lim.max_records) return SK_E_INVALID;
    const uint32_t nc = s->n_chunks[which == 0 ? SK_IN_R1 : SK_IN_R2];
    if (nc) CK(cudaMemcpyAsync(rows, s->rows[which], (uint64_t)nc * sizeof(ChunkRow), cudaMemcpyDeviceToHost, s->stream));
    if (n_records)
        CK(cudaMemcpyAsync(groups, s->groups[which], n_records * sizeof(Group), cudaMemcpyDeviceToHost, s->stream));
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

extern "C" uint64_t sk_demux_gather(const uint8_t *out_host, const sk_chunk_row *rows, const sk_group *groups,
                                    uint32_t n_chunks, uint32_t smp, uint8_t *dst, uint64_t dst_cap) {
    uint64_t total = 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        uint64_t off = rows[c].base;
        const sk_group *g = groups + rows[c].first_group;
        for (uint32_t k = 0; k < rows[c].n_groups; k++) {
            if (g[k].sample == smp) {
                if (dst && total + g[k].len <= dst_cap) memcpy(dst + total, out_host + off, g[k].len);
                total += g[k].len;
            }
            off += g[k].len;
        }
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
// Run totals: the counters of a finished batch are added to the context's device-side totals on the slot's
// stream (call it once the batch's outcome is final, i.e. after sk_wait and any replay); at the end of a run
// the totals of all GPUs of the process are merged with ONE grouped ncclAllReduce (fasta_demultiplex.rs:263-264
// needs total / identified; the dry-run table needs the per-sample counts).
__global__ void sk_add_counts_kernel(unsigned long long *totals, const unsigned long long *counts, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) totals[i] += counts[i];
}
extern "C" int sk_counts_accumulate(sk_ctx *ctx, uint32_t slot) {
    Slot *s = get_slot(ctx, slot);
    if (!s || !s->counts || !ctx->d_totals) return SK_E_INVALID;
    const uint32_t n = ctx->S + 2;
    sk_add_counts_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(ctx->d_totals, s->counts, n);
    return cudaGetLastError() == cudaSuccess ? SK_OK : SK_E_CUDA;
}
extern "C" int sk_totals_reset(sk_ctx *ctx) {
    if (!ctx || !ctx->d_totals) return SK_E_INVALID;
    cudaSetDevice(ctx->device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(ctx->d_totals, 0, (uint64_t)(ctx->lim.max_samples + 2) * 8));
    return SK_OK;
}
extern "C" int sk_download_totals(sk_ctx *ctx, uint64_t *totals) {
    if (!ctx || !ctx->d_totals || !totals) return SK_E_INVALID;
    cudaSetDevice(ctx->device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(totals, ctx->d_totals, (uint64_t)(ctx->S + 2) * 8, cudaMemcpyDeviceToHost));
    return SK_OK;
}
extern "C" int sk_allreduce_totals(sk_ctx **ctxs, int n) {
    if (!ctxs || n < 1 || !ctxs[0]) return SK_E_INVALID;
    sk_ctx *ctx = ctxs[0];
    for (int i = 0; i < n; i++) {
        if (!ctxs[i] || !ctxs[i]->d_totals || ctxs[i]->S != ctx->S) return SK_E_INVALID;
        cudaSetDevice(ctxs[i]->device);
        CK(cudaDeviceSynchronize());
    }
    if (n == 1) return SK_OK;
    typedef int (*init_all_fn)(void **, int, const int *);
    typedef int (*group_fn)(void);
    typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    typedef int (*destroy_fn)(void *);
    init_all_fn init_all = (init_all_fn)nccl_sym(ctx, "ncclCommInitAll");
    group_fn gstart = (group_fn)nccl_sym(ctx, "ncclGroupStart"), gend = (group_fn)nccl_sym(ctx, "ncclGroupEnd");
    allreduce_fn allreduce = (allreduce_fn)nccl_sym(ctx, "ncclAllReduce");
    destroy_fn destroy = (destroy_fn)nccl_sym(ctx, "ncclCommDestroy");
    if (!init_all || !gstart || !gend || !allreduce || !destroy) return SK_E_UNSUPPORTED;
    std::vector<void *> comms((size_t)n, nullptr);
    std::vector<int> devs((size_t)n);
    for (int i = 0; i < n; i++) devs[(size_t)i] = ctxs[i]->device;
    int rc = init_all(comms.data(), n, devs.data());
    if (rc != 0) {
        ctx->err = "ncclCommInitAll failed with code " + std::to_string(rc);
        return SK_E_CUDA;
    }
    gstart();
    for (int i = 0; i < n && rc == 0; i++) {
        cudaSetDevice(ctxs[i]->device);
        // ncclUint64 = 5, ncclSum = 0; the context's first slot stream carries the collective
        rc = allreduce(ctxs[i]->d_totals, ctxs[i]->d_totals, (size_t)ctx->S + 2, 5, 0, comms[(size_t)i], ctxs[i]->slots[0].stream);
    }
    const int rc2 = gend();
    for (int i = 0; i < n; i++) {
        cudaSetDevice(ctxs[i]->device);
        cudaStreamSynchronize(ctxs[i]->slots[0].stream);
        destroy(comms[(size_t)i]);
    }
    if (rc != 0 || rc2 != 0) {
        ctx->err = "ncclAllReduce (grouped) failed with code " + std::to_string(rc ? rc : rc2);
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
