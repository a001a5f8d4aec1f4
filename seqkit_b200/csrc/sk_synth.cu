// sk_synth.cu -- counter-based synthetic FASTQ generator on the device.
// Bench/test utility, not part of the hot path: every read is a pure function of
// (seed, global pair index, mate), so any shard can be regenerated on any GPU, and the host gets the
// identical bytes with a D2H copy (sk_download_in) to feed the CPU oracle.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_scan.cuh>

#include "../../include/seqkit_b200.h"
#include "sk_internal.h"

namespace sk {

struct Rng {
    uint64_t s;
    __device__ __forceinline__ uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};
__device__ __forceinline__ Rng make_rng(uint64_t seed, uint64_t pair, uint64_t stream) {
    Rng r{seed ^ (pair * 0xD1342543DE82EF95ull) ^ (stream * 0xA24BAED4963EE407ull)};
    r.next();
    return r;
}
__device__ __forceinline__ uint32_t ndigits(uint32_t v) {
    return v >= 10000 ? 5 : v >= 1000 ? 4 : v >= 100 ? 3 : v >= 10 ? 2 : 1;
}
struct Hdr {
    uint32_t lane, tile, x, y;
};
__device__ __forceinline__ Hdr make_hdr(uint64_t seed, uint64_t pair) {
    Rng r = make_rng(seed, pair, 0);
    Hdr h;
    h.lane = 1 + (uint32_t)(r.next() % 8);
    h.tile = 1101 + (uint32_t)(r.next() % 1578);
    h.x = 1000 + (uint32_t)(r.next() % 29000);
    h.y = 1000 + (uint32_t)(r.next() % 29000);
    return h;
}
// "@SIM:1:FC:<lane>:<tile>:<x>:<y> <mate>:N:0:1"
__device__ __forceinline__ uint32_t hdr_len(const Hdr &h) { return 10 + 1 + 1 + 4 + 1 + ndigits(h.x) + 1 + ndigits(h.y) + 1 + 1 + 6; }

__global__ void synth_len_kernel(sk_synth_spec sp, uint32_t L, uint64_t *lens) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n_pairs) return;
    const Hdr h = make_hdr(sp.seed, sp.first_pair + i);
    lens[i] = hdr_len(h) + (sp.with_bc ? 4 + L : 0) + 2ull * sp.read_len + 5;
}

__device__ __forceinline__ uint8_t *put_dec(uint8_t *d, uint32_t v) {
    const uint32_t n = ndigits(v);
    for (uint32_t k = 0; k < n; k++) {
        d[n - 1 - k] = (uint8_t)('0' + v % 10);
        v /= 10;
    }
    return d + n;
}

__global__ void synth_write_kernel(sk_synth_spec sp, const uint8_t *sheet, uint32_t S, uint32_t L,
                                   const uint64_t *offs, uint8_t *dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n_pairs) return;
    const uint64_t pair = sp.first_pair + i;
    uint8_t *d = dst + offs[i];
    const Hdr h = make_hdr(sp.seed, pair);
    const char *pre = "@SIM:1:FC:";
    for (int k = 0; k < 10; k++) *d++ = (uint8_t)pre[k];
    *d++ = (uint8_t)('0' + h.lane);
    *d++ = ':';
    d = put_dec(d, h.tile);
    *d++ = ':';
    d = put_dec(d, h.x);
    *d++ = ':';
    d = put_dec(d, h.y);
    *d++ = ' ';
    *d++ = (uint8_t)('0' + sp.mate);
    const char *suf = ":N:0:1";
    for (int k = 0; k < 6; k++) *d++ = (uint8_t)suf[k];
    if (sp.with_bc) {
        // observed barcode: the same for both mates of a pair
        Rng rb = make_rng(sp.seed, pair, 3);
        const uint64_t u = rb.next() & 0xFFFFFu;
        const uint32_t smp = (uint32_t)(((uint64_t)S * u * u) >> 40);  // skewed sample abundance
        const bool all_random = (rb.next() % 1000000u) < sp.p_random_ppm;
        *d++ = ' ';
        *d++ = 'B';
        *d++ = 'C';
        *d++ = ':';
        for (uint32_t q = 0; q < L; q++) {
            uint8_t c = sheet[(uint64_t)smp * L + q];
            const uint64_t r = rb.next();
            const uint8_t rnd = (uint8_t)"ACGT"[r & 3];
            if (c != '+') {
                if (c == 'U' || c == 'N' || all_random) c = rnd;
                if (((r >> 8) % 1000000u) < sp.p_sub_ppm) c = (uint8_t)"ACGT"[(r >> 2) & 3];
                if (((r >> 32) % 1000000u) < sp.p_n_ppm) c = 'N';
            }
            *d++ = c;
        }
    }
    *d++ = '\n';
    Rng rq = make_rng(sp.seed, pair, sp.mate);
    const uint32_t n = sp.read_len;
    uint8_t *seq = d, *qual = d + n + 3;
    const uint64_t r0 = rq.next();
    const uint32_t sc = (uint32_t)(r0 & 3) == 0 ? 0 : (2u << ((r0 & 3) - 1));  // noise scale 0,2,4,8
    const uint32_t crash = ((r0 >> 8) % 100u) < 5u ? (uint32_t)((r0 >> 16) % (n ? n : 1)) : n;
    const uint64_t den = n > 1 ? (uint64_t)(n - 1) * (n - 1) * (n - 1) : 1;
    for (uint32_t k = 0; k < n; k++) {
        const uint64_t r = rq.next();
        seq[k] = (((r >> 2) % 1000u) == 0) ? (uint8_t)'N' : (uint8_t)"ACGT"[r & 3];
        int q = 38 - (int)((30ull * k * k * k) / den) + (sc ? (int)((r >> 16) % (2 * sc + 1)) - (int)sc : 0);
        q = q < 2 ? 2 : q > 41 ? 41 : q;
        if (k >= crash) q = 2;
        if (sp.qual_profile == 1) q = q < 7 ? 2 : q < 18 ? 11 : q < 31 ? 25 : 37;  // RTA3 bins
        qual[k] = (uint8_t)(33 + q);
    }
    seq[n] = '\n';
    seq[n + 1] = '+';
    seq[n + 2] = '\n';
    qual[n] = '\n';
}

int launch_synth(void *stream_, uint8_t *dst, uint64_t cap, const sk_synth_spec &spec, const uint8_t *sheet_raw,
                 uint32_t S, uint32_t L, uint64_t *tmp, uint64_t *n_out, const char **err) {
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    if (spec.n_pairs == 0) return SK_OK;
    if (spec.mate < 1 || spec.mate > 2 || spec.read_len == 0 || spec.read_len > 4000) {
        *err = "bad spec";
        return SK_E_INVALID;
    }
    const uint32_t nb = (uint32_t)((spec.n_pairs + 255) / 256);
    synth_len_kernel<<<nb, 256, 0, stream>>>(spec, L, tmp);
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scratch_bytes, tmp, tmp, spec.n_pairs + 1, stream);
    if (cudaMalloc(&scratch, scratch_bytes) != cudaSuccess) {
        *err = "cudaMalloc(scan scratch)";
        return SK_E_NOMEM;
    }
    // tmp[n_pairs] is scanned too so that it ends up holding the total
    cudaMemsetAsync(tmp + spec.n_pairs, 0, 8, stream);
    cub::DeviceScan::ExclusiveSum(scratch, scratch_bytes, tmp, tmp, spec.n_pairs + 1, stream);
    uint64_t total = 0;
    cudaMemcpyAsync(&total, tmp + spec.n_pairs, 8, cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(scratch);
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return SK_E_CUDA;
    }
    if (total > cap) {
        *err = "synthetic batch larger than the slot's input capacity";
        return SK_E_TOO_LARGE;
    }
    synth_write_kernel<<<nb, 256, 0, stream>>>(spec, sheet_raw, S, L, tmp, dst);
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return SK_E_CUDA;
    }
    *n_out = total;
    return SK_OK;
}

}  // namespace sk
