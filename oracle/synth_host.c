/* synth_host.c -- host-side twin of the device generator (seqkit_b200/csrc/sk_synth.cu).
 *
 * TEST / BENCH INFRASTRUCTURE ONLY, like everything under oracle/: `bench.py --impl reference` builds its
 * input with it, so that the CPU arm never loads the product's library, and tests/ check that it yields the
 * device generator's bytes.  Every read is a pure function of (seed, global pair index, mate): the same
 * counter-based generator, restated in plain C. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t seed, first_pair, n_pairs;
    uint32_t read_len, mate, with_bc, qual_profile, p_sub_ppm, p_n_ppm, p_random_ppm, reserved;
} synth_spec; /* == sk_synth_spec */

typedef struct { uint64_t s; } rng_t;
static uint64_t rng_next(rng_t *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static rng_t make_rng(uint64_t seed, uint64_t pair, uint64_t stream) {
    rng_t r = {seed ^ (pair * 0xD1342543DE82EF95ull) ^ (stream * 0xA24BAED4963EE407ull)};
    rng_next(&r);
    return r;
}
static uint8_t *put_dec(uint8_t *d, uint32_t v) {
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *d++ = (uint8_t)tmp[--n];
    return d;
}
/* one record at d; returns the end */
static uint8_t *record(uint8_t *d, const synth_spec *sp, uint64_t pair, const uint8_t *sheet, uint32_t S, uint32_t L) {
    rng_t rh = make_rng(sp->seed, pair, 0);
    const uint32_t lane = 1 + (uint32_t)(rng_next(&rh) % 8), tile = 1101 + (uint32_t)(rng_next(&rh) % 1578);
    const uint32_t x = 1000 + (uint32_t)(rng_next(&rh) % 29000), y = 1000 + (uint32_t)(rng_next(&rh) % 29000);
    memcpy(d, "@SIM:1:FC:", 10);
    d += 10;
    *d++ = (uint8_t)('0' + lane);
    *d++ = ':';
    d = put_dec(d, tile);
    *d++ = ':';
    d = put_dec(d, x);
    *d++ = ':';
    d = put_dec(d, y);
    *d++ = ' ';
    *d++ = (uint8_t)('0' + sp->mate);
    memcpy(d, ":N:0:1", 6);
    d += 6;
    if (sp->with_bc) {
        rng_t rb = make_rng(sp->seed, pair, 3);
        const uint64_t u = rng_next(&rb) & 0xFFFFFu;
        const uint32_t smp = (uint32_t)(((uint64_t)S * u * u) >> 40);
        const int all_random = (rng_next(&rb) % 1000000u) < sp->p_random_ppm;
        memcpy(d, " BC:", 4);
        d += 4;
        for (uint32_t q = 0; q < L; q++) {
            uint8_t c = sheet[(uint64_t)smp * L + q];
            const uint64_t r = rng_next(&rb);
            const uint8_t rnd = (uint8_t)"ACGT"[r & 3];
            if (c != '+') {
                if (c == 'U' || c == 'N' || all_random) c = rnd;
                if (((r >> 8) % 1000000u) < sp->p_sub_ppm) c = (uint8_t)"ACGT"[(r >> 2) & 3];
                if (((r >> 32) % 1000000u) < sp->p_n_ppm) c = 'N';
            }
            *d++ = c;
        }
    }
    *d++ = '\n';
    rng_t rq = make_rng(sp->seed, pair, sp->mate);
    const uint32_t n = sp->read_len;
    uint8_t *seq = d, *qual = d + n + 3;
    const uint64_t r0 = rng_next(&rq);
    const uint32_t sc = (uint32_t)(r0 & 3) == 0 ? 0 : (2u << ((r0 & 3) - 1));
    const uint32_t crash = ((r0 >> 8) % 100u) < 5u ? (uint32_t)((r0 >> 16) % (n ? n : 1)) : n;
    const uint64_t den = n > 1 ? (uint64_t)(n - 1) * (n - 1) * (n - 1) : 1;
    for (uint32_t k = 0; k < n; k++) {
        const uint64_t r = rng_next(&rq);
        seq[k] = (((r >> 2) % 1000u) == 0) ? (uint8_t)'N' : (uint8_t)"ACGT"[r & 3];
        int q = 38 - (int)((30ull * k * k * k) / den) + (sc ? (int)((r >> 16) % (2 * sc + 1)) - (int)sc : 0);
        q = q < 2 ? 2 : q > 41 ? 41 : q;
        if (k >= crash) q = 2;
        if (sp->qual_profile == 1) q = q < 7 ? 2 : q < 18 ? 11 : q < 31 ? 25 : 37;
        qual[k] = (uint8_t)(33 + q);
    }
    seq[n] = '\n';
    seq[n + 1] = '+';
    seq[n + 2] = '\n';
    qual[n] = '\n';
    return qual + n + 1;
}
/* Writes the FASTQ text of `sp` into dst (capacity cap); returns the bytes written, 0 when cap is too small.
 * Upper bound per record: 48 + 4 + L + 2 * read_len + 6. */
uint64_t orc_synth_fastq(uint8_t *dst, uint64_t cap, const synth_spec *sp, const uint8_t *sheet, uint32_t S, uint32_t L) {
    uint8_t *d = dst;
    const uint64_t worst = 64 + L + 2ull * sp->read_len;
    for (uint64_t i = 0; i < sp->n_pairs; i++) {
        if ((uint64_t)(d - dst) + worst > cap) return 0;
        d = record(d, sp, sp->first_pair + i, sheet, S, L);
    }
    return (uint64_t)(d - dst);
}
