/* kz_sampler.h -- device samplers, integer core bit-exact with the reference
 * (hash.h:15-108, pcg32.h:41-166, common.cpp:316-344, sampler.cpp:18-390).
 * State carried per path: pixel, sampleIndex, dimension, pcg32 (state, inc). */
#ifndef KZ_SAMPLER_H
#define KZ_SAMPLER_H
#include "kz_scene.h"

#define KZ_MURMUR_M 0xc6a4a7935bd1e995ull

KZ_HD uint64_t kz_murmur_block(uint64_t h, uint64_t k) {
    k *= KZ_MURMUR_M; k ^= k >> 47; k *= KZ_MURMUR_M;
    h ^= k; h *= KZ_MURMUR_M;
    return h;
}
KZ_HD uint64_t kz_murmur_final(uint64_t h) {
    h ^= h >> 47; h *= KZ_MURMUR_M; h ^= h >> 47;
    return h;
}
/* Hash(Point2i, uint64 seed): 16 bytes = two little-endian 64-bit blocks */
KZ_HD uint64_t kz_hash_pixel_seed(int32_t x, int32_t y, uint64_t seed) {
    uint64_t h = 0ull ^ (16ull * KZ_MURMUR_M);
    uint64_t k0 = (uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)y << 32);
    h = kz_murmur_block(h, k0);
    h = kz_murmur_block(h, seed);
    return kz_murmur_final(h);
}
/* Hash(Point2i, uint32 dim, uint64 seed): 20 bytes = blocks (x|y), (dim|seed.lo), tail 4 bytes seed.hi */
KZ_HD uint64_t kz_hash_pixel_dim_seed(int32_t x, int32_t y, uint32_t dim, uint64_t seed) {
    uint64_t h = 0ull ^ (20ull * KZ_MURMUR_M);
    uint64_t k0 = (uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)y << 32);
    uint64_t k1 = (uint64_t)dim | ((seed & 0xffffffffull) << 32);
    h = kz_murmur_block(h, k0);
    h = kz_murmur_block(h, k1);
    h ^= (seed >> 32);          /* tail bytes 0..3 (len & 7 == 4) */
    h *= KZ_MURMUR_M;
    return kz_murmur_final(h);
}
KZ_HD uint64_t kz_mix_bits(uint64_t v) {
    v ^= (v >> 31); v *= 0x7fb5d329728ea185ull;
    v ^= (v >> 27); v *= 0x81dadef4bc2dd44dull;
    v ^= (v >> 33);
    return v;
}
/* n / d and n % d; d is a per-scene constant and almost always a power of two (16, 64, 256 spp) */
KZ_HD void kz_divmod(uint32_t n, uint32_t d, uint32_t &q, uint32_t &r) {
    if ((d & (d - 1u)) == 0u) { q = n >> kz_bfind(d); r = n & (d - 1u); }
    else { q = n / d; r = n - q * d; }
}
KZ_HD uint32_t kz_permute(uint32_t i, uint32_t l, uint32_t p) {
    /* w = l-1 with every bit below its top bit set (common.cpp:318-323 builds it with five shift-ors) */
    const uint32_t w = l > 1u ? 0xFFFFFFFFu >> (31u - kz_bfind(l - 1u)) : 0u;
    do {
        i ^= p;             i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;        i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1;  i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2;  i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;  i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    const uint32_t r = i + p;
    return ((l & (l - 1u)) == 0u) ? (r & (l - 1u)) : r % l;
}

#define KZ_PCG_MULT 0x5851f42d4c957f2dULL

#if defined(__CUDACC__)
#define KZ_CONSTEXPR_HD __host__ __device__ constexpr
#else
#define KZ_CONSTEXPR_HD constexpr
#endif
/* cur_mult and cur_plus / inc of pcg32::advance after 16 rounds with no bit set in delta */
KZ_CONSTEXPR_HD uint64_t kz_pcg_mult_2_16() { uint64_t m = KZ_PCG_MULT; for (int i = 0; i < 16; ++i) m *= m; return m; }
KZ_CONSTEXPR_HD uint64_t kz_pcg_plus_2_16() { uint64_t m = KZ_PCG_MULT, g = 1u; for (int i = 0; i < 16; ++i) { g = (m + 1u) * g; m *= m; } return g; }

struct KzSampler {
    uint64_t state, inc;
    int32_t  px, py;
    uint32_t sample_index, dim;
};

KZ_HD uint32_t kz_pcg_next(KzSampler &s) {
    uint64_t old = s.state;
    s.state = old * KZ_PCG_MULT + s.inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
KZ_HD float kz_pcg_float(KzSampler &s) {
    return kz_sub(kz_u2f((kz_pcg_next(s) >> 9) | 0x3f800000u), 1.0f);
}
KZ_HD void kz_pcg_seed_advance(KzSampler &s, uint64_t initseq, uint64_t delta) {
    /* seed(MixBits(initseq), initseq) */
    s.state = 0u;
    s.inc = (initseq << 1u) | 1u;
    kz_pcg_next(s);
    s.state += kz_mix_bits(initseq);
    kz_pcg_next(s);
    /* advance(delta), pcg32.h:145-166.  generateSample always jumps by sampleIndex * 65536, so the first 16 rounds of the
     * loop only square the multiplier: they are folded into two compile-time constants (same arithmetic mod 2^64). */
    uint64_t cur_mult = KZ_PCG_MULT, cur_plus = s.inc, acc_mult = 1u, acc_plus = 0u;
    if ((delta & 0xFFFFull) == 0ull) {
        constexpr uint64_t m16 = kz_pcg_mult_2_16(), g16 = kz_pcg_plus_2_16();      /* evaluated by the compiler */
        cur_mult = m16; cur_plus = g16 * s.inc; delta >>= 16;
    }
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    s.state = acc_mult * s.state + acc_plus;
}

KZ_HD void kz_sampler_start(const KzScene &sc, KzSampler &s, int32_t px, int32_t py, uint32_t sample_index) {
    s.px = px; s.py = py; s.sample_index = sample_index;
    if (sc.sampler_type == KZ_SAMPLER_PMJ02BN) { s.dim = 2; s.state = 0; s.inc = 0; return; }   /* max(2, 0) */
    s.dim = 0;
    kz_pcg_seed_advance(s, kz_hash_pixel_seed(px, py, sc.seed), (uint64_t)sample_index * 65536ull);
}

KZ_HD float kz_blue_noise(const KzScene &sc, int tex, int px, int py) {
    tex %= 48;
    int x = px % 128, y = py % 128;
    return kz_div((float)sc.blue_noise[((size_t)tex * 128 + x) * 128 + y], 65535.f);
}
KZ_HD kz2 kz_pmj_sample(const KzScene &sc, int set, int idx) {
    set %= 5; idx %= 65536;
    const uint32_t *t = sc.pmj02bn + ((size_t)set * 65536 + idx) * 2;
    return mk2((float)(t[0] * 0x1p-32), (float)(t[1] * 0x1p-32));
}

KZ_HD float kz_next1d(const KzScene &sc, KzSampler &s) {
    switch (sc.sampler_type) {
        case KZ_SAMPLER_INDEPENDENT: return kz_pcg_float(s);
        case KZ_SAMPLER_STRATIFIED: {
            uint64_t h = kz_hash_pixel_dim_seed(s.px, s.py, s.dim, sc.seed);
            int stratum = (int)kz_permute(s.sample_index, sc.sample_count, (uint32_t)h);
            ++s.dim;
            float delta = kz_pcg_float(s);
            return kz_div(kz_add((float)stratum, delta), (float)sc.sample_count);
        }
        case KZ_SAMPLER_CORRELATED: {
            uint64_t h = kz_hash_pixel_dim_seed(s.px, s.py, s.dim, sc.seed);
            int p = (int)kz_permute(s.sample_index, sc.sample_count, (uint32_t)(h * 0x45fbe943ull));
            float j = kz_pcg_float(s);
            ++s.dim;
            return kz_div(kz_add((float)p, j), (float)sc.sample_count);
        }
        default: {
            uint64_t h = kz_hash_pixel_dim_seed(s.px, s.py, s.dim, sc.seed);
            int index = (int)kz_permute(s.sample_index, sc.sample_count, (uint32_t)h);
            float delta = kz_blue_noise(sc, (int)s.dim, s.px, s.py);
            ++s.dim;
            return fminf(kz_div(kz_add((float)index, delta), (float)sc.sample_count), KZ_ONE_MINUS_EPS);
        }
    }
}

KZ_HD kz2 kz_next2d(const KzScene &sc, KzSampler &s) {
    switch (sc.sampler_type) {
        case KZ_SAMPLER_INDEPENDENT: {      /* sampler.cpp:52-57 as GCC evaluates it: x is the second draw (see oracle/kzo_sampler.h) */
            float second = kz_pcg_float(s);
            float first = kz_pcg_float(s);
            return mk2(first, second);
        }
        case KZ_SAMPLER_STRATIFIED: {
            uint64_t h = kz_hash_pixel_dim_seed(s.px, s.py, s.dim, sc.seed);
            int stratum = (int)kz_permute(s.sample_index, sc.sample_count, (uint32_t)h);
            s.dim += 2;
            uint32_t x, y;
            kz_divmod((uint32_t)stratum, (uint32_t)sc.res_x, y, x);
            float dx = kz_pcg_float(s);
            float dy = kz_pcg_float(s);
            return mk2(kz_div(kz_add((float)x, dx), (float)sc.res_x), kz_div(kz_add((float)y, dy), (float)sc.res_x));
        }
        case KZ_SAMPLER_CORRELATED: {
            uint64_t h = kz_hash_pixel_dim_seed(s.px, s.py, s.dim, sc.seed);
            uint32_t sidx = kz_permute(s.sample_index, sc.sample_count, (uint32_t)(h * 0x51633e2dull));
            uint32_t x, y;
            kz_divmod(sidx, (uint32_t)sc.res_x, y, x);
            uint32_t sx = kz_permute(x, (uint32_t)sc.res_x, (uint32_t)(h * 0x68bc21ebull));
            uint32_t sy = kz_permute(y, (uint32_t)sc.res_y, (uint32_t)(h * 0x02e5be93ull));
            float jx = kz_pcg_float(s);
            float jy = kz_pcg_float(s);
            s.dim += 2;
            float fx = kz_div(kz_add((float)x, kz_div(kz_add((float)sy, jx), (float)sc.res_y)), (float)sc.res_x);
            float fy = kz_div(kz_add((float)y, kz_div(kz_add((float)sx, jy), (float)sc.res_x)), (float)sc.res_y);
            return mk2(fx, fy);
        }
        default: {
            int index = (int)s.sample_index;
            int inst = (int)(s.dim / 2);
            if (inst >= 5) {
                uint64_t h = kz_hash_pixel_dim_seed(s.px, s.py, s.dim, sc.seed);
                index = (int)kz_permute(s.sample_index, sc.sample_count, (uint32_t)h);
            }
            kz2 u = kz_pmj_sample(sc, inst, index);
            u.x = kz_add(u.x, kz_blue_noise(sc, (int)s.dim, s.px, s.py));
            u.y = kz_add(u.y, kz_blue_noise(sc, (int)s.dim + 1, s.px, s.py));
            if (u.x >= 1) u.x = kz_sub(u.x, 1.f);
            if (u.y >= 1) u.y = kz_sub(u.y, 1.f);
            s.dim += 2;
            return mk2(fminf(u.x, KZ_ONE_MINUS_EPS), fminf(u.y, KZ_ONE_MINUS_EPS));
        }
    }
}

KZ_HD kz2 kz_next_pixel2d(const KzScene &sc, KzSampler &s) {
    if (sc.sampler_type != KZ_SAMPLER_PMJ02BN) return kz_next2d(sc, s);
    int T = sc.pmj_tile_size;
    int x = s.px % T, y = s.py % T;
    int offset = (x + y * T) * (int)sc.sample_count;
    return sc.pmj_pixel_samples[offset + s.sample_index];
}

#endif
