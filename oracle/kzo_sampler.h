/* ORACLE (test infrastructure) -- samplers restated from the reference:
 *   hash.h:15-108 (MurmurHash64A, MixBits, Hash), pcg32.h:41-166, common.cpp:316-344 (permute),
 *   sampler.cpp:18-71 (independent), :81-156 (stratified), :176-269 (correlated), :273-390 (pmj02bn),
 *   bluenoise.h:16-23, pmj02table.h:17-29. */
#ifndef KZO_SAMPLER_H
#define KZO_SAMPLER_H
#include "kzo_math.h"
#include "../include/kzgpu.h"
#include <vector>

namespace kzo {

/* hash.h:15-62 */
inline uint64_t murmur64a(const unsigned char *key, size_t len, uint64_t seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    const int r = 47;
    uint64_t h = seed ^ (len * m);
    const unsigned char *end = key + 8 * (len / 8);
    while (key != end) {
        uint64_t k;
        std::memcpy(&k, key, 8);
        key += 8;
        k *= m; k ^= k >> r; k *= m;
        h ^= k; h *= m;
    }
    switch (len & 7) {
        case 7: h ^= uint64_t(key[6]) << 48; /* fallthrough */
        case 6: h ^= uint64_t(key[5]) << 40; /* fallthrough */
        case 5: h ^= uint64_t(key[4]) << 32; /* fallthrough */
        case 4: h ^= uint64_t(key[3]) << 24; /* fallthrough */
        case 3: h ^= uint64_t(key[2]) << 16; /* fallthrough */
        case 2: h ^= uint64_t(key[1]) << 8;  /* fallthrough */
        case 1: h ^= uint64_t(key[0]); h *= m;
    }
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}
/* hash.h:66-73 */
inline uint64_t mixBits(uint64_t v) {
    v ^= (v >> 31); v *= 0x7fb5d329728ea185ull;
    v ^= (v >> 27); v *= 0x81dadef4bc2dd44dull;
    v ^= (v >> 33);
    return v;
}
/* Hash(Point2i p, uint64_t seed): 16 bytes, no padding (hash.h:92-108, sampler.cpp:44,115) */
inline uint64_t hashPixelSeed(int32_t x, int32_t y, uint64_t seed) {
    unsigned char buf[16];
    std::memcpy(buf, &x, 4); std::memcpy(buf + 4, &y, 4); std::memcpy(buf + 8, &seed, 8);
    return murmur64a(buf, 16, 0);
}
/* Hash(Point2i p, uint32_t dim, uint64_t seed): 20 bytes (sampler.cpp:121,132) */
inline uint64_t hashPixelDimSeed(int32_t x, int32_t y, uint32_t dim, uint64_t seed) {
    unsigned char buf[24];
    std::memcpy(buf, &x, 4); std::memcpy(buf + 4, &y, 4); std::memcpy(buf + 8, &dim, 4); std::memcpy(buf + 12, &seed, 8);
    return murmur64a(buf, 20, 0);
}

/* pcg32.h:41-166 */
struct Pcg32 {
    uint64_t state, inc;
    static constexpr uint64_t kMult = 0x5851f42d4c957f2dULL;
    uint32_t nextUInt() {
        uint64_t old = state;
        state = old * kMult + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    void seed(uint64_t initstate, uint64_t initseq) {
        state = 0U;
        inc = (initseq << 1u) | 1u;
        nextUInt();
        state += initstate;
        nextUInt();
    }
    void seed(uint64_t initseq) { seed(mixBits(initseq), initseq); }
    float nextFloat() {
        uint32_t u = (nextUInt() >> 9) | 0x3f800000u;
        float f;
        std::memcpy(&f, &u, 4);
        return f - 1.0f;
    }
    void advance(uint64_t delta) {
        uint64_t cur_mult = kMult, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
        while (delta > 0) {
            if (delta & 1) {
                acc_mult *= cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta /= 2;
        }
        state = acc_mult * state + acc_plus;
    }
};

/* common.cpp:316-344 (Kensler permutation; note uint32_t p: callers truncate 64-bit hashes) */
inline uint32_t permute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p;             i *= 0xe170893d;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;        i *= 0x0929eb3f;
        i ^= p >> 23;
        i ^= (i & w) >> 1;  i *= 1 | p >> 27;
        i *= 0x6935fa69;
        i ^= (i & w) >> 11; i *= 0x74dcb303;
        i ^= (i & w) >> 2;  i *= 0x9e501cc3;
        i ^= (i & w) >> 2;  i *= 0xc860a3df;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}

/* pmj02bn per-pixel sample buckets, sampler.cpp:275-315 */
struct PmjPixelSamples {
    int tileSize = 0;
    std::vector<V2> samples;
};

inline int log2i_int(int v) { int r = 0; while (v > 1) { v >>= 1; ++r; } return r; }
inline bool isPow4(int n) { /* common.h:271-289 */
    if (n <= 0) return false;
    int x = (int)std::sqrt((double)n);
    if (x * x != n) return false;
    return !(n & (n - 1));
}
inline int roundUpPow4(int v) { return isPow4(v) ? v : (1 << (2 * (1 + log2i_int(v) / 2))); } /* common.h:318-320 */

struct SamplerCfg {
    kz_sampler_desc d;
    PmjPixelSamples pmjPix;
};

inline V2 pmjSample(const SamplerCfg &c, int set, int idx) {   /* pmj02table.h:17-29 */
    set %= 5; idx %= 65536;
    const uint32_t *t = c.d.pmj02bn + ((size_t)set * 65536 + idx) * 2;
    return V2{(float)(t[0] * 0x1p-32), (float)(t[1] * 0x1p-32)};
}
inline float blueNoise(const SamplerCfg &c, int tex, int px, int py) {   /* bluenoise.h:16-23 */
    tex %= 48;
    int x = px % 128, y = py % 128;
    return c.d.blue_noise[((size_t)tex * 128 + x) * 128 + y] / 65535.f;
}

inline void buildPmjPixelSamples(SamplerCfg &c) {   /* sampler.cpp:290-314 */
    int spp = (int)c.d.sample_count;
    int tile = 1 << (log2i_int(65536) / 2 - log2i_int(roundUpPow4(spp)) / 2);
    c.pmjPix.tileSize = tile;
    c.pmjPix.samples.assign((size_t)tile * tile * spp, V2{0, 0});
    std::vector<int> nStored((size_t)tile * tile, 0);
    for (int i = 0; i < 65536; ++i) {
        V2 p = pmjSample(c, 0, i);
        p.x *= tile; p.y *= tile;
        int pixelOffset = int(p.x) + int(p.y) * tile;
        if (nStored[pixelOffset] == spp) continue;
        int sampleOffset = pixelOffset * spp + nStored[pixelOffset];
        c.pmjPix.samples[sampleOffset] = V2{p.x - std::floor(p.x), p.y - std::floor(p.y)};
        ++nStored[pixelOffset];
    }
}

/* One sampler instance = the mutable per-thread clone of renderer.cpp:102 */
struct Sampler {
    const SamplerCfg *cfg;
    Pcg32 rng;
    int32_t px, py;
    uint32_t sampleIndex, dim;
    /* probe mode (kzo_light_sample_dump): next1D() replays these numbers instead of drawing */
    const float *fixed = nullptr; int fixedPos = 0;

    void generateSample(int32_t x, int32_t y, int sampleIdx, int dimension = 0) {
        const kz_sampler_desc &d = cfg->d;
        px = x; py = y; sampleIndex = (uint32_t)sampleIdx;
        if (d.type == KZ_SAMPLER_PMJ02BN) {           /* sampler.cpp:333-337 */
            dim = (uint32_t)std::max(2, dimension);
            return;
        }
        dim = (uint32_t)dimension;                    /* sampler.cpp:43-46,111-117,207-213 */
        rng.seed(hashPixelSeed(x, y, d.seed));
        rng.advance((uint64_t)sampleIdx * 65536ull + (uint64_t)dimension);
    }
    float next1D() {
        if (fixed) return fixed[fixedPos++];
        const kz_sampler_desc &d = cfg->d;
        switch (d.type) {
            case KZ_SAMPLER_INDEPENDENT: return rng.nextFloat();           /* sampler.cpp:48-50 */
            case KZ_SAMPLER_STRATIFIED: {                                   /* sampler.cpp:119-127 */
                uint64_t h = hashPixelDimSeed(px, py, dim, d.seed);
                int stratum = (int)permute(sampleIndex, d.sample_count, (uint32_t)h);
                ++dim;
                float delta = rng.nextFloat();
                return (stratum + delta) / d.sample_count;
            }
            case KZ_SAMPLER_CORRELATED: {                                   /* sampler.cpp:215-227 */
                uint64_t h = hashPixelDimSeed(px, py, dim, d.seed);
                int p = (int)permute(sampleIndex, d.sample_count, (uint32_t)(h * 0x45fbe943));
                float j = rng.nextFloat();
                ++dim;
                return (p + j) / d.sample_count;
            }
            default: {                                                      /* sampler.cpp:339-347 */
                uint64_t h = hashPixelDimSeed(px, py, dim, d.seed);
                int index = (int)permute(sampleIndex, d.sample_count, (uint32_t)h);
                float delta = blueNoise(*cfg, (int)dim, px, py);
                ++dim;
                return std::min((index + delta) / d.sample_count, kOneMinusEpsilon);
            }
        }
    }
    V2 next2D() {
        const kz_sampler_desc &d = cfg->d;
        switch (d.type) {
            case KZ_SAMPLER_INDEPENDENT: {                                  /* sampler.cpp:52-57 */
                /* `Point2f(m_random.nextFloat(), m_random.nextFloat())`: the two calls are indeterminately sequenced constructor
                 * arguments; GCC (the reference's toolchain) evaluates them right to left, so x receives the SECOND draw
                 * (observed by running the reference's own body, oracle/ref_math_kat.cpp).  Stratified / Correlated return a
                 * braced list, which is left to right by the standard. */
                float second = rng.nextFloat();
                float first = rng.nextFloat();
                return V2{first, second};
            }
            case KZ_SAMPLER_STRATIFIED: {                                   /* sampler.cpp:129-139 */
                uint64_t h = hashPixelDimSeed(px, py, dim, d.seed);
                int stratum = (int)permute(sampleIndex, d.sample_count, (uint32_t)h);
                dim += 2;
                int x = stratum % d.res_x, y = stratum / d.res_x;
                float dx = rng.nextFloat();
                float dy = rng.nextFloat();
                return V2{(x + dx) / d.res_x, (y + dy) / d.res_x};
            }
            case KZ_SAMPLER_CORRELATED: {                                   /* sampler.cpp:229-251 */
                uint64_t h = hashPixelDimSeed(px, py, dim, d.seed);
                int s = (int)permute(sampleIndex, d.sample_count, (uint32_t)(h * 0x51633e2d));
                uint32_t y = (uint32_t)s / (uint32_t)d.res_x;
                uint32_t x = (uint32_t)s % (uint32_t)d.res_x;
                uint32_t sx = permute(x, (uint32_t)d.res_x, (uint32_t)(h * 0x68bc21eb));
                uint32_t sy = permute(y, (uint32_t)d.res_y, (uint32_t)(h * 0x02e5be93));
                float jx = rng.nextFloat();
                float jy = rng.nextFloat();
                dim += 2;
                return V2{(x + (sy + jx) / d.res_y) / d.res_x, (y + (sx + jy) / d.res_x) / d.res_y};
            }
            default: {                                                      /* sampler.cpp:349-371 */
                int index = (int)sampleIndex;
                int pmjInstance = (int)(dim / 2);
                if (pmjInstance >= 5) {
                    uint64_t h = hashPixelDimSeed(px, py, dim, d.seed);
                    index = (int)permute(sampleIndex, d.sample_count, (uint32_t)h);
                }
                V2 u = pmjSample(*cfg, pmjInstance, index);
                u.x += blueNoise(*cfg, (int)dim, px, py);
                u.y += blueNoise(*cfg, (int)dim + 1, px, py);
                if (u.x >= 1) u.x -= 1;
                if (u.y >= 1) u.y -= 1;
                dim += 2;
                return V2{std::min(u.x, kOneMinusEpsilon), std::min(u.y, kOneMinusEpsilon)};
            }
        }
    }
    V2 nextPixel2D() {
        if (cfg->d.type != KZ_SAMPLER_PMJ02BN) return next2D();          /* sampler.cpp:59,141,253 */
        int T = cfg->pmjPix.tileSize;                                     /* sampler.cpp:373-377 */
        int x = px % T, y = py % T;
        int offset = (x + y * T) * (int)cfg->d.sample_count;
        return cfg->pmjPix.samples[offset + sampleIndex];
    }
};

}  // namespace kzo
#endif
