/* ORACLE tooling -- known-answer-vector generator built from the REFERENCE'S OWN sources where
 * they lie (never copied into this repo): include/kazen/hash.h, include/kazen/pcg32.h are
 * #included from /root/reference/include, and the body of random::permute is extracted at build
 * time by oracle/Makefile from /root/reference/src/kazen/common.cpp:316-344 into
 * oracle/_ref/permute_extract.inc (git-ignored).  Output: JSON on stdout ->
 * tests/golden/sampler_kat.json (committed).  Runs only in the build container. */
#include <cstdint>
#include <cstdio>
#include <cstring>
#define NAMESPACE_BEGIN(n) namespace n {
#define NAMESPACE_END(n) }
#include <kazen/hash.h>
#include <kazen/pcg32.h>
namespace refx {
#include "_ref/permute_extract.inc"
}
/* Point2i is Eigen::Matrix<int,2,1>: two packed int32 (x then y); a POD stand-in hashes the same bytes. */
struct P2i { int32_t x, y; };

static uint64_t lcg(uint64_t &s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return s >> 11; }

int main() {
    uint64_t rs = 0x1234567ULL;
    printf("{\n");
    printf("\"hash16\": [");
    for (int i = 0; i < 512; ++i) {
        P2i p{(int32_t)(lcg(rs) % 4096), (int32_t)(lcg(rs) % 4096)};
        if (i % 7 == 0) { p.x = i; p.y = 2 * i + 1; }
        uint64_t seed = (i % 3 == 0) ? 1ull : lcg(rs);
        uint64_t h = kazen::Hash(p, seed);
        printf("%s[%d,%d,\"%llu\",\"%llu\"]", i ? "," : "", p.x, p.y, (unsigned long long)seed, (unsigned long long)h);
    }
    printf("],\n\"hash20\": [");
    for (int i = 0; i < 512; ++i) {
        P2i p{(int32_t)(lcg(rs) % 4096), (int32_t)(lcg(rs) % 4096)};
        uint32_t dim = (uint32_t)(lcg(rs) % 64);
        uint64_t seed = (i % 3 == 0) ? 1ull : lcg(rs);
        uint64_t h = kazen::Hash(p, dim, seed);
        printf("%s[%d,%d,%u,\"%llu\",\"%llu\"]", i ? "," : "", p.x, p.y, dim, (unsigned long long)seed, (unsigned long long)h);
    }
    printf("],\n\"mixbits\": [");
    for (int i = 0; i < 256; ++i) {
        uint64_t v = i < 4 ? (uint64_t)i : lcg(rs) * lcg(rs);
        printf("%s[\"%llu\",\"%llu\"]", i ? "," : "", (unsigned long long)v, (unsigned long long)kazen::MixBits(v));
    }
    printf("],\n\"permute\": [");
    for (int i = 0; i < 2048; ++i) {
        uint32_t l = (i % 5 == 0) ? 64u : (uint32_t)(1 + lcg(rs) % 70000);
        uint32_t idx = (uint32_t)(lcg(rs) % l);
        uint32_t p = (uint32_t)lcg(rs);
        printf("%s[%u,%u,%u,%u]", i ? "," : "", idx, l, p, refx::permute(idx, l, p));
    }
    printf("],\n\"pcg32\": [");
    for (int i = 0; i < 256; ++i) {
        uint64_t seed = lcg(rs) * lcg(rs);
        uint64_t delta = (uint64_t)(lcg(rs) % 4096) * 65536ull + (lcg(rs) % 16);
        pcg32 r; r.seed(seed); r.advance((int64_t)delta);
        uint32_t a = r.nextUInt(), b = r.nextUInt();
        float f = r.nextFloat();
        uint32_t fb; memcpy(&fb, &f, 4);
        printf("%s[\"%llu\",\"%llu\",%u,%u,%u]", i ? "," : "", (unsigned long long)seed, (unsigned long long)delta, a, b, fb);
    }
    /* the six vectors quoted in SURVEY.md Appendix B */
    P2i p37{3, 7};
    uint64_t h16 = kazen::Hash(p37, (uint64_t)1), h20 = kazen::Hash(p37, (uint32_t)4, (uint64_t)1);
    pcg32 r; r.seed(h16); r.advance(5 * 65536);
    uint32_t u = r.nextUInt(); float f = r.nextFloat(); uint32_t fb; memcpy(&fb, &f, 4);
    printf("],\n\"survey\": {\"hash16\":\"%llu\",\"hash20\":\"%llu\",\"mix\":\"%llu\",\"pcg_u\":%u,\"pcg_f_bits\":%u,\"perm_a\":%u,\"perm_b\":%u}\n}\n",
           (unsigned long long)h16, (unsigned long long)h20, (unsigned long long)kazen::MixBits(h16), u, fb,
           refx::permute(5, 64, (uint32_t)h20), refx::permute(5, 64, (uint32_t)(h20 * 0x51633e2d)));
    return 0;
}
