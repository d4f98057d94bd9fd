/* jpeg.cpp -- baseline / extended-sequential Huffman JPEG decoder for imagetexture (texture.cpp:46-64 read every texture
 * through OpenImageIO, whose JPEG reader is libjpeg with its defaults; the reference's own look-dev material,
 * doc/2022_q1/2022_q1_report.md:183-220, uses .jpg textures).
 *
 * To decode the SAME texels as the reference, the arithmetic is the published one of libjpeg's defaults:
 *   - the "slow integer" inverse DCT (Loeffler-Ligtenberg-Moschytz, 13-bit constants, 2 extra bits after the column pass);
 *   - "fancy" chroma upsampling: the triangle filter for 2:1 horizontal (3/4, 1/4) and 2x2 (9/16, 3/16, 3/16, 1/16) with the two
 *     rounding biases per output pair, replication for every other ratio and for components of width <= 2;
 *   - the 16-bit fixed-point YCbCr -> RGB tables (1.402, 1.772, 0.71414, 0.34414).
 * tests/test_host.py compares the result with Pillow's libjpeg-turbo decode bit for bit.
 * Not supported (clear error): progressive, arithmetic coding, 12-bit, lossless, CMYK. */
#include <kazen/scene.h>
#include <cstring>
#include <fstream>
#include <iterator>

namespace kazen {
namespace {

const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool present = false;
    uint8_t bits[17] = {0}, vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    uint16_t look[512];                 /* 9-bit lookahead: (length << 8) | symbol, 0 = longer code */
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7FFFFFFF;
        memset(look, 0, sizeof(look));
        code = 0; k = 0;
        for (int l = 1; l <= 9; ++l) {
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                const int first = code << (9 - l);
                for (int f = 0; f < (1 << (9 - l)); ++f) look[first + f] = (uint16_t)((l << 8) | vals[k]);
            }
            code <<= 1;
        }
    }
};

struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0; int bw = 0, bh = 0; /* blocks per row / column of the padded plane */
              int dw = 0, dh = 0;   /* real (downsampled) size */ std::vector<uint8_t> plane; int stride = 0; };

struct BitReader {
    const uint8_t *p, *end; uint32_t acc = 0; int cnt = 0; bool marker = false;
    void fill() {
        while (cnt <= 24) {
            int b = 0;
            if (!marker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else { marker = true; b = 0; }              /* a marker ends the segment: feed zeros */
                } else ++p;
            }
            acc |= (uint32_t)b << (24 - cnt); cnt += 8;
        }
    }
    int peek(int n) { if (cnt < n) fill(); return (int)(acc >> (32 - n)); }
    void skip(int n) { acc <<= n; cnt -= n; }
    int get(int n) { if (n == 0) return 0; const int v = peek(n); skip(n); return v; }
    void align() { acc = 0; cnt = 0; }
};

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

inline int decodeSym(BitReader &br, const Huff &h, bool &ok) {
    const int pk = br.peek(9);
    const uint16_t e = h.look[pk];
    if (e) { br.skip(e >> 8); return e & 0xFF; }
    int code = br.peek(16), l = 10;
    for (; l <= 16; ++l) if ((code >> (16 - l)) <= h.maxcode[l] && h.maxcode[l] >= 0) break;
    if (l > 16) { ok = false; return 0; }
    br.skip(l);
    return h.vals[h.valptr[l] + (code >> (16 - l)) - h.mincode[l]];
}

/* libjpeg's jidctint.c ("ISLOW"), dequantisation folded in as there */
inline int descale(long x, int n) { return (int)((x + (1L << (n - 1))) >> n); }
inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
void idctIslow(const int16_t *coef, const uint16_t *q, uint8_t *out, int stride) {
    const long F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299, F1_847 = 15137,
               F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
    const int CB = 13, P1 = 2;
    long ws[64];
    for (int c = 0; c < 8; ++c) {
        const int16_t *in = coef + c; const uint16_t *qq = q + c; long *w = ws + c;
        if (!in[8] && !in[16] && !in[24] && !in[32] && !in[40] && !in[48] && !in[56]) {
            const long dc = ((long)in[0] * qq[0]) << P1;
            for (int r = 0; r < 8; ++r) w[8 * r] = dc;
            continue;
        }
        long z2 = (long)in[16] * qq[16], z3 = (long)in[48] * qq[48];
        long z1 = (z2 + z3) * F0_541;
        long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        z2 = (long)in[0] * qq[0]; z3 = (long)in[32] * qq[32];
        long tmp0 = (z2 + z3) << CB, tmp1 = (z2 - z3) << CB;
        const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = (long)in[56] * qq[56]; tmp1 = (long)in[40] * qq[40]; tmp2 = (long)in[24] * qq[24]; tmp3 = (long)in[8] * qq[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; long z4 = tmp1 + tmp3;
        const long z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        w[0] = descale(tmp10 + tmp3, CB - P1); w[56] = descale(tmp10 - tmp3, CB - P1);
        w[8] = descale(tmp11 + tmp2, CB - P1); w[48] = descale(tmp11 - tmp2, CB - P1);
        w[16] = descale(tmp12 + tmp1, CB - P1); w[40] = descale(tmp12 - tmp1, CB - P1);
        w[24] = descale(tmp13 + tmp0, CB - P1); w[32] = descale(tmp13 - tmp0, CB - P1);
    }
    for (int r = 0; r < 8; ++r) {
        const long *w = ws + 8 * r; uint8_t *o = out + (size_t)r * stride;
        long z2 = w[2], z3 = w[6];
        long z1 = (z2 + z3) * F0_541;
        long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        long tmp0 = (w[0] + w[4]) << CB, tmp1 = (w[0] - w[4]) << CB;
        const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; long z4 = tmp1 + tmp3;
        const long z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int S = CB + P1 + 3;
        o[0] = clamp8(descale(tmp10 + tmp3, S) + 128); o[7] = clamp8(descale(tmp10 - tmp3, S) + 128);
        o[1] = clamp8(descale(tmp11 + tmp2, S) + 128); o[6] = clamp8(descale(tmp11 - tmp2, S) + 128);
        o[2] = clamp8(descale(tmp12 + tmp1, S) + 128); o[5] = clamp8(descale(tmp12 - tmp1, S) + 128);
        o[3] = clamp8(descale(tmp13 + tmp0, S) + 128); o[4] = clamp8(descale(tmp13 - tmp0, S) + 128);
    }
}

/* Full-resolution plane of one component (W x H) from its downsampled plane: libjpeg's jdsample.c */
void upsample(const Comp &c, int hmax, int vmax, int W, int H, std::vector<uint8_t> &out) {
    out.assign((size_t)W * H, 0);
    const int hr = hmax / c.h, vr = vmax / c.v, dw = c.dw, dh = c.dh;
    auto src = [&](int x, int y) -> int { return c.plane[(size_t)y * c.stride + x]; };
    const bool fancy = dw > 2;
    if (hr == 1 && vr == 1) {
        for (int y = 0; y < H; ++y) memcpy(&out[(size_t)y * W], &c.plane[(size_t)y * c.stride], (size_t)W);
    } else if (hr == 2 && vr == 1 && fancy) {                                          /* h2v1_fancy_upsample */
        std::vector<uint8_t> row((size_t)2 * dw);
        for (int y = 0; y < H; ++y) {
            for (int x = 0; x < dw; ++x) {
                const int v = src(x, y);
                row[2 * x] = x == 0 ? (uint8_t)v : (uint8_t)((3 * v + src(x - 1, y) + 1) >> 2);
                row[2 * x + 1] = x == dw - 1 ? (uint8_t)v : (uint8_t)((3 * v + src(x + 1, y) + 2) >> 2);
            }
            memcpy(&out[(size_t)y * W], row.data(), (size_t)W);
        }
    } else if (hr == 2 && vr == 2 && fancy) {                                          /* h2v2_fancy_upsample */
        std::vector<int> sum((size_t)dw); std::vector<uint8_t> row((size_t)2 * dw);
        for (int y = 0; y < H; ++y) {
            const int i = y >> 1, near = i, far = (y & 1) ? (i + 1 < dh ? i + 1 : dh - 1) : (i > 0 ? i - 1 : 0);
            for (int x = 0; x < dw; ++x) sum[x] = 3 * src(x, near) + src(x, far);
            for (int x = 0; x < dw; ++x) {
                const int t = sum[x];
                row[2 * x] = x == 0 ? (uint8_t)((t * 4 + 8) >> 4) : (uint8_t)((3 * t + sum[x - 1] + 8) >> 4);
                row[2 * x + 1] = x == dw - 1 ? (uint8_t)((t * 4 + 7) >> 4) : (uint8_t)((3 * t + sum[x + 1] + 7) >> 4);
            }
            memcpy(&out[(size_t)y * W], row.data(), (size_t)W);
        }
    } else if (hr == 1 && vr == 2) {                                                   /* h1v2_fancy_upsample (libjpeg-turbo >= 2.0) */
        for (int y = 0; y < H; ++y) {
            const int i = y >> 1, far = (y & 1) ? (i + 1 < dh ? i + 1 : dh - 1) : (i > 0 ? i - 1 : 0), bias = (y & 1) ? 2 : 1;
            for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = (uint8_t)((3 * src(x, i) + src(x, far) + bias) >> 2);
        }
    } else {                                                                           /* int_upsample: replication */
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = (uint8_t)src(std::min(x / hr, dw - 1), std::min(y / vr, dh - 1));
    }
}

}  // namespace

bool readJPEG(const std::string &path, int &W, int &H, std::vector<float> &rgb, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (d.size() < 4 || d[0] != 0xFF || d[1] != 0xD8) { err = "not a JPEG file"; return false; }
    uint16_t quant[4][64]; bool haveQ[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    std::vector<Comp> comps;
    int restart = 0, hmax = 1, vmax = 1, adobeTransform = -1;
    bool haveFrame = false, jfif = false;
    W = H = 0;
    size_t p = 2;
    auto u16 = [&](size_t o) { return (int)d[o] << 8 | d[o + 1]; };
    for (;;) {
        while (p < d.size() && d[p] != 0xFF) ++p;
        while (p < d.size() && d[p] == 0xFF) ++p;
        if (p >= d.size()) { err = "truncated JPEG (no EOI)"; return false; }
        const int m = d[p++];
        if (m == 0xD9) break;                                   /* EOI */
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;    /* stand-alone markers */
        if (p + 2 > d.size()) { err = "truncated JPEG"; return false; }
        const int len = u16(p);
        if (len < 2 || p + (size_t)len > d.size()) { err = "truncated JPEG segment"; return false; }
        const size_t seg = p + 2, segEnd = p + (size_t)len;
        if (m == 0xDB) {                                        /* DQT */
            size_t q = seg;
            while (q < segEnd) {
                const int pq = d[q] >> 4, tq = d[q] & 15; ++q;
                if (tq > 3 || q + (size_t)(pq ? 128 : 64) > segEnd) { err = "bad DQT"; return false; }
                for (int i = 0; i < 64; ++i) { quant[tq][kZigzag[i]] = (uint16_t)(pq ? u16(q) : d[q]); q += pq ? 2 : 1; }
                haveQ[tq] = true;
            }
        } else if (m == 0xC4) {                                 /* DHT */
            size_t q = seg;
            while (q < segEnd) {
                const int tc = d[q] >> 4, th = d[q] & 15; ++q;
                if (tc > 1 || th > 3 || q + 16 > segEnd) { err = "bad DHT"; return false; }
                Huff &h = tc ? ac[th] : dc[th];
                int n = 0;
                for (int i = 1; i <= 16; ++i) { h.bits[i] = d[q++]; n += h.bits[i]; }
                if (n > 256 || q + (size_t)n > segEnd) { err = "bad DHT"; return false; }
                memcpy(h.vals, &d[q], (size_t)n); q += (size_t)n;
                h.present = true; h.build();
            }
        } else if (m == 0xC0 || m == 0xC1) {                    /* SOF0 / SOF1 */
            if (len < 8 || d[seg] != 8) { err = "only 8-bit JPEG is supported"; return false; }
            H = u16(seg + 1); W = u16(seg + 3);
            const int n = d[seg + 5];
            if (W <= 0 || H <= 0) { err = "bad JPEG size"; return false; }
            if (n != 1 && n != 3) { err = "only grayscale and 3-component JPEG are supported (CMYK is not)"; return false; }
            if (len < 8 + 3 * n) { err = "bad SOF"; return false; }
            comps.resize((size_t)n);
            for (int i = 0; i < n; ++i) {
                Comp &c = comps[(size_t)i];
                c.id = d[seg + 6 + 3 * i]; c.h = d[seg + 7 + 3 * i] >> 4; c.v = d[seg + 7 + 3 * i] & 15; c.tq = d[seg + 8 + 3 * i];
                if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) { err = "bad SOF component"; return false; }
                hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v);
            }
            for (const Comp &c : comps) if (hmax % c.h || vmax % c.v) { err = "fractional chroma sampling ratios are not supported"; return false; }
            const int mx = (W + 8 * hmax - 1) / (8 * hmax), my = (H + 8 * vmax - 1) / (8 * vmax);
            for (Comp &c : comps) {
                c.bw = mx * c.h; c.bh = my * c.v; c.stride = c.bw * 8;
                c.dw = (W * c.h + hmax - 1) / hmax; c.dh = (H * c.v + vmax - 1) / vmax;
                c.plane.assign((size_t)c.stride * c.bh * 8, 0);
            }
            haveFrame = true;
        } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            err = m == 0xC2 ? "progressive JPEG is not supported (re-save as baseline)" : "this JPEG process (arithmetic / lossless / hierarchical) is not supported";
            return false;
        } else if (m == 0xDD) { restart = u16(seg); }
        else if (m == 0xE0 && len >= 7 && !memcmp(&d[seg], "JFIF", 5)) jfif = true;
        else if (m == 0xEE && len >= 14 && !memcmp(&d[seg], "Adobe", 5)) adobeTransform = d[seg + 11];
        else if (m == 0xDA) {                                   /* SOS + entropy-coded segment */
            if (!haveFrame) { err = "SOS before SOF"; return false; }
            const int ns = d[seg];
            if (ns < 1 || ns > (int)comps.size() || len < 6 + 2 * ns) { err = "bad SOS"; return false; }
            std::vector<Comp *> sc;
            for (int i = 0; i < ns; ++i) {
                const int id = d[seg + 1 + 2 * i], t = d[seg + 2 + 2 * i];
                Comp *c = nullptr;
                for (Comp &k : comps) if (k.id == id) c = &k;
                if (!c) { err = "SOS names an unknown component"; return false; }
                c->td = t >> 4; c->ta = t & 15; c->pred = 0;
                if (c->td > 3 || c->ta > 3 || !dc[c->td].present || !ac[c->ta].present || !haveQ[c->tq]) { err = "missing Huffman / quantisation table"; return false; }
                sc.push_back(c);
            }
            BitReader br; br.p = &d[segEnd]; br.end = d.data() + d.size();
            const int mx = (W + 8 * hmax - 1) / (8 * hmax), my = (H + 8 * vmax - 1) / (8 * vmax);
            /* a single-component scan is not interleaved: its MCU is one block, over the component's own block grid */
            const bool inter = ns > 1;
            const int ux = inter ? mx : (sc[0]->dw + 7) / 8, uy = inter ? my : (sc[0]->dh + 7) / 8;
            int16_t coef[64];
            int left = restart; bool ok = true;
            for (int y = 0; y < uy && ok; ++y)
                for (int x = 0; x < ux && ok; ++x) {
                    if (restart && left == 0) {
                        br.align();
                        const uint8_t *q = br.p;                 /* the reader stops in front of the marker */
                        while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                        if (q + 1 >= br.end) { ok = false; break; }
                        br.p = q + 2; br.marker = false;
                        for (Comp *c : sc) c->pred = 0;
                        left = restart;
                    }
                    for (Comp *c : sc) {
                        const int nh = inter ? c->h : 1, nv = inter ? c->v : 1;
                        for (int by = 0; by < nv; ++by)
                            for (int bx = 0; bx < nh; ++bx) {
                                memset(coef, 0, sizeof(coef));
                                int s = decodeSym(br, dc[c->td], ok);
                                if (s > 11) ok = false;
                                if (!ok) break;
                                c->pred += s ? extend(br.get(s), s) : 0;
                                coef[0] = (int16_t)c->pred;
                                for (int k = 1; k < 64; ++k) {
                                    const int rs = decodeSym(br, ac[c->ta], ok);
                                    if (!ok) break;
                                    const int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) { if (r == 15) { k += 15; continue; } break; }
                                    k += r;
                                    if (k > 63) { ok = false; break; }
                                    coef[kZigzag[k]] = (int16_t)extend(br.get(sz), sz);
                                }
                                const int gx = x * nh + bx, gy = y * nv + by;
                                if (gx < c->bw && gy < c->bh) idctIslow(coef, quant[c->tq], &c->plane[((size_t)gy * 8) * c->stride + (size_t)gx * 8], c->stride);
                            }
                    }
                    if (restart) --left;
                }
            if (!ok) { err = "corrupt JPEG entropy data"; return false; }
            p = (size_t)(br.p - d.data());
            continue;
        }
        p = segEnd;
    }
    if (!haveFrame) { err = "JPEG without a frame"; return false; }
    rgb.assign((size_t)W * H * 3, 0.f);
    std::vector<uint8_t> full[3];
    for (size_t i = 0; i < comps.size(); ++i) upsample(comps[i], hmax, vmax, W, H, full[i]);
    const float k = 1.0f / 255.0f;
    if (comps.size() == 1) {
        for (size_t i = 0; i < (size_t)W * H; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = full[0][i] * k;
        return true;
    }
    /* jdcolor.c: Adobe transform 0 (or component ids 'R','G','B' without JFIF) means the file stores RGB */
    const bool isRGB = adobeTransform == 0 || (!jfif && adobeTransform < 0 && comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B');
    int crR[256], cbB[256]; long crG[256], cbG[256];
    for (int i = 0; i < 256; ++i) {
        const long x = i - 128;
        crR[i] = (int)((91881L * x + 32768L) >> 16); cbB[i] = (int)((116130L * x + 32768L) >> 16);
        crG[i] = -46802L * x; cbG[i] = -22554L * x + 32768L;
    }
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        const int y = full[0][i], cb = full[1][i], cr = full[2][i];
        if (isRGB) { rgb[3 * i] = y * k; rgb[3 * i + 1] = cb * k; rgb[3 * i + 2] = cr * k; continue; }
        rgb[3 * i] = clamp8(y + crR[cr]) * k;
        rgb[3 * i + 1] = clamp8(y + (int)((cbG[cb] + crG[cr]) >> 16)) * k;
        rgb[3 * i + 2] = clamp8(y + cbB[cb]) * k;
    }
    return true;
}

}  // namespace kazen
