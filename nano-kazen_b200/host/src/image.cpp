/* image.cpp -- minimal image I/O for the host: PNG writer (bitmap.cpp:38-64 wrote PNG through
 * OpenImageIO), PNG (8/16-bit, non-interlaced) and PFM readers for imagetexture, and the stand-in
 * pmj02bn / blue-noise tables.  Only zlib is used. */
#include <kazen/scene.h>
#include <cmath>
#include <cstring>
#include <fstream>
#include <zlib.h>

namespace kazen {

static void put32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static void chunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data) {
    put32(out, (uint32_t)data.size());
    std::vector<uint8_t> td(type, type + 4);
    td.insert(td.end(), data.begin(), data.end());
    out.insert(out.end(), td.begin(), td.end());
    put32(out, (uint32_t)crc32(0L, td.data(), (uInt)td.size()));
}
void writePNG(const std::string &path, int w, int h, const uint8_t *rgb8) {
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)w); put32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    std::vector<uint8_t> raw((size_t)h * (1 + (size_t)w * 3));
    for (int y = 0; y < h; ++y) { raw[(size_t)y * (1 + (size_t)w * 3)] = 0; memcpy(&raw[(size_t)y * (1 + (size_t)w * 3) + 1], rgb8 + (size_t)y * w * 3, (size_t)w * 3); }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw Exception("PNG: deflate failed");
    comp.resize(clen);
    chunk(out, "IDAT", comp);
    chunk(out, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Exception("cannot write \"" + path + "\"");
    f.write((const char *)out.data(), (std::streamsize)out.size());
}

static bool readPFM(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    std::string magic; float scale;
    f >> magic >> w >> h >> scale;
    f.get();
    const int ch = magic == "PF" ? 3 : (magic == "Pf" ? 1 : 0);
    if (!f || !ch || w <= 0 || h <= 0) { err = "bad PFM header"; return false; }
    std::vector<float> buf((size_t)w * h * ch);
    if (!f.read((char *)buf.data(), (std::streamsize)(buf.size() * 4))) { err = "truncated PFM"; return false; }
    if (scale > 0) { for (float &v : buf) { uint32_t u; memcpy(&u, &v, 4); u = __builtin_bswap32(u); memcpy(&v, &u, 4); } }
    rgb.resize((size_t)w * h * 3);
    for (int y = 0; y < h; ++y)          /* PFM stores bottom-to-top; row 0 of kz_image_desc = first scanline of the picture */
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) rgb[3 * ((size_t)y * w + x) + c] = buf[((size_t)(h - 1 - y) * w + x) * ch + (ch == 3 ? c : 0)];
    return true;
}

static bool readPNG(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (d.size() < 33 || memcmp(d.data(), sig, 8) != 0) { err = "not a PNG file"; return false; }
    auto rd32 = [&](size_t o) { return ((uint32_t)d[o] << 24) | ((uint32_t)d[o + 1] << 16) | ((uint32_t)d[o + 2] << 8) | d[o + 3]; };
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    for (size_t p = 8; p + 12 <= d.size();) {
        const uint32_t len = rd32(p); const std::string type((const char *)&d[p + 4], 4);
        if (p + 12 + len > d.size()) { err = "truncated PNG"; return false; }
        if (type == "IHDR") { w = (int)rd32(p + 8); h = (int)rd32(p + 12); depth = d[p + 16]; ctype = d[p + 17]; interlace = d[p + 20]; }
        else if (type == "PLTE") plte.assign(d.begin() + p + 8, d.begin() + p + 8 + len);
        else if (type == "IDAT") idat.insert(idat.end(), d.begin() + p + 8, d.begin() + p + 8 + len);
        else if (type == "IEND") break;
        p += 12 + len;
    }
    if (interlace) { err = "interlaced PNG is not supported"; return false; }
    if (depth != 8 && depth != 16) { err = "only 8/16-bit PNG is supported"; return false; }
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch || (ctype == 3 && depth != 8)) { err = "unsupported PNG colour type"; return false; }
    const size_t bpp = (size_t)ch * depth / 8, stride = (size_t)w * bpp;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rl = (uLongf)raw.size();
    if (uncompress(raw.data(), &rl, idat.data(), (uLong)idat.size()) != Z_OK || rl != raw.size()) { err = "PNG inflate failed"; return false; }
    std::vector<uint8_t> img(stride * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y]; const uint8_t *src = &raw[(stride + 1) * y + 1];
        uint8_t *dst = &img[stride * y]; const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= bpp ? dst[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int pr = 0;
            switch (ft) {
                case 0: pr = 0; break; case 1: pr = a; break; case 2: pr = b; break; case 3: pr = (a + b) / 2; break;
                case 4: { const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c); pr = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
                default: err = "bad PNG filter"; return false;
            }
            dst[x] = (uint8_t)(src[x] + pr);
        }
    }
    rgb.resize((size_t)w * h * 3);
    const float s = depth == 8 ? 1.f / 255.f : 1.f / 65535.f;
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        float v[4] = {0, 0, 0, 1};
        for (int c = 0; c < ch; ++c) v[c] = (depth == 8 ? img[i * bpp + c] : (img[i * bpp + 2 * c] << 8 | img[i * bpp + 2 * c + 1])) * s;
        if (ctype == 3) { const size_t k = img[i]; if (3 * k + 2 >= plte.size()) { err = "PNG palette index out of range"; return false; } for (int c = 0; c < 3; ++c) v[c] = plte[3 * k + c] / 255.f; }
        else if (ch <= 2) v[1] = v[2] = v[0];
        rgb[3 * i] = v[0]; rgb[3 * i + 1] = v[1]; rgb[3 * i + 2] = v[2];
    }
    return true;
}

bool readImage(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err) {
    std::ifstream probe(path, std::ios::binary);
    if (!probe) { err = "cannot open file"; return false; }
    char m[2] = {0, 0}; probe.read(m, 2);
    if (m[0] == 'P' && (m[1] == 'F' || m[1] == 'f')) return readPFM(path, w, h, rgb, err);
    if ((uint8_t)m[0] == 0x89 && m[1] == 'P') return readPNG(path, w, h, rgb, err);
    err = "unsupported image format (this build decodes PNG and PFM; the reference used OpenImageIO)";
    return false;
}

/* ---- stand-in sample tables ------------------------------------------------------------------ */
static uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
static uint32_t reverseBits(uint32_t x) {
    x = (x >> 16) | (x << 16); x = ((x & 0x00ff00ffu) << 8) | ((x & 0xff00ff00u) >> 8); x = ((x & 0x0f0f0f0fu) << 4) | ((x & 0xf0f0f0f0u) >> 4);
    x = ((x & 0x33333333u) << 2) | ((x & 0xccccccccu) >> 2); x = ((x & 0x55555555u) << 1) | ((x & 0xaaaaaaaau) >> 1);
    return x;
}
/* Laine-Karras style nested uniform (Owen) scramble of a 32-bit fixed-point value */
static uint32_t owenScramble(uint32_t x, uint32_t seed) {
    x = reverseBits(x);
    x += seed; x ^= x * 0x6c50b47cu; x ^= x * 0xb82f1e52u; x ^= x * 0xc7afe638u; x ^= x * 0x8d22f6e6u;
    return reverseBits(x);
}
void fallbackPmjTables(std::vector<uint16_t> &blueNoise, std::vector<uint32_t> &pmj) {
    blueNoise.resize((size_t)48 * 128 * 128);
    for (size_t i = 0; i < blueNoise.size(); ++i) blueNoise[i] = (uint16_t)(mix32((uint32_t)i * 0x9e3779b9u + 0x1234567u) >> 16);
    pmj.resize((size_t)5 * 65536 * 2);
    for (uint32_t set = 0; set < 5; ++set)
        for (uint32_t i = 0; i < 65536; ++i) {
            /* Sobol' (0,2)-sequence: dimension 0 = van der Corput, dimension 1 = second Sobol' matrix */
            const uint32_t x = reverseBits(i);
            uint32_t y = 0, v = 1u << 31;
            for (uint32_t k = i; k; k >>= 1, v ^= v >> 1) if (k & 1u) y ^= v;
            pmj[((size_t)set * 65536 + i) * 2] = owenScramble(x, mix32(2 * set + 1));
            pmj[((size_t)set * 65536 + i) * 2 + 1] = owenScramble(y, mix32(2 * set + 2));
        }
}

}  // namespace kazen
