/* image.cpp -- minimal image I/O for the host: PNG writer (bitmap.cpp:38-64 wrote PNG through
 * OpenImageIO), PNG (8/16-bit, non-interlaced), PFM, Radiance HDR and scanline OpenEXR readers for imagetexture, and the stand-in
 * pmj02bn / blue-noise tables.  Only zlib is used. */
#include <kazen/scene.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iterator>
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
/* Linear scene-referred output (bitmap.cpp:23-36 saved EXR through OpenImageIO): single-part scanline OpenEXR, three 32-bit float
 * channels, no compression, increasing Y -- the plainest file every EXR reader (ours included) accepts. */
void writeEXR(const std::string &path, int w, int h, const float *rgb) {
    std::vector<uint8_t> o;
    auto put = [&](const void *p, size_t n) { const uint8_t *b = (const uint8_t *)p; o.insert(o.end(), b, b + n); };
    auto str = [&](const char *t) { put(t, strlen(t) + 1); };
    auto i32 = [&](int32_t v) { put(&v, 4); };
    auto f32 = [&](float v) { put(&v, 4); };
    auto attr = [&](const char *name, const char *type, int32_t size) { str(name); str(type); i32(size); };
    const uint8_t magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    put(magic, 8);
    attr("channels", "chlist", 3 * 18 + 1);
    for (const char *c : {"B", "G", "R"}) { str(c); i32(2); const uint8_t lin[4] = {0, 0, 0, 0}; put(lin, 4); i32(1); i32(1); }
    o.push_back(0);
    attr("compression", "compression", 1); o.push_back(0);
    attr("dataWindow", "box2i", 16); i32(0); i32(0); i32(w - 1); i32(h - 1);
    attr("displayWindow", "box2i", 16); i32(0); i32(0); i32(w - 1); i32(h - 1);
    attr("lineOrder", "lineOrder", 1); o.push_back(0);
    attr("pixelAspectRatio", "float", 4); f32(1.f);
    attr("screenWindowCenter", "v2f", 8); f32(0.f); f32(0.f);
    attr("screenWindowWidth", "float", 4); f32(1.f);
    o.push_back(0);
    const uint64_t table = o.size(), block = 8 + (uint64_t)w * 12;
    for (int y = 0; y < h; ++y) { const uint64_t off = table + 8ull * h + block * y; put(&off, 8); }
    std::vector<float> row((size_t)w);
    for (int y = 0; y < h; ++y) {
        i32(y); i32((int32_t)(w * 12));
        for (int c = 2; c >= 0; --c) {          /* B, G, R */
            for (int x = 0; x < w; ++x) row[x] = rgb[3 * ((size_t)y * w + x) + c];
            put(row.data(), (size_t)w * 4);
        }
    }
    std::ofstream f(path, std::ios::binary);
    f.write((const char *)o.data(), (std::streamsize)o.size());
    if (!f) throw std::runtime_error("cannot write " + path);
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

/* header fields come from untrusted files: positive, at most 65536 per side (kzgpu's texture atlas limit), so that the
 * w * h * channels products below cannot overflow */
static bool saneSize(long long w, long long h) { return w > 0 && h > 0 && w <= 65536 && h <= 65536; }

static bool readPFM(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    std::string magic; float scale;
    f >> magic >> w >> h >> scale;
    f.get();
    const int ch = magic == "PF" ? 3 : (magic == "Pf" ? 1 : 0);
    if (!f || !ch || !saneSize(w, h)) { err = "bad PFM header"; return false; }
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
    bool have_ihdr = false;
    std::vector<uint8_t> idat, plte;
    for (size_t p = 8; p + 12 <= d.size();) {
        const uint32_t len = rd32(p); const std::string type((const char *)&d[p + 4], 4);
        if (p + 12 + len > d.size()) { err = "truncated PNG"; return false; }
        if (type == "IHDR") {
            if (len < 13) { err = "bad PNG IHDR"; return false; }
            const uint32_t uw = rd32(p + 8), uh = rd32(p + 12);
            if (!saneSize(uw, uh)) { err = "PNG dimensions out of range"; return false; }
            w = (int)uw; h = (int)uh; depth = d[p + 16]; ctype = d[p + 17]; interlace = d[p + 20]; have_ihdr = true;
        }
        else if (type == "PLTE") plte.assign(d.begin() + p + 8, d.begin() + p + 8 + len);
        else if (type == "IDAT") idat.insert(idat.end(), d.begin() + p + 8, d.begin() + p + 8 + len);
        else if (type == "IEND") break;
        p += 12 + len;
    }
    if (!have_ihdr) { err = "PNG without IHDR"; return false; }
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

/* Radiance RGBE (.hdr): flat and new-style RLE scanlines, -Y H +X W orientation */
static bool readHDR(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    std::string line;
    bool fmt_ok = false;
    while (std::getline(f, line) && !line.empty() && line != "\r") if (line.find("32-bit_rle_rgbe") != std::string::npos) fmt_ok = true;
    if (!std::getline(f, line)) { err = "truncated HDR header"; return false; }
    if (!fmt_ok || sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || !saneSize(w, h)) { err = "unsupported HDR variant (need 32-bit_rle_rgbe, -Y H +X W)"; return false; }
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t p = 0;
    std::vector<uint8_t> scan((size_t)w * 4);
    rgb.resize((size_t)w * h * 3);
    for (int y = 0; y < h; ++y) {
        if (p + 4 <= d.size() && d[p] == 2 && d[p + 1] == 2 && !(d[p + 2] & 0x80) && ((d[p + 2] << 8) | d[p + 3]) == w && w >= 8 && w < 32768) {
            p += 4;
            for (int c = 0; c < 4; ++c)
                for (int x = 0; x < w;) {
                    if (p >= d.size()) { err = "truncated HDR"; return false; }
                    int n = d[p++];
                    if (n > 128) { n -= 128; if (p >= d.size() || x + n > w) { err = "bad HDR run"; return false; } const uint8_t v = d[p++]; while (n--) scan[(size_t)4 * x++ + c] = v; }
                    else { if (n == 0 || p + n > d.size() || x + n > w) { err = "bad HDR run"; return false; } while (n--) scan[(size_t)4 * x++ + c] = d[p++]; }
                }
        } else {
            if (p + (size_t)w * 4 > d.size()) { err = "truncated HDR"; return false; }
            memcpy(scan.data(), &d[p], (size_t)w * 4); p += (size_t)w * 4;
        }
        for (int x = 0; x < w; ++x) {
            const uint8_t *q = &scan[(size_t)4 * x];
            const float sc = q[3] ? std::ldexp(1.0f, (int)q[3] - 136) : 0.f;
            for (int c = 0; c < 3; ++c) rgb[3 * ((size_t)y * w + x) + c] = q[c] * sc;
        }
    }
    return true;
}

/* OpenEXR, single-part scanline files: compression NONE / RLE / ZIPS / ZIP, HALF or FLOAT channels R,G,B (or Y) */
static float halfToFloat(uint16_t hbits) {
    const uint32_t s = (hbits >> 15) & 1u, e = (hbits >> 10) & 31u, m = hbits & 1023u;
    uint32_t u;
    if (e == 0) {
        if (m == 0) u = s << 31;
        else { int k = 0; uint32_t mm = m; while (!(mm & 1024u)) { mm <<= 1; ++k; } u = (s << 31) | ((uint32_t)(113 - k) << 23) | ((mm & 1023u) << 13); }
    } else if (e == 31) u = (s << 31) | 0x7f800000u | (m << 13);
    else u = (s << 31) | ((e + 112u) << 23) | (m << 13);
    float fl; memcpy(&fl, &u, 4); return fl;
}
static bool readEXR(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (d.size() < 16 || d[0] != 0x76 || d[1] != 0x2f || d[2] != 0x31 || d[3] != 0x01) { err = "not an OpenEXR file"; return false; }
    if (d[5] & 0x1E) { err = "tiled / deep / multi-part EXR is not supported"; return false; }
    size_t p = 8;
    auto cstr = [&](std::string &out) { out.clear(); while (p < d.size() && d[p]) out += (char)d[p++]; ++p; return p <= d.size(); };
    auto i32 = [&](size_t o) { int32_t v; memcpy(&v, &d[o], 4); return v; };
    struct Chan { std::string name; int type; };
    std::vector<Chan> chans;
    int comp = -1, xmin = 0, ymin = 0, xmax = -1, ymax = -1;
    for (;;) {
        std::string name, type;
        if (!cstr(name)) { err = "truncated EXR header"; return false; }
        if (name.empty()) break;
        if (!cstr(type) || p + 4 > d.size()) { err = "truncated EXR header"; return false; }
        const int size = i32(p); p += 4;
        if (size < 0 || p + (size_t)size > d.size()) { err = "truncated EXR header"; return false; }
        if (name == "channels") {
            /* name\0, pixel type (4), pLinear + reserved (4), x / y sampling (8) per channel, a lone \0 ends the list; every read stays
             * inside the attribute, which lies inside the file (checked above) */
            size_t q = p; const size_t end = p + (size_t)size;
            while (q < end && d[q]) {
                Chan c; while (q < end && d[q]) c.name += (char)d[q++];
                if (q >= end || q + 1 + 16 > end) { err = "bad EXR channel list"; return false; }
                ++q;
                c.type = i32(q); q += 16;
                chans.push_back(c);
            }
        } else if (name == "compression") { if (size < 1) { err = "bad EXR compression attribute"; return false; } comp = d[p]; }
        else if (name == "dataWindow") { if (size < 16) { err = "bad EXR dataWindow"; return false; } xmin = i32(p); ymin = i32(p + 4); xmax = i32(p + 8); ymax = i32(p + 12); }
        p += (size_t)size;
    }
    const long long lw = (long long)xmax - xmin + 1, lh = (long long)ymax - ymin + 1;
    if (!saneSize(lw, lh) || chans.empty()) { err = "bad EXR header"; return false; }
    w = (int)lw; h = (int)lh;
    if (comp < 0 || comp > 3) { err = "EXR compression other than NONE / RLE / ZIPS / ZIP is not supported"; return false; }
    size_t bytesPerLine = 0;
    std::vector<size_t> chanOff;
    for (const Chan &c : chans) {
        if (c.type != 1 && c.type != 2) { err = "EXR channel type must be HALF or FLOAT"; return false; }
        chanOff.push_back(bytesPerLine); bytesPerLine += (size_t)w * (c.type == 1 ? 2 : 4);
    }
    int idx[3] = {-1, -1, -1};
    for (size_t k = 0; k < chans.size(); ++k) { if (chans[k].name == "R") idx[0] = (int)k; if (chans[k].name == "G") idx[1] = (int)k; if (chans[k].name == "B") idx[2] = (int)k; }
    if (idx[0] < 0) for (size_t k = 0; k < chans.size(); ++k) if (chans[k].name == "Y") idx[0] = idx[1] = idx[2] = (int)k;
    if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0) { err = "EXR needs R,G,B or Y channels"; return false; }
    const int linesPerBlock = comp == 3 ? 16 : 1;
    const int nblocks = (h + linesPerBlock - 1) / linesPerBlock;
    if (p + (size_t)nblocks * 8 > d.size()) { err = "truncated EXR offset table"; return false; }
    rgb.assign((size_t)w * h * 3, 0.f);
    std::vector<uint8_t> tmp, raw;
    for (int b = 0; b < nblocks; ++b) {
        uint64_t off; memcpy(&off, &d[p + (size_t)b * 8], 8);
        if (off + 8 > d.size()) { err = "bad EXR block offset"; return false; }
        const int y0 = i32((size_t)off) - ymin, size = i32((size_t)off + 4);
        if (size < 0 || off + 8 + (size_t)size > d.size() || y0 < 0 || y0 >= h) { err = "bad EXR block"; return false; }
        const int lines = std::min(linesPerBlock, h - y0);
        const size_t expect = bytesPerLine * lines;
        const uint8_t *src = &d[(size_t)off + 8];
        raw.resize(expect);
        if (comp == 0 || (size_t)size == expect) memcpy(raw.data(), src, expect);
        else {
            tmp.resize(expect);
            if (comp == 1) {      /* RLE */
                size_t o = 0; int q = 0;
                while (q < size && o < expect) {
                    const int n = (int8_t)src[q++];
                    if (n < 0) { const int c = -n; if (q + c > size || o + c > expect) { err = "bad EXR RLE"; return false; } memcpy(&tmp[o], src + q, (size_t)c); q += c; o += (size_t)c; }
                    else { const int c = n + 1; if (q >= size || o + c > expect) { err = "bad EXR RLE"; return false; } memset(&tmp[o], src[q++], (size_t)c); o += (size_t)c; }
                }
                if (o != expect) { err = "bad EXR RLE"; return false; }
            } else {
                uLongf rl = (uLongf)expect;
                if (uncompress(tmp.data(), &rl, src, (uLong)size) != Z_OK || rl != expect) { err = "EXR inflate failed"; return false; }
            }
            for (size_t i = 1; i < expect; ++i) tmp[i] = (uint8_t)(tmp[i - 1] + tmp[i] - 128);      /* predictor */
            const size_t half = (expect + 1) / 2;                                                     /* de-interleave */
            for (size_t i = 0; i < expect; ++i) raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
        }
        for (int l = 0; l < lines; ++l)
            for (int c = 0; c < 3; ++c) {
                const Chan &ch = chans[(size_t)idx[c]];
                const uint8_t *q = raw.data() + bytesPerLine * l + chanOff[(size_t)idx[c]];
                for (int x = 0; x < w; ++x) {
                    float v;
                    if (ch.type == 1) { uint16_t hb; memcpy(&hb, q + 2 * (size_t)x, 2); v = halfToFloat(hb); } else memcpy(&v, q + 4 * (size_t)x, 4);
                    rgb[3 * ((size_t)(y0 + l) * w + x) + c] = v;
                }
            }
    }
    return true;
}

bool readImage(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err) {
    std::ifstream probe(path, std::ios::binary);
    if (!probe) { err = "cannot open file"; return false; }
    char m[4] = {0, 0, 0, 0}; probe.read(m, 4);
    if (m[0] == 'P' && (m[1] == 'F' || m[1] == 'f')) return readPFM(path, w, h, rgb, err);
    if ((uint8_t)m[0] == 0x89 && m[1] == 'P') return readPNG(path, w, h, rgb, err);
    if ((uint8_t)m[0] == 0xFF && (uint8_t)m[1] == 0xD8) return readJPEG(path, w, h, rgb, err);
    if (m[0] == '#' && m[1] == '?') return readHDR(path, w, h, rgb, err);
    if ((uint8_t)m[0] == 0x76 && (uint8_t)m[1] == 0x2f && (uint8_t)m[2] == 0x31 && (uint8_t)m[3] == 0x01) return readEXR(path, w, h, rgb, err);
    err = "unsupported image format (this build decodes PNG, JPEG, PFM, Radiance HDR and scanline OpenEXR; the reference used OpenImageIO)";
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
