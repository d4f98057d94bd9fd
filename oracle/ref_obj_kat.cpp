/* ref_obj_kat.cpp -- known-answer generator for the OBJ loader (TEST INFRASTRUCTURE; built only where /root/reference is mounted).
 *
 * Compiles the reference's OWN WavefrontOBJ class (src/kazen/mesh.cpp:200-343, extracted by line range at build time into
 * oracle/_ref/obj_extract.inc) and its string::toUInt / string::tokenize (src/kazen/common.cpp:255-261,281-297) over small stand-ins
 * for what the class touches -- PropertyList, the file resolver, Transform, Timer, LOG, the Eigen matrix members -- and prints, for
 * every tests/golden/obj/*.obj, the vertex / normal / texture-coordinate / index arrays the reference builds, as raw float bits.
 * The toWorld transform is the identity (the stand-in Transform leaves points alone and normalises normals as Eigen's normalized()
 * does: v / sqrt(v.v), the dot product summed as a0*b0 + (a1*b1 + a2*b2)); what is pinned is the parsing, the quad split, the
 * vertex de-duplication order and the index order.  tests/test_host.py compares the C++ host's loader with tests/golden/obj_kat.json. */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
using std::cout; using std::endl;
namespace kazen {
struct Exception : std::runtime_error { template <typename... A> Exception(const char *f, const A &...) : std::runtime_error(f) {} };
namespace string {
#include "_ref/common_string_extract.inc"
}
struct V3 { float v[3] = {0, 0, 0}; float &x() { return v[0]; } float &y() { return v[1]; } float &z() { return v[2]; }
            V3 normalized() const { const float z = v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]); if (!(z > 0.f)) return *this; const float s = std::sqrt(z); V3 r; r.v[0] = v[0] / s; r.v[1] = v[1] / s; r.v[2] = v[2] / s; return r; } };
struct V2 { float v[2] = {0, 0}; float &x() { return v[0]; } float &y() { return v[1]; } };
typedef V3 Vector3f; typedef V3 Point3f; typedef V3 Normal3f; typedef V2 Vector2f; typedef V2 Point2f;
struct Transform { V3 operator*(const V3 &p) const { return p; } };
struct BoundingBox3f { void expandBy(const V3 &) {} };
template <typename S> struct Mat {
    int r = 0, c = 0; std::vector<S> d;
    void resize(int rows, size_t cols) { r = rows; c = (int)cols; d.assign((size_t)rows * cols, S(0)); }
    S *data() { return d.data(); } size_t size() const { return d.size(); } int cols() const { return c; }
    struct Col { Mat *m; int j; void operator=(V3 p) { for (int i = 0; i < m->r; ++i) m->d[(size_t)j * m->r + i] = p.v[i]; } void operator=(V2 p) { for (int i = 0; i < m->r; ++i) m->d[(size_t)j * m->r + i] = p.v[i]; } };
    Col col(int j) { return Col{this, j}; }
};
typedef Mat<float> MatrixXf; typedef Mat<uint32_t> MatrixXu;
struct PropertyList { std::string file; std::string getString(const std::string &) const { return file; } Transform getTransform(const std::string &, const Transform &d) const { return d; } };
namespace filesystem { struct path { std::string s; const std::string &str() const { return s; } }; }
struct Resolver { filesystem::path resolve(const std::string &s) const { return filesystem::path{s}; } };
static Resolver *getFileResolver() { static Resolver r; return &r; }
struct Timer { double elapsed() const { return 0; } };
namespace util { static std::string timeString(double) { return ""; } }
#define LOG(...) ((void)0)
struct Mesh { MatrixXf m_V, m_N, m_UV; MatrixXu m_F; BoundingBox3f m_bbox; std::string m_name; };
#include "_ref/obj_extract.inc"
}

int main(int argc, char **argv) {
    printf("{\n \"generator\": \"oracle/ref_obj_kat.cpp over the reference's mesh.cpp:200-343 and common.cpp:255-261,281-297\",\n \"files\": {\n");
    for (int a = 1; a < argc; ++a) {
        kazen::PropertyList pl; pl.file = argv[a];
        kazen::WavefrontOBJ *m = new kazen::WavefrontOBJ(pl);
        const char *base = strrchr(argv[a], '/'); base = base ? base + 1 : argv[a];
        printf("  \"%s\": {", base);
        auto dumpf = [&](const char *k, kazen::MatrixXf &M, bool comma) { printf("\"%s\": [", k); for (size_t i = 0; i < M.size(); ++i) { uint32_t u; memcpy(&u, &M.d[i], 4); printf("%s%u", i ? "," : "", u); } printf("]%s", comma ? ", " : ""); };
        dumpf("V", m->m_V, true); dumpf("N", m->m_N, true); dumpf("UV", m->m_UV, true);
        printf("\"F\": ["); for (size_t i = 0; i < m->m_F.size(); ++i) printf("%s%u", i ? "," : "", m->m_F.d[i]); printf("]}%s\n", a + 1 < argc ? "," : "");
    }
    printf(" }\n}\n");
    return 0;
}
