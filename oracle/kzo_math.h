/* ORACLE (test infrastructure) -- scalar float math restating the Eigen-based types the
 * reference uses (include/kazen/vector.h, color.h, frame.h, common.cpp).
 * Compile with -ffp-contract=off: every reference-side float op is an individually rounded
 * IEEE single op (CMakeLists.txt:31-40 sets no -mfma/-ffast-math). */
#ifndef KZO_MATH_H
#define KZO_MATH_H
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace kzo {

static const float kEpsilon = 1e-5f;                       /* common.h:27 */
static const float kOneMinusEpsilon = 0x1.fffffep-1f;      /* common.h:28 */
static const float kPi = 3.14159265358979323846f;          /* common.h:33 */
static const float kInvPi = 0.31830988618379067154f;       /* common.h:34 */

struct V2 { float x, y; };
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a) : x(a), y(a), z(a) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 &operator+=(V3 &a, V3 b) { a = a + b; return a; }
inline V3 &operator-=(V3 &a, V3 b) { a = a - b; return a; }
inline V3 &operator*=(V3 &a, V3 b) { a = a * b; return a; }
inline V3 &operator/=(V3 &a, float s) { a = a / s; return a; }
/* Eigen dot of a fixed 3-vector: ((x*x' + y*y') + z*z') */
/* Eigen evaluates a fixed-size dot product as a fully unrolled reduction that splits the range in halves (Redux.h,
 * redux_novec_unroller): three terms sum as a0 + (a1 + a2). */
inline float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline V3 cross(V3 a, V3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float sqnorm(V3 a) { return dot(a, a); }
inline float norm(V3 a) { return std::sqrt(sqnorm(a)); }
/* Eigen::MatrixBase::normalized(): v / sqrt(squaredNorm) when squaredNorm > 0 */
inline V3 normalized(V3 a) {
    float z = sqnorm(a);
    if (z > 0.f) return a / std::sqrt(z);
    return a;
}
inline float maxcoeff(V3 a) { return std::max(a.x, std::max(a.y, a.z)); }
inline bool iszero(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }

inline float sqr(float x) { return x * x; }                                  /* common.h:456 */
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); } /* common.h:237 */
inline float lerpf(float t, float a, float b) { return (1.f - t) * a + t * b; } /* common.h:255 */

/* common.cpp:436-445 */
inline void coordinateSystem(const V3 &a, V3 &b, V3 &c) {
    if (std::fabs(a.x) > std::fabs(a.y)) {
        float invLen = 1.0f / std::sqrt(a.x * a.x + a.z * a.z);
        c = V3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / std::sqrt(a.y * a.y + a.z * a.z);
        c = V3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}

/* frame.h:13-44 */
struct Frame {
    V3 s, t, n;
    Frame() {}
    explicit Frame(const V3 &n_) : n(n_) { coordinateSystem(n, s, t); }
    V3 toLocal(const V3 &v) const { return V3(dot(v, s), dot(v, t), dot(v, n)); }
    V3 toWorld(const V3 &v) const { return s * v.x + t * v.y + n * v.z; }
};

/* common.cpp:536-538 */
inline V3 reflect(const V3 &wi, const V3 &n) { return 2 * dot(n, wi) * n - wi; }

/* common.cpp:393-395 */
inline float luminance(V3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }

/* common.cpp:368-382 */
inline V3 toLinearRGB(V3 c) {
    float v[3] = {c.x, c.y, c.z}, r[3];
    for (int i = 0; i < 3; ++i)
        r[i] = v[i] <= 0.04045f ? v[i] * (1.0f / 12.92f) : std::pow((v[i] + 0.055f) * (1.0f / 1.055f), 2.4f);
    return V3(r[0], r[1], r[2]);
}
/* common.cpp:352-366 */
inline V3 toSRGB(V3 c) {
    float v[3] = {c.x, c.y, c.z}, r[3];
    for (int i = 0; i < 3; ++i)
        r[i] = v[i] <= 0.0031308f ? 12.92f * v[i] : (1.0f + 0.055f) * std::pow(v[i], 1.0f / 2.4f) - 0.055f;
    return V3(r[0], r[1], r[2]);
}
/* common.cpp:384-391 */
inline bool colorValid(V3 c) {
    float v[3] = {c.x, c.y, c.z};
    for (int i = 0; i < 3; ++i)
        if (v[i] < 0 || !std::isfinite(v[i])) return false;
    return true;
}

/* Row-major 4x4 applied like Transform (transform.h:49-62). */
struct M44 { float m[16]; };
inline V3 xformPoint(const M44 &M, V3 p) {
    float r[4];
    for (int i = 0; i < 4; ++i)
        r[i] = M.m[i * 4 + 0] * p.x + M.m[i * 4 + 1] * p.y + M.m[i * 4 + 2] * p.z + M.m[i * 4 + 3] * 1.0f;
    return V3(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
}
inline V3 xformVector(const M44 &M, V3 v) {
    /* topLeftCorner<3,3>() * v: each coefficient is the same unrolled reduction as dot() */
    return V3(M.m[0] * v.x + (M.m[1] * v.y + M.m[2] * v.z),
              M.m[4] * v.x + (M.m[5] * v.y + M.m[6] * v.z),
              M.m[8] * v.x + (M.m[9] * v.y + M.m[10] * v.z));
}

/* warp.cpp:41-50 (math::sincosf -> sinf/cosf, common.h:231) */
inline V2 squareToUniformDisk(V2 s) {
    float r = std::sqrt(s.x);
    float a = 2.0f * kPi * s.y;
    float sinPhi = std::sin(a), cosPhi = std::cos(a);
    return V2{cosPhi * r, sinPhi * r};
}
/* warp.cpp:85-115 */
inline V3 squareToCosineHemisphere(V2 s) {
    float r1 = 2.0f * s.x - 1.0f;
    float r2 = 2.0f * s.y - 1.0f;
    float phi, r;
    if (r1 == 0 && r2 == 0) {
        r = phi = 0;
    } else if (r1 * r1 > r2 * r2) {
        r = r1;
        phi = (kPi / 4.0f) * (r2 / r1);
    } else {
        r = r2;
        phi = (kPi / 2.0f) - (r1 / r2) * (kPi / 4.0f);
    }
    float cosPhi = std::cos(phi), sinPhi = std::sin(phi);
    float px = r * cosPhi, py = r * sinPhi;
    float z = std::sqrt(1.0f - px * px - py * py);
    if (z == 0) z = 1e-10f;
    return V3(px, py, z);
}

}  // namespace kzo
#endif
