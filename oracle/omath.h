// oracle/omath.h — TEST INFRASTRUCTURE ONLY (CPU oracle).  Not part of the product path.
//
// Scalar restatement of the arithmetic the reference gets from glam 0.15.1 (SSE2) through
// sub_crates/math3d, in the same operation order, for bit-level comparisons with the CUDA path.
// glam itself is not vendored under /root/reference (Cargo.lock: glam 0.15.1); the semantics below
// are restated from the crate and pinned by the reference's own math3d/lerp unit tests
// (tests/test_oracle_pins.py).  Compile with -ffp-contract=off.
//
// Follows: sub_crates/math3d/src/transform.rs:12-118, point.rs:132-141, vector.rs:44-48,141-150,
//          normal.rs:130-137, src/lerp.rs:20-129, src/fp_utils.rs:9-82.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>

#include "../include/psy_cuda.h"
#include "../include/psy_detmath.h"

namespace orc {

// The eight libm functions of the path.  Default: glibc, as the reference (Rust std) uses.  Strict mode
// (oracle_set_strict_math): the deterministic functions of include/psy_detmath.h, which the strict build of the
// CUDA path uses as well — then CPU and GPU compute every sample with the same bits.
// (plain globals, tested with an unlikely hint: the default path must cost what calling libm costs — this code is
// also the timed CPU baseline)
inline bool g_strict_math = false;
inline bool g_debug_trace = false;  // oracle_set_debug: per-sample trace of one path
#define ORC_UNLIKELY(x) __builtin_expect(!!(x), 0)
inline bool& strict_math() { return g_strict_math; }
inline float m_sinf(float x) { return ORC_UNLIKELY(g_strict_math) ? psy_det_sinf(x) : std::sin(x); }
inline float m_cosf(float x) { return ORC_UNLIKELY(g_strict_math) ? psy_det_cosf(x) : std::cos(x); }
inline float m_acosf(float x) { return ORC_UNLIKELY(g_strict_math) ? psy_det_acosf(x) : std::acos(x); }
inline float m_asinf(float x) { return ORC_UNLIKELY(g_strict_math) ? psy_det_asinf(x) : std::asin(x); }
inline double m_sin(double x) { return ORC_UNLIKELY(g_strict_math) ? psy_det_sin(x) : std::sin(x); }
inline double m_cos(double x) { return ORC_UNLIKELY(g_strict_math) ? psy_det_cos(x) : std::cos(x); }
inline double m_acos(double x) { return ORC_UNLIKELY(g_strict_math) ? psy_det_acos(x) : std::acos(x); }
inline double m_log(double x) { return ORC_UNLIKELY(g_strict_math) ? psy_det_log(x) : std::log(x); }

struct V3 {
    float x, y, z;
    float get(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length2(V3 a) { return dot(a, a); }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { return a / std::sqrt(dot(a, a)); }
inline V3 recip(V3 a) { return {1.0f / a.x, 1.0f / a.y, 1.0f / a.z}; }
// SSE minps / maxps: returns the second operand unless the first compares strictly smaller / larger
inline float sse_min(float a, float b) { return a < b ? a : b; }
inline float sse_max(float a, float b) { return a > b ? a : b; }
// glam Vec3A::signum (SSE2): +1 where v >= 0 (incl. -0.0), -1 elsewhere, NaN stays NaN
inline float sse_signum(float v) { return v != v ? v : (v >= 0.0f ? 1.0f : -1.0f); }
// Vector::abs = v * signum(v) (vector.rs:44-48)
inline V3 vabs(V3 a) { return {a.x * sse_signum(a.x), a.y * sse_signum(a.y), a.z * sse_signum(a.z)}; }
// Vec3A::max_element: max(max(x, y), z) with maxps semantics
inline float max_element(V3 a) { return sse_max(sse_max(a.x, a.y), a.z); }
// Rust f32::max / min (IEEE maxNum/minNum: a NaN operand is ignored)
inline float rmax(float a, float b) { return std::fmax(a, b); }
inline float rmin(float a, float b) { return std::fmin(a, b); }
inline double rmax(double a, double b) { return std::fmax(a, b); }
inline double rmin(double a, double b) { return std::fmin(a, b); }
// Rust f32::clamp
inline float rclamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
// Rust f32::signum
inline float rsignum(float v) { return v != v ? v : std::copysign(1.0f, v); }

struct Xf {
    V3 X{1, 0, 0}, Y{0, 1, 0}, Z{0, 0, 1}, T{0, 0, 0};
};
inline Xf from_abi(const PsyXform& p) {
    Xf m;
    m.X = mk(p.r[0][0], p.r[1][0], p.r[2][0]);
    m.Y = mk(p.r[0][1], p.r[1][1], p.r[2][1]);
    m.Z = mk(p.r[0][2], p.r[1][2], p.r[2][2]);
    m.T = mk(p.r[0][3], p.r[1][3], p.r[2][3]);
    return m;
}
inline V3 mat3_mul(const Xf& m, V3 v) { return (m.X * v.x + m.Y * v.y) + m.Z * v.z; }
inline V3 xf_point(const Xf& m, V3 p) { return mat3_mul(m, p) + m.T; }   // Point * Transform
inline V3 xf_vector(const Xf& m, V3 v) { return mat3_mul(m, v); }          // Vector * Transform
inline Xf affine_mul(const Xf& a, const Xf& b) {
    Xf r;
    r.X = mat3_mul(a, b.X);
    r.Y = mat3_mul(a, b.Y);
    r.Z = mat3_mul(a, b.Z);
    r.T = mat3_mul(a, b.T) + a.T;
    return r;
}
// Transform::mul(self, other) = other.0 * self.0
inline Xf xf_mul(const Xf& self, const Xf& other) { return affine_mul(other, self); }
inline Xf xf_inverse(const Xf& m) {
    V3 t0 = cross(m.Y, m.Z), t1 = cross(m.Z, m.X), t2 = cross(m.X, m.Y);
    float det = dot(m.Z, t2);
    float inv_det = 1.0f / det;
    V3 c0 = t0 * inv_det, c1 = t1 * inv_det, c2 = t2 * inv_det;
    Xf r;
    r.X = mk(c0.x, c1.x, c2.x);
    r.Y = mk(c0.y, c1.y, c2.y);
    r.Z = mk(c0.z, c1.z, c2.z);
    r.T = -mat3_mul(r, m.T);
    return r;
}
// Normal * Transform = transpose(inverse(M3)) * n (normal.rs:130-137)
inline V3 xf_normal(const Xf& m, V3 n) {
    V3 t0 = cross(m.Y, m.Z), t1 = cross(m.Z, m.X), t2 = cross(m.X, m.Y);
    float det = dot(m.Z, t2);
    float inv_det = 1.0f / det;
    V3 c0 = t0 * inv_det, c1 = t1 * inv_det, c2 = t2 * inv_det;
    return (c0 * n.x + c1 * n.y) + c2 * n.z;
}
inline Xf xf_scale(const Xf& m, float s) { return Xf{m.X * s, m.Y * s, m.Z * s, m.T * s}; }
inline Xf xf_add(const Xf& a, const Xf& b) { return Xf{a.X + b.X, a.Y + b.Y, a.Z + b.Z, a.T + b.T}; }

inline float lerp(float a, float b, float alpha) { return (a * (1.0f - alpha)) + (b * alpha); }
inline V3 lerp(V3 a, V3 b, float alpha) { return (a * (1.0f - alpha)) + (b * alpha); }
inline Xf lerp(const Xf& a, const Xf& b, float alpha) { return xf_add(xf_scale(a, 1.0f - alpha), xf_scale(b, alpha)); }

PsyColor lerp(const PsyColor& x, const PsyColor& y, float alpha);  // oshade.h

template <typename T>
inline T lerp_slice(const T* s, size_t n, float alpha) {
    if (n == 1 || alpha == 1.0f) return s[n - 1];
    float tmp = alpha * (float)(n - 1);
    size_t i1 = (size_t)tmp;
    float alpha2 = tmp - (float)i1;
    return lerp(s[i1], s[i1 + 1], alpha2);
}
// lerp_slice over a strided source converted on the fly
template <typename T, typename Get>
inline T lerp_slice_fn(size_t n, float alpha, Get get) {
    if (n == 1 || alpha == 1.0f) return get(n - 1);
    float tmp = alpha * (float)(n - 1);
    size_t i1 = (size_t)tmp;
    float alpha2 = tmp - (float)i1;
    return lerp(get(i1), get(i1 + 1), alpha2);
}

// fp_gamma (fp_utils.rs:9-13)
inline float fp_gamma(uint32_t n) {
    const float e = std::numeric_limits<float>::epsilon() * 0.5f;
    return (e * (float)n) / (1.0f - (e * (float)n));
}
inline uint32_t f2u(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline float increment_ulp(float v) {
    if (std::isfinite(v)) {
        if (v > 0.0f) return u2f(f2u(v) + 1);
        if (v < -0.0f) return u2f(f2u(v) - 1);
        return u2f(0x00000001u);
    }
    return v;
}
inline float decrement_ulp(float v) {
    if (std::isfinite(v)) {
        if (v > 0.0f) return u2f(f2u(v) - 1);
        if (v < -0.0f) return u2f(f2u(v) + 1);
        return u2f(0x80000001u);
    }
    return v;
}
// robust_ray_origin (fp_utils.rs:45-82)
inline V3 robust_ray_origin(V3 pos, float pos_err, V3 nor, V3 ray_dir) {
    V3 n = dot(nor, ray_dir) >= 0.0f ? nor : -nor;
    float d = dot(vabs(n), mk(pos_err, pos_err, pos_err));
    V3 offset = n * d;
    V3 p = pos + offset;
    float x = n.x >= 0.0f ? increment_ulp(p.x) : decrement_ulp(p.x);
    float y = n.y >= 0.0f ? increment_ulp(p.y) : decrement_ulp(p.y);
    float z = n.z >= 0.0f ? increment_ulp(p.z) : decrement_ulp(p.z);
    return mk(x, y, z);
}

// coordinate_system_from_vector (math.rs:26-34): returns (v, v2, v3)
inline void coordinate_system_from_vector(V3 v, V3& v2, V3& v3) {
    float sign = rsignum(v.z);
    float a = -1.0f / (sign + v.z);
    float b = v.x * v.y * a;
    v2 = mk(1.0f + sign * v.x * v.x * a, sign * b, -sign * v.x);
    v3 = mk(b, sign + v.y * v.y * a, -v.y);
}
// zup_to_vec (math.rs:46-53)
inline V3 zup_to_vec(V3 from, V3 toz) {
    V3 tz = normalized(toz), tx, ty;
    coordinate_system_from_vector(tz, tx, ty);
    return (tx * from.x) + (ty * from.y) + (tz * from.z);
}

}  // namespace orc
