// dmath.cuh — device arithmetic of the render core.  Every expression keeps the operation order
// of the reference's glam/SSE2 math (sub_crates/math3d, src/lerp.rs, src/fp_utils.rs); the file is
// compiled with -fmad=false and IEEE div/sqrt so that box tests, lerps, 3x4 inverses and the
// watertight triangle test produce the bits the CPU produces.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define PSY_DEV __device__ __forceinline__
// Large, rarely-hot routines (double-precision trig, spectral lookups): one copy per kernel keeps the
// shade kernels' instruction footprint down (ncu: stall_no_instruction dominated k_shade before).
#define PSY_NOINLINE __device__ __noinline__

namespace psyd {

struct V3 {
    float x, y, z;
};
PSY_DEV V3 mk3(float x, float y, float z) {
    V3 v;
    v.x = x, v.y = y, v.z = z;
    return v;
}
PSY_DEV V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PSY_DEV V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PSY_DEV V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
PSY_DEV V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PSY_DEV V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
PSY_DEV float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
PSY_DEV V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PSY_DEV float length2(V3 a) { return dot(a, a); }
PSY_DEV float length(V3 a) { return sqrtf(dot(a, a)); }
PSY_DEV V3 normalized(V3 a) { return a / sqrtf(dot(a, a)); }
PSY_DEV V3 recip3(V3 a) { return mk3(1.0f / a.x, 1.0f / a.y, 1.0f / a.z); }
PSY_DEV float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// SSE minps/maxps semantics (second operand unless the first is strictly smaller/larger)
PSY_DEV float sse_min(float a, float b) { return a < b ? a : b; }
PSY_DEV float sse_max(float a, float b) { return a > b ? a : b; }
PSY_DEV float sse_signum(float v) { return v != v ? v : (v >= 0.0f ? 1.0f : -1.0f); }
PSY_DEV V3 vabs(V3 a) { return mk3(a.x * sse_signum(a.x), a.y * sse_signum(a.y), a.z * sse_signum(a.z)); }
PSY_DEV float max_element(V3 a) { return sse_max(sse_max(a.x, a.y), a.z); }
// Rust f32::max/min (maxNum/minNum) == CUDA fmaxf/fminf
PSY_DEV float rmax(float a, float b) { return fmaxf(a, b); }
PSY_DEV float rmin(float a, float b) { return fminf(a, b); }
PSY_DEV double rmax(double a, double b) { return fmax(a, b); }
PSY_DEV double rmin(double a, double b) { return fmin(a, b); }
PSY_DEV float rclamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
PSY_DEV float rsignum(float v) { return v != v ? v : copysignf(1.0f, v); }

// 3x4 affine transform as glam's column axes + translation
struct Xf {
    V3 X, Y, Z, T;
};
PSY_DEV Xf xf_identity() {
    Xf m;
    m.X = mk3(1, 0, 0), m.Y = mk3(0, 1, 0), m.Z = mk3(0, 0, 1), m.T = mk3(0, 0, 0);
    return m;
}
// rows as stored on the device: r_i = (X_i, Y_i, Z_i, T_i)
PSY_DEV Xf xf_from_rows(float4 r0, float4 r1, float4 r2) {
    Xf m;
    m.X = mk3(r0.x, r1.x, r2.x);
    m.Y = mk3(r0.y, r1.y, r2.y);
    m.Z = mk3(r0.z, r1.z, r2.z);
    m.T = mk3(r0.w, r1.w, r2.w);
    return m;
}
PSY_DEV Xf xf_load(const float4* __restrict__ p) { return xf_from_rows(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
PSY_DEV V3 mat3_mul(const Xf& m, V3 v) { return (m.X * v.x + m.Y * v.y) + m.Z * v.z; }
PSY_DEV V3 xf_point(const Xf& m, V3 p) { return mat3_mul(m, p) + m.T; }
PSY_DEV V3 xf_vector(const Xf& m, V3 v) { return mat3_mul(m, v); }
PSY_DEV Xf affine_mul(const Xf& a, const Xf& b) {
    Xf r;
    r.X = mat3_mul(a, b.X);
    r.Y = mat3_mul(a, b.Y);
    r.Z = mat3_mul(a, b.Z);
    r.T = mat3_mul(a, b.T) + a.T;
    return r;
}
PSY_DEV Xf xf_mul(const Xf& self, const Xf& other) { return affine_mul(other, self); }
PSY_DEV Xf xf_inverse(const Xf& m) {
    V3 t0 = cross(m.Y, m.Z), t1 = cross(m.Z, m.X), t2 = cross(m.X, m.Y);
    float det = dot(m.Z, t2);
    float inv_det = 1.0f / det;
    V3 c0 = t0 * inv_det, c1 = t1 * inv_det, c2 = t2 * inv_det;
    Xf r;
    r.X = mk3(c0.x, c1.x, c2.x);
    r.Y = mk3(c0.y, c1.y, c2.y);
    r.Z = mk3(c0.z, c1.z, c2.z);
    r.T = -mat3_mul(r, m.T);
    return r;
}
PSY_DEV V3 xf_normal(const Xf& m, V3 n) {
    V3 t0 = cross(m.Y, m.Z), t1 = cross(m.Z, m.X), t2 = cross(m.X, m.Y);
    float det = dot(m.Z, t2);
    float inv_det = 1.0f / det;
    V3 c0 = t0 * inv_det, c1 = t1 * inv_det, c2 = t2 * inv_det;
    return (c0 * n.x + c1 * n.y) + c2 * n.z;
}
PSY_DEV float lerpf(float a, float b, float alpha) { return (a * (1.0f - alpha)) + (b * alpha); }
PSY_DEV V3 lerp3(V3 a, V3 b, float alpha) { return (a * (1.0f - alpha)) + (b * alpha); }
PSY_DEV Xf xf_lerp(const Xf& a, const Xf& b, float alpha) {
    float ia = 1.0f - alpha;
    Xf r;
    r.X = (a.X * ia) + (b.X * alpha);
    r.Y = (a.Y * ia) + (b.Y * alpha);
    r.Z = (a.Z * ia) + (b.Z * alpha);
    r.T = (a.T * ia) + (b.T * alpha);
    return r;
}

// lerp_slice index arithmetic (src/lerp.rs:20-35): returns true when two keys must be blended
PSY_DEV bool lerp_slice_keys(uint32_t n, float alpha, uint32_t& i1, float& alpha2) {
    if (n == 1 || alpha == 1.0f) {
        i1 = n - 1;
        alpha2 = 0.0f;
        return false;
    }
    float tmp = alpha * (float)(n - 1);
    i1 = (uint32_t)tmp;
    alpha2 = tmp - (float)i1;
    return true;
}
PSY_DEV float lerp_slice_f(const float* __restrict__ s, uint32_t n, float alpha) {
    uint32_t i1;
    float a2;
    if (!lerp_slice_keys(n, alpha, i1, a2)) return __ldg(s + i1);
    return lerpf(__ldg(s + i1), __ldg(s + i1 + 1), a2);
}
// transforms stored as 3 float4 rows each
PSY_DEV Xf lerp_slice_xf(const float4* __restrict__ s, uint32_t n, float alpha) {
    uint32_t i1;
    float a2;
    if (!lerp_slice_keys(n, alpha, i1, a2)) return xf_load(s + 3 * i1);
    return xf_lerp(xf_load(s + 3 * i1), xf_load(s + 3 * (i1 + 1)), a2);
}

PSY_DEV float fp_gamma(uint32_t n) {
    const float e = 1.1920928955078125e-07f * 0.5f;
    return (e * (float)n) / (1.0f - (e * (float)n));
}
PSY_DEV float increment_ulp(float v) {
    if (isfinite(v)) {
        if (v > 0.0f) return __uint_as_float(__float_as_uint(v) + 1u);
        if (v < -0.0f) return __uint_as_float(__float_as_uint(v) - 1u);
        return __uint_as_float(0x00000001u);
    }
    return v;
}
PSY_DEV float decrement_ulp(float v) {
    if (isfinite(v)) {
        if (v > 0.0f) return __uint_as_float(__float_as_uint(v) - 1u);
        if (v < -0.0f) return __uint_as_float(__float_as_uint(v) + 1u);
        return __uint_as_float(0x80000001u);
    }
    return v;
}
// robust_ray_origin (src/fp_utils.rs:45-82)
PSY_DEV V3 robust_ray_origin(V3 pos, float pos_err, V3 nor, V3 ray_dir) {
    V3 n = dot(nor, ray_dir) >= 0.0f ? nor : -nor;
    float d = dot(vabs(n), mk3(pos_err, pos_err, pos_err));
    V3 offset = n * d;
    V3 p = pos + offset;
    float x = n.x >= 0.0f ? increment_ulp(p.x) : decrement_ulp(p.x);
    float y = n.y >= 0.0f ? increment_ulp(p.y) : decrement_ulp(p.y);
    float z = n.z >= 0.0f ? increment_ulp(p.z) : decrement_ulp(p.z);
    return mk3(x, y, z);
}
// coordinate_system_from_vector / zup_to_vec (src/math.rs:26-53)
PSY_DEV void coordinate_system_from_vector(V3 v, V3& v2, V3& v3) {
    float sign = rsignum(v.z);
    float a = -1.0f / (sign + v.z);
    float b = v.x * v.y * a;
    v2 = mk3(1.0f + sign * v.x * v.x * a, sign * b, -sign * v.x);
    v3 = mk3(b, sign + v.y * v.y * a, -v.y);
}
PSY_DEV V3 zup_to_vec(V3 from, V3 toz) {
    V3 tz = normalized(toz), tx, ty;
    coordinate_system_from_vector(tz, tx, ty);
    return (tx * from.x) + (ty * from.y) + (tz * from.z);
}

}  // namespace psyd
