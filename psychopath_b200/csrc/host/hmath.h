// hmath.h — host-side arithmetic for the scene builders, in the operation order of glam 0.15.1 /
// math3d so that bounds, instance boxes and composed transforms come out as the reference's would
// (SURVEY.md §8c "third-party arithmetic").  Must be compiled with -ffp-contract=off.
//
// Reference: sub_crates/math3d/src/{point,vector,normal,transform}.rs, src/bbox.rs, src/lerp.rs.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

#include "../../../include/psy_cuda.h"

namespace psyh {

struct V3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// glam Vec3A::dot: (x*x' + y*y') + z*z'
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length2(V3 a) { return dot(a, a); }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { return a / std::sqrt(dot(a, a)); }
// SSE minps/maxps semantics: second operand unless first strictly smaller/larger.
inline float sse_min(float a, float b) { return a < b ? a : b; }
inline float sse_max(float a, float b) { return a > b ? a : b; }
inline V3 vmin(V3 a, V3 b) { return {sse_min(a.x, b.x), sse_min(a.y, b.y), sse_min(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {sse_max(a.x, b.x), sse_max(a.y, b.y), sse_max(a.z, b.z)}; }

// Affine transform held as glam's column axes + translation.
struct Xf {
    V3 X{1, 0, 0}, Y{0, 1, 0}, Z{0, 0, 1}, T{0, 0, 0};
};

inline V3 mat3_mul(const Xf& m, V3 v) { return (m.X * v.x + m.Y * v.y) + m.Z * v.z; }
// Point * Transform (point.rs:132-141)
inline V3 xf_point(const Xf& m, V3 p) { return mat3_mul(m, p) + m.T; }
// Vector * Transform (vector.rs:141-150)
inline V3 xf_vector(const Xf& m, V3 v) { return mat3_mul(m, v); }
// Affine3A a*b = {a.M*b.M, a.M*b.T + a.T}
inline Xf affine_mul(const Xf& a, const Xf& b) {
    Xf r;
    r.X = mat3_mul(a, b.X);
    r.Y = mat3_mul(a, b.Y);
    r.Z = mat3_mul(a, b.Z);
    r.T = mat3_mul(a, b.T) + a.T;
    return r;
}
// math3d Transform::mul(self, other) = other.0 * self.0 (transform.rs:89-96)
inline Xf xf_mul(const Xf& self, const Xf& other) { return affine_mul(other, self); }
// Affine3A::inverse
inline Xf xf_inverse(const Xf& m) {
    V3 t0 = cross(m.Y, m.Z), t1 = cross(m.Z, m.X), t2 = cross(m.X, m.Y);
    float det = dot(m.Z, t2);
    float inv_det = 1.0f / det;
    V3 c0 = t0 * inv_det, c1 = t1 * inv_det, c2 = t2 * inv_det;
    Xf r;
    // transpose of columns (c0,c1,c2)
    r.X = {c0.x, c1.x, c2.x};
    r.Y = {c0.y, c1.y, c2.y};
    r.Z = {c0.z, c1.z, c2.z};
    r.T = -mat3_mul(r, m.T);
    return r;
}
inline Xf xf_scale(const Xf& m, float s) { return Xf{m.X * s, m.Y * s, m.Z * s, m.T * s}; }
inline Xf xf_add(const Xf& a, const Xf& b) { return Xf{a.X + b.X, a.Y + b.Y, a.Z + b.Z, a.T + b.T}; }

inline PsyXform to_abi(const Xf& m) {
    PsyXform o;
    const float* c[4] = {&m.X.x, &m.Y.x, &m.Z.x, &m.T.x};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) o.r[i][j] = c[j][i];
    return o;
}

struct BBox {
    V3 min{std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity(),
           std::numeric_limits<float>::infinity()};
    V3 max{-std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(),
           -std::numeric_limits<float>::infinity()};
};
inline BBox bb_union(const BBox& a, const BBox& b) { return BBox{vmin(a.min, b.min), vmax(a.max, b.max)}; }
inline BBox bb_union_pt(const BBox& a, V3 p) { return BBox{vmin(a.min, p), vmax(a.max, p)}; }
inline float bb_area(const BBox& b) {
    V3 d = b.max - b.min;
    return ((d.x * d.y) + (d.y * d.z) + (d.z * d.x)) * 2.0f;
}
inline float bb_diagonal(const BBox& b) { return length(b.max - b.min); }

// Lerp impls (src/lerp.rs:57-129): a*(1-alpha) + b*alpha
inline float lerp(float a, float b, float alpha) { return (a * (1.0f - alpha)) + (b * alpha); }
inline V3 lerp(V3 a, V3 b, float alpha) { return (a * (1.0f - alpha)) + (b * alpha); }
inline BBox lerp(const BBox& a, const BBox& b, float alpha) {
    return BBox{lerp(a.min, b.min, alpha), lerp(a.max, b.max, alpha)};
}
inline Xf lerp(const Xf& a, const Xf& b, float alpha) {
    return xf_add(xf_scale(a, 1.0f - alpha), xf_scale(b, alpha));
}
inline V3 bb_center(const BBox& b) { return lerp(b.min, b.max, 0.5f); }

// lerp_slice (src/lerp.rs:20-35)
template <typename T>
inline T lerp_slice(const T* s, size_t n, float alpha) {
    if (n == 1 || alpha == 1.0f) return s[n - 1];
    float tmp = alpha * (float)(n - 1);
    size_t i1 = (size_t)tmp;
    float alpha2 = tmp - (float)i1;
    return lerp(s[i1], s[i1 + 1], alpha2);
}

// BBox::transformed (src/bbox.rs:58-80)
inline BBox bb_transformed(const BBox& b, const Xf& xf) {
    V3 vs[8] = {{b.min.x, b.min.y, b.min.z}, {b.min.x, b.min.y, b.max.z}, {b.min.x, b.max.y, b.min.z},
                {b.min.x, b.max.y, b.max.z}, {b.max.x, b.min.y, b.min.z}, {b.max.x, b.min.y, b.max.z},
                {b.max.x, b.max.y, b.min.z}, {b.max.x, b.max.y, b.max.z}};
    BBox o;
    for (auto& v0 : vs) {
        V3 v = xf_point(xf, v0);
        o.min = vmin(v, o.min);
        o.max = vmax(v, o.max);
    }
    return o;
}

// transform_bbox_slice_from (src/bbox.rs:153-174)
inline void transform_bbox_slice_from(const std::vector<BBox>& in, const Xf* xforms, size_t nx,
                                      std::vector<BBox>& out) {
    out.clear();
    if (nx == 0) {
        out = in;
    } else if (in.size() == nx) {
        for (size_t i = 0; i < nx; ++i) out.push_back(bb_transformed(in[i], xf_inverse(xforms[i])));
    } else if (in.size() > nx) {
        float s = (float)(in.size() - 1);
        for (size_t i = 0; i < in.size(); ++i)
            out.push_back(bb_transformed(in[i], xf_inverse(lerp_slice(xforms, nx, (float)i / s))));
    } else {
        float s = (float)(nx - 1);
        for (size_t i = 0; i < nx; ++i)
            out.push_back(bb_transformed(lerp_slice(in.data(), in.size(), (float)i / s), xf_inverse(xforms[i])));
    }
}

// merge_slices_append (src/algorithm.rs:234-262)
template <typename T, typename F>
inline void merge_slices_append(const T* s1, size_t n1, const T* s2, size_t n2, std::vector<T>& out, F merge) {
    if (n1 == 0 || n2 == 0) return;
    if (n1 == n2) {
        for (size_t i = 0; i < n1; ++i) out.push_back(merge(s1[i], s2[i]));
    } else if (n1 > n2) {
        float s = (float)(n1 - 1);
        for (size_t i = 0; i < n1; ++i) out.push_back(merge(s1[i], lerp_slice(s2, n2, (float)i / s)));
    } else {
        float s = (float)(n2 - 1);
        for (size_t i = 0; i < n2; ++i) out.push_back(merge(lerp_slice(s1, n1, (float)i / s), s2[i]));
    }
}

inline PsyBBox to_abi(const BBox& b) {
    return PsyBBox{{b.min.x, b.min.y, b.min.z}, {b.max.x, b.max.y, b.max.z}};
}

// hash_u64 (src/hash.rs:12-22)
inline uint64_t hash_u64(uint64_t n, uint64_t seed) {
    uint64_t h = n;
    for (int i = 0; i < 4; ++i) {
        h *= (uint64_t)32416190071ull * (uint64_t)314604959ull;
        h ^= h >> 32;
        h ^= seed;
    }
    return h;
}

}  // namespace psyh
