// dshade.cuh — sampling, hero-wavelength spectral colour, closures, light sampling and the
// light-tree walk, as device functions.
//
// Replaces (file:line in the reference):
//   src/renderer.rs:698-737       get_sample_4d (sobol_burley::sample_4d), golden_ratio_sample
//   src/hash.rs:1-29, src/math.rs:67-132 (probit), src/mis.rs:12-21
//   src/sampling/monte_carlo.rs:10-199
//   src/color.rs:15-43,84-142,336-383,520-526,602-628; spectral_upsampling/src/meng.rs:177-328
//   src/shading/surface_closure.rs:57-155,281-633,714-719
//   src/light/sphere_light.rs:51-192, rectangle_light.rs:52-243, distant_disk_light.rs:59-85
//   src/scene/mod.rs:28-99, src/scene/assembly.rs:46-140, src/accel/light_tree.rs:135-201
#pragma once
#include "dtrace.cuh"
#include "../../../include/psy_detmath.h"

// The eight libm functions of the path.  The product build calls CUDA's libdevice; the strict build
// (-DPSY_STRICT=1, libpsy_cuda_strict.so) calls the deterministic functions the oracle's strict mode uses as well,
// so that the CUDA path and the restated reference path agree bit for bit (DESIGN.md "Strict mode").
#ifndef PSY_STRICT
#define PSY_STRICT 0
#endif
#if PSY_STRICT
#define PSY_SINF(x) psy_det_sinf(x)
#define PSY_COSF(x) psy_det_cosf(x)
#define PSY_ACOSF(x) psy_det_acosf(x)
#define PSY_ASINF(x) psy_det_asinf(x)
#define PSY_SIN(x) psy_det_sin(x)
#define PSY_COS(x) psy_det_cos(x)
#define PSY_ACOS(x) psy_det_acos(x)
#define PSY_LOG(x) psy_det_log(x)
#else
#define PSY_SINF(x) sinf(x)
#define PSY_COSF(x) cosf(x)
#define PSY_ACOSF(x) acosf(x)
#define PSY_ASINF(x) asinf(x)
#define PSY_SIN(x) sin(x)
#define PSY_COS(x) cos(x)
#define PSY_ACOS(x) acos(x)
#define PSY_LOG(x) log(x)
#endif

namespace psyd {

__constant__ unsigned short c_sobol_rev[2][16][4];
__constant__ short c_meng_grid[168][8];
__constant__ float c_meng_uv[186][2];
__constant__ float c_meng_mat[6];

struct V4 {
    float v[4];
};
PSY_DEV V4 splat4(float a) {
    V4 r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = a;
    return r;
}
PSY_DEV V4 operator*(V4 a, float s) {
    V4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] * s;
    return r;
}
PSY_DEV V4 operator*(V4 a, V4 b) {
    V4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] * b.v[i];
    return r;
}
PSY_DEV V4 operator/(V4 a, float s) {
    V4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] / s;
    return r;
}
PSY_DEV V4 operator+(V4 a, V4 b) {
    V4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] + b.v[i];
    return r;
}
PSY_DEV float max_element4(V4 a) { return sse_max(sse_max(a.v[0], a.v[2]), sse_max(a.v[1], a.v[3])); }
PSY_DEV V4 from_f4(float4 f) {
    V4 r;
    r.v[0] = f.x, r.v[1] = f.y, r.v[2] = f.z, r.v[3] = f.w;
    return r;
}
PSY_DEV float4 to_f4(V4 a) { return make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }

//---- hashing and the sampler ---------------------------------------------------------------------
PSY_DEV uint32_t hash_u32(uint32_t n, uint32_t seed) {
    uint32_t h = n;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        h *= 0x736caf6fu;
        h ^= h >> 16;
        h ^= seed;
    }
    return h;
}
PSY_DEV float hash_u32_to_f32(uint32_t n, uint32_t seed) {
    const float INV_MAX = 1.0f / 4294967295.0f;
    return (float)hash_u32(n, seed) * INV_MAX;
}
PSY_DEV uint32_t sb_hash(uint32_t n) {
    uint32_t h = n ^ 0x79c68e4au;
    h ^= h >> 16;
    h *= 0x7feb352du;
    h ^= h >> 15;
    h *= 0x846ca68bu;
    h ^= h >> 16;
    return h;
}
PSY_DEV uint32_t sb_lk_scramble(uint32_t n, uint32_t hashed_seed) {
    n ^= n * 0x3d20adeau;
    n += hashed_seed;
    n *= (hashed_seed >> 16) | 1u;
    n ^= n * 0x05526c56u;
    n ^= n * 0x53a22864u;
    return n;
}
// Owen-scrambled Sobol, 4 dimensions at once (sobol_burley::sample_4d; see oracle/osample.h for the
// provenance note — the integer recipe is shared with the oracle bit for bit)
PSY_DEV void sobol_sample_4d(uint32_t index, uint32_t dimension_set, uint32_t seed, float out[4]) {
    uint32_t shuffled_rev = sb_lk_scramble(__brev(index), sb_hash(seed));
    uint32_t sob[4] = {0u, 0u, 0u, 0u};
    uint32_t top = shuffled_rev >> 16;
#pragma unroll
    for (int bit = 0; bit < 16; ++bit) {
        if (top & (0x8000u >> bit)) {
#pragma unroll
            for (int l = 0; l < 4; ++l) sob[l] ^= (uint32_t)c_sobol_rev[dimension_set][bit][l];
        }
    }
    const uint32_t LANE_SALT[4] = {0x912f69bau, 0x174f18abu, 0x691e72cau, 0xb40cc1b8u};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        uint32_t lane_seed = (dimension_set ^ seed) ^ LANE_SALT[l];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            lane_seed *= 0x736caf6fu;
            lane_seed ^= lane_seed >> 16;
        }
        uint32_t v = __brev(sb_lk_scramble(sob[l], lane_seed));
        out[l] = __uint_as_float((v >> 9) | 0x3f800000u) - 1.0f;  // sobol_burley's to_norm_floats (23-bit mantissa trick)
    }
}
// get_sample_4d (renderer.rs:698-725)
PSY_NOINLINE void get_sample_4d(uint32_t i, uint32_t dimension_set, uint32_t px, uint32_t py, uint32_t seed, float out[4]) {
    uint32_t s = px ^ (py << 16) ^ (seed * 0x736caf6fu);
    if (dimension_set < 2u) {
        sobol_sample_4d(i, dimension_set, s, out);
    } else {
        for (uint32_t k = 0; k < 4; ++k) out[k] = hash_u32_to_f32((dimension_set * 4u + k) ^ (i << 16), s);
    }
}
PSY_DEV float golden_ratio_sample(uint32_t i, uint32_t scramble) {
    uint32_t n = (i + hash_u32(scramble, 0u)) * 2654435769u;
    return (float)n * (1.0f / 4294967296.0f);
}
// probit (math.rs:67-132), f64 inside like the reference
PSY_DEV float probit(float nf, float width) {
    double n = (double)nf;
    const double A1 = -3.969683028665376e+01, A2 = 2.209460984245205e+02, A3 = -2.759285104469687e+02,
                 A4 = 1.383577518672690e+02, A5 = -3.066479806614716e+01, A6 = 2.506628277459239e+00;
    const double B1 = -5.447609879822406e+01, B2 = 1.615858368580409e+02, B3 = -1.556989798598866e+02,
                 B4 = 6.680131188771972e+01, B5 = -1.328068155288572e+01;
    const double C1 = -7.784894002430293e-03, C2 = -3.223964580411365e-01, C3 = -2.400758277161838e+00,
                 C4 = -2.549732539343734e+00, C5 = 4.374664141464968e+00, C6 = 2.938163982698783e+00;
    const double D1 = 7.784695709041462e-03, D2 = 3.224671290700398e-01, D3 = 2.445134137142996e+00,
                 D4 = 3.754408661907416e+00;
    const double N_LOW = 0.02425, N_HIGH = 1.0 - N_LOW;
    double x;
    if (0.0 < n && n < N_LOW) {
        double q = sqrt(-2.0 * PSY_LOG(n));
        x = (((((C1 * q + C2) * q + C3) * q + C4) * q + C5) * q + C6) / ((((D1 * q + D2) * q + D3) * q + D4) * q + 1.0);
    } else if (n <= N_HIGH) {
        double q = n - 0.5;
        double r = q * q;
        x = (((((A1 * r + A2) * r + A3) * r + A4) * r + A5) * r + A6) * q /
            (((((B1 * r + B2) * r + B3) * r + B4) * r + B5) * r + 1.0);
    } else if (n < 1.0) {
        double q = sqrt(-2.0 * PSY_LOG(1.0 - n));
        x = -(((((C1 * q + C2) * q + C3) * q + C4) * q + C5) * q + C6) / ((((D1 * q + D2) * q + D3) * q + D4) * q + 1.0);
    } else if (n == 1.0) {
        x = CUDART_INF;
    } else {
        x = CUDART_NAN;
    }
    return (float)x * width;
}
PSY_DEV float power_heuristic(float a, float b) {
    if (isinf(a)) return a;
    float a2 = a * a, b2 = b * b;
    float mis_fac = a2 / (a2 + b2);
    return a / mis_fac;
}

//---- Monte-Carlo warps (monte_carlo.rs) ---------------------------------------------------------
PSY_NOINLINE void square_to_circle(float x, float y, float& ox, float& oy) {
    const float QPI = 0.78539816339744830962f;
    if (x == 0.0f && y == 0.0f) {
        ox = 0.0f, oy = 0.0f;
        return;
    }
    float radius, angle;
    if (x > fabsf(y)) {
        radius = x;
        angle = QPI * (y / x);
    } else if (y > fabsf(x)) {
        radius = y;
        angle = QPI * (2.0f - (x / y));
    } else if (x < -fabsf(y)) {
        radius = -x;
        angle = QPI * (4.0f + (y / x));
    } else {
        radius = -y;
        angle = QPI * (6.0f - (x / y));
    }
    ox = radius * PSY_COSF(angle);
    oy = radius * PSY_SINF(angle);
}
PSY_DEV V3 cosine_sample_hemisphere(float u, float v) {
    float a, b;
    square_to_circle((u * 2.0f) - 1.0f, (v * 2.0f) - 1.0f, a, b);
    float z = sqrtf(rmax(1.0f - ((a * a) + (b * b)), 0.0f));
    return mk3(a, b, z);
}
#define PSY_PI_32 3.14159265358979323846f
#define PSY_PI_64 3.14159265358979323846
PSY_DEV V3 uniform_sample_sphere(float u, float v) {
    float z = 1.0f - (2.0f * u);
    float r = sqrtf(rmax(1.0f - (z * z), 0.0f));
    float phi = 2.0f * PSY_PI_32 * v;
    return mk3(r * PSY_COSF(phi), r * PSY_SINF(phi), z);
}
PSY_NOINLINE V3 uniform_sample_cone(float u, float v, double cos_theta_max) {
    double cos_theta = (1.0 - (double)u) + ((double)u * cos_theta_max);
    double sin_theta = sqrt(1.0 - (cos_theta * cos_theta));
    double phi = (double)v * 2.0 * PSY_PI_64;
    return mk3((float)(PSY_COS(phi) * sin_theta), (float)(PSY_SIN(phi) * sin_theta), (float)cos_theta);
}
PSY_DEV double uniform_sample_cone_pdf(double cos_theta_max) { return 1.0 / (2.0 * PSY_PI_64 * (1.0 - cos_theta_max)); }
PSY_DEV V3 uniform_sample_triangle(V3 va, V3 vb, V3 vc, float i, float j) {
    float isqrt = sqrtf(i);
    float a = 1.0f - isqrt;
    float b = isqrt * (1.0f - j);
    float c = j * isqrt;
    return (va * a) + (vb * b) + (vc * c);
}
PSY_DEV float triangle_surface_area(V3 p0, V3 p1, V3 p2) { return 0.5f * length(cross(p1 - p0, p2 - p0)); }

struct SphTri {
    double cos_a, cos_b, cos_c, sin_a, sin_b, sin_c;
};
PSY_DEV SphTri sph_tri_edges(V3 va, V3 vb, V3 vc) {
    SphTri s;
    s.cos_a = (double)rmin(rmax(dot(vb, vc), -1.0f), 1.0f);
    s.cos_b = (double)rmin(rmax(dot(vc, va), -1.0f), 1.0f);
    s.cos_c = (double)rmin(rmax(dot(va, vb), -1.0f), 1.0f);
    s.sin_a = sqrt(1.0 - (s.cos_a * s.cos_a));
    s.sin_b = sqrt(1.0 - (s.cos_b * s.cos_b));
    s.sin_c = sqrt(1.0 - (s.cos_c * s.cos_c));
    return s;
}
PSY_NOINLINE float spherical_triangle_solid_angle(V3 va, V3 vb, V3 vc) {
    SphTri s = sph_tri_edges(va, vb, vc);
    if (s.cos_a == 1.0 || s.cos_b == 1.0 || s.cos_c == 1.0) return 0.0f;
    double cos_va = rmin(rmax((s.cos_a - (s.cos_b * s.cos_c)) / (s.sin_b * s.sin_c), -1.0), 1.0);
    double cos_vb = rmin(rmax((s.cos_b - (s.cos_c * s.cos_a)) / (s.sin_c * s.sin_a), -1.0), 1.0);
    double cos_vc = rmin(rmax((s.cos_c - (s.cos_a * s.cos_b)) / (s.sin_a * s.sin_b), -1.0), 1.0);
    double ang_va = PSY_ACOS(cos_va), ang_vb = PSY_ACOS(cos_vb), ang_vc = PSY_ACOS(cos_vc);
    return (float)(ang_va + ang_vb + ang_vc - PSY_PI_64);
}
PSY_NOINLINE V3 uniform_sample_spherical_triangle(V3 va, V3 vb, V3 vc, float i, float j) {
    SphTri s = sph_tri_edges(va, vb, vc);
    if (s.cos_a == 1.0 || s.cos_b == 1.0 || s.cos_c == 1.0) return va;
    double cos_va = rmin(rmax((s.cos_a - (s.cos_b * s.cos_c)) / (s.sin_b * s.sin_c), -1.0), 1.0);
    double cos_vb = rmin(rmax((s.cos_b - (s.cos_c * s.cos_a)) / (s.sin_c * s.sin_a), -1.0), 1.0);
    double cos_vc = rmin(rmax((s.cos_c - (s.cos_a * s.cos_b)) / (s.sin_a * s.sin_b), -1.0), 1.0);
    double sin_va = sqrt(1.0 - (cos_va * cos_va));
    double ang_va = PSY_ACOS(cos_va), ang_vb = PSY_ACOS(cos_vb), ang_vc = PSY_ACOS(cos_vc);
    double area = ang_va + ang_vb + ang_vc - PSY_PI_64;
    double area_2 = area * (double)i;
    double ss = PSY_SIN(area_2 - ang_va);
    double t = PSY_COS(area_2 - ang_va);
    double u = t - cos_va;
    double v = ss + (sin_va * s.cos_c);
    double q_top = (((v * t) - (u * ss)) * cos_va) - v;
    double q_bottom = ((v * ss) + (u * t)) * sin_va;
    double q = q_top / q_bottom;
    V3 vc_2 = (va * (float)q) + (normalized(vc - (va * dot(vc, va))) * (float)sqrt(1.0 - (q * q)));
    float z = 1.0f - (j * (1.0f - dot(vc_2, vb)));
    return (vb * z) + (normalized(vc_2 - (vb * dot(vc_2, vb))) * sqrtf(1.0f - (z * z)));
}

//---- fastapprox + colour -------------------------------------------------------------------------
PSY_DEV float fast_pow2(float p) {
    float offset = (p < 0.0f) ? 1.0f : 0.0f;
    float clipp = (p < -126.0f) ? -126.0f : p;
    int w = (int)clipp;
    float z = clipp - (float)w + offset;
    uint32_t bits =
        (uint32_t)((float)(1 << 23) * (clipp + 121.2740575f + 27.7280233f / (4.84252568f - z) - 1.49012907f * z));
    return __uint_as_float(bits);
}
PSY_DEV float fast_exp(float p) { return fast_pow2(1.442695040f * p); }

#define PSY_WL_MIN 380.0f
#define PSY_WL_MAX 700.0f
#define PSY_WL_RANGE (PSY_WL_MAX - PSY_WL_MIN)
#define PSY_WL_RANGE_Q (PSY_WL_RANGE / 4.0f)
PSY_DEV float map_0_1_to_wavelength(float n) { return n * PSY_WL_RANGE + PSY_WL_MIN; }
PSY_DEV float nth_wavelength(float hero, int n) {
    float wl = hero + (PSY_WL_RANGE_Q * (float)n);
    return wl > PSY_WL_MAX ? wl - PSY_WL_RANGE : wl;
}

#define PSY_MENG_W 12
#define PSY_MENG_H 14
#define PSY_MENG_N 95
#define PSY_MENG_EER 0.009355121400914532f

// spectrum_xyz_to_p_4 (meng.rs:177-328)
PSY_DEV V4 spectrum_xyz_to_p_4(const DScene& S, const float lambdas[4], float x, float y, float z) {
    float inv_norm = x + y + z;
    float norm = 1.0f / inv_norm;
    if (!(norm < 3.402823466e+38f)) return splat4(0.0f);
    float xyy0 = x * norm, xyy1 = y * norm;
    float uv0 = c_meng_mat[0] * xyy0 + c_meng_mat[1] * xyy1 + c_meng_mat[2];
    float uv1 = c_meng_mat[3] * xyy0 + c_meng_mat[4] * xyy1 + c_meng_mat[5];
    if (uv0 < 0.0f || uv0 >= (float)PSY_MENG_W || uv1 < 0.0f || uv1 >= (float)PSY_MENG_H) return splat4(0.0f);
    int uvi0 = (int)uv0, uvi1 = (int)uv1;
    int cell_idx = uvi0 + PSY_MENG_W * uvi1;
    const short* cell = c_meng_grid[cell_idx];
    bool inside = cell[0] != 0;
    int num = cell[1];
    if (num == 0) return splat4(0.0f);
    int idx[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) idx[i] = cell[2 + i];

    float sbf[4];
    int sb0[4], sb1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float sb = (lambdas[k] - 360.0f) / (830.0f - 360.0f) * ((float)PSY_MENG_N - 1.0f);
        sb0[k] = (int)sb;
        sb1[k] = sb0[k] + 1 < PSY_MENG_N - 1 ? sb0[k] + 1 : PSY_MENG_N - 1;
        sbf[k] = sb - (float)sb0[k];
    }
    V4 p[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        p[i] = splat4(0.0f);
        if (i < num) {
            const float* spectrum = S.meng_spectra + idx[i] * PSY_MENG_N;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                p[i].v[k] = __ldg(spectrum + sb0[k]) * (1.0f - sbf[k]) + __ldg(spectrum + sb1[k]) * sbf[k];
        }
    }
    V4 interpolated = splat4(0.0f);
    if (inside) {
        float u2 = uv0 - (float)uvi0, v2 = uv1 - (float)uvi1;
        interpolated = p[0] * ((1.0f - u2) * (1.0f - v2)) + p[2] * ((1.0f - u2) * v2) + p[3] * (u2 * v2) +
                       p[1] * (u2 * (1.0f - v2));
    } else {
        float ex = uv0 - c_meng_uv[idx[0]][0];
        float ey = uv1 - c_meng_uv[idx[0]][1];
        float e0x = c_meng_uv[idx[1]][0] - c_meng_uv[idx[0]][0];
        float e0y = c_meng_uv[idx[1]][1] - c_meng_uv[idx[0]][1];
        float uu = e0x * ey - ex * e0y;
        for (int i = 0; i < num - 1; ++i) {
            float e1x, e1y;
            int last = (i == num - 2) ? 1 : i + 2;
            e1x = c_meng_uv[idx[last]][0] - c_meng_uv[idx[0]][0];
            e1y = c_meng_uv[idx[last]][1] - c_meng_uv[idx[0]][1];
            float vv = ex * e1y - e1x * ey;
            float area = e0x * e1y - e1x * e0y;
            float u = uu / area;
            float v = vv / area;
            float w = 1.0f - u - v;
            if (u < 0.0f || v < 0.0f || w < 0.0f) {
                uu = -vv;
                e0x = e1x;
                e0y = e1y;
                continue;
            }
            V4 pa = p[0], pb = p[0], pc = p[0];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                if (q == i + 1) pb = p[q];
                if (q == last) pc = p[q];
            }
            interpolated = pa * w + pb * v + pc * u;
            break;
        }
    }
    return interpolated * inv_norm;
}
PSY_DEV float plancks_law(float temperature, float wavelength) {
    if (temperature < 200.0f) return 0.0f;
    float wl = wavelength * 1.0e-9f;
    const float TMP1 = (float)(2.0 * (double)6.62607015e-34f * (double)299792458.0f * (double)299792458.0f);
    const float TMP2 = (float)((double)6.62607015e-34f * (double)299792458.0f / (double)1.38064852e-23f);
    float wl2 = wl * wl;
    float wl5 = wl2 * wl2 * wl;
    float tmp3 = wl5 * (fast_exp(TMP2 / (wl * temperature)) - 1.0f);
    float energy = TMP1 / tmp3;
    return rmax(energy * 1.0e-6f, 0.0f);
}
PSY_DEV float plancks_law_normalized(float temperature, float wavelength) {
    float t2 = temperature * temperature;
    return plancks_law(temperature, wavelength) * 4.0e7f / (t2 * t2);
}
PSY_NOINLINE V4 color_to_spectral(const DScene& S, const PsyColor& c, float hero) {
    float wls[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wls[i] = nth_wavelength(hero, i);
    V4 r;
    if (c.kind == PSY_COLOR_BAKED) {
        // the colour's own spectrum, interpolated over the chromaticity grid at upload (dscene.cuh)
        const float* __restrict__ sp = S.meng_baked + (size_t)__float_as_uint(c.a) * PSY_MENG_N;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float sb = (wls[k] - 360.0f) / (830.0f - 360.0f) * ((float)PSY_MENG_N - 1.0f);
            const int s0 = (int)sb;
            const int s1 = s0 + 1 < PSY_MENG_N - 1 ? s0 + 1 : PSY_MENG_N - 1;
            const float f = sb - (float)s0;
            r.v[k] = __ldg(sp + s0) * (1.0f - f) + __ldg(sp + s1) * f;
        }
        r = (r * c.b) * splat4(1.0f / PSY_MENG_EER);
    } else if (c.kind == PSY_COLOR_XYZ) {
        r = spectrum_xyz_to_p_4(S, wls, c.a, c.b, c.c) * splat4(1.0f / PSY_MENG_EER);
    } else if (c.kind == PSY_COLOR_BLACKBODY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r.v[i] = plancks_law(c.a, wls[i]) * c.b;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) r.v[i] = plancks_law_normalized(c.a, wls[i]) * c.b;
    }
    return r;
}
PSY_DEV PsyColor color_scale(const PsyColor& c, float s) {
    PsyColor r = c;
    if (c.kind == PSY_COLOR_BAKED) {
        r.b = c.b * s;
    } else if (c.kind == PSY_COLOR_XYZ) {
        r.a = c.a * s, r.b = c.b * s, r.c = c.c * s;
    } else {
        r.b = c.b * s;
    }
    return r;
}
PSY_DEV PsyColor color_lerp(const PsyColor& x, const PsyColor& y, float alpha) {
    float inv = 1.0f - alpha;
    PsyColor r;
    r.kind = x.kind;
    r.a = (x.a * inv) + (y.a * alpha);
    r.b = (x.b * inv) + (y.b * alpha);
    r.c = x.kind == PSY_COLOR_XYZ ? (x.c * inv) + (y.c * alpha) : 0.0f;
    return r;
}
PSY_DEV PsyColor lerp_slice_color(const PsyColor* s, uint32_t n, float alpha) {
    uint32_t i1;
    float a2;
    if (!lerp_slice_keys(n, alpha, i1, a2)) return s[i1];
    return color_lerp(s[i1], s[i1 + 1], a2);
}
PSY_DEV float x_1931(float wl) {
    float t1 = (wl - 442.0f) * ((wl < 442.0f) ? 0.0624f : 0.0374f);
    float t2 = (wl - 599.8f) * ((wl < 599.8f) ? 0.0264f : 0.0323f);
    float t3 = (wl - 501.1f) * ((wl < 501.1f) ? 0.0490f : 0.0382f);
    return (0.362f * fast_exp(-0.5f * t1 * t1)) + (1.056f * fast_exp(-0.5f * t2 * t2)) - (0.065f * fast_exp(-0.5f * t3 * t3));
}
PSY_DEV float y_1931(float wl) {
    float t1 = (wl - 568.8f) * ((wl < 568.8f) ? 0.0213f : 0.0247f);
    float t2 = (wl - 530.9f) * ((wl < 530.9f) ? 0.0613f : 0.0322f);
    return (0.821f * fast_exp(-0.5f * t1 * t1)) + (0.286f * fast_exp(-0.5f * t2 * t2));
}
PSY_DEV float z_1931(float wl) {
    float t1 = (wl - 437.0f) * ((wl < 437.0f) ? 0.0845f : 0.0278f);
    float t2 = (wl - 459.0f) * ((wl < 459.0f) ? 0.0385f : 0.0725f);
    return (1.217f * fast_exp(-0.5f * t1 * t1)) + (0.681f * fast_exp(-0.5f * t2 * t2));
}
PSY_DEV V3 xyz_from_spectral(V4 e, float hero) {
    float X[4], Y[4], Z[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float wl = nth_wavelength(hero, i);
        X[i] = x_1931(wl) * e.v[i];
        Y[i] = y_1931(wl) * e.v[i];
        Z[i] = z_1931(wl) * e.v[i];
    }
    return mk3((((X[0] + X[1]) + X[2]) + X[3]) * 0.75f, (((Y[0] + Y[1]) + Y[2]) + Y[3]) * 0.75f,
               (((Z[0] + Z[1]) + Z[2]) + Z[3]) * 0.75f);
}

//---- closures ------------------------------------------------------------------------------------
struct Closure {
    uint32_t kind;
    PsyColor color;
    float roughness, fresnel;
};
#define PSY_INV_PI (1.0f / PSY_PI_32)
#define PSY_H_PI (PSY_PI_32 / 2.0f)

PSY_DEV float ggx_d(float nh, float rough) {
    if (nh <= 0.0f) return 0.0f;
    float rough2 = rough * rough;
    float tmp = 1.0f + ((rough2 - 1.0f) * (nh * nh));
    return rough2 / (PSY_PI_32 * tmp * tmp);
}
PSY_DEV float ggx_g(float vh, float vn, float rough) {
    if ((vh * vn) <= 0.0f) return 0.0f;
    return 2.0f / (1.0f + sqrtf(1.0f + rough * rough * (1.0f - vn * vn) / (vn * vn)));
}
PSY_DEV float half_theta_sample(float u, float rough) {
    float rough2 = rough * rough;
    float top = 1.0f - u;
    float bottom = 1.0f + ((rough2 - 1.0f) * u);
    return sqrtf(top / bottom);
}
PSY_DEV float schlick_fresnel_from_fac(float fac, float c) {
    float c1 = 1.0f - c;
    float c2 = c1 * c1;
    return fac + ((1.0f - fac) * c1 * c2 * c2);
}
PSY_DEV void flip_normals(V3 inc, V3 nor, V3 nor_g, V3& nn, V3& fng) {
    if (dot(nor_g, inc) <= 0.0f) {
        nn = normalized(nor);
        fng = nor_g;
    } else {
        nn = -normalized(nor);
        fng = -nor_g;
    }
}
// KIND is the closure kind when the caller knows it at compile time (the shade kernels are
// specialised per closure class), or CLOSURE_ANY to branch on cl.kind.
constexpr uint32_t CLOSURE_ANY = 0xFFu;
#define PSY_KIND_IS(K) ((KIND == CLOSURE_ANY) ? (cl.kind == (K)) : (KIND == (K)))

// What sample() and evaluate() of one hit share: the closure colour at the path's hero wavelengths and
// the flipped normals.  The reference recomputes them in every call (surface_closure.rs:284-332,
// 420-533); they depend only on (closure, inc, nor, nor_g, wavelength), so one evaluation per hit gives
// the same bits.
struct ShadeCtx {
    V4 spec;     // color.to_spectral_sample(wavelength).e
    V3 nn, fng;  // flipped normalised shading normal / flipped geometric normal
    V3 aa;       // ggx: -inc.normalized()
};
template <uint32_t KIND>
PSY_DEV ShadeCtx closure_prepare(const DScene& S, const Closure& cl, V3 inc, V3 nor, V3 nor_g, float wl) {
    ShadeCtx c;
    c.spec = color_to_spectral(S, cl.color, wl);
    flip_normals(inc, nor, nor_g, c.nn, c.fng);
    c.aa = mk3(0, 0, 0);
    if (PSY_KIND_IS(PSY_SHADER_GGX)) c.aa = -normalized(inc);
    return c;
}
template <uint32_t KIND>
PSY_DEV void closure_evaluate(const Closure& cl, const ShadeCtx& sc, V3 out, V4& filter, float& pdf) {
    if (PSY_KIND_IS(PSY_SHADER_LAMBERT)) {
        if (dot(sc.fng, out) >= 0.0f) {
            float fac = rmax(dot(sc.nn, normalized(out)), 0.0f) * PSY_INV_PI;
            filter = sc.spec * fac;
            pdf = fac;
        } else {
            filter = splat4(0.0f);
            pdf = 0.0f;
        }
    } else if (PSY_KIND_IS(PSY_SHADER_GGX)) {
        V3 aa = sc.aa;
        V3 bb = normalized(out);
        V3 hh = normalized(aa + bb);
        V3 nn = sc.nn, fng = sc.fng;
        if (dot(nn, aa) < 0.0f || dot(nn, bb) < 0.0f || dot(fng, bb) < 0.0f) {
            filter = splat4(0.0f);
            pdf = 0.0f;
            return;
        }
        float na = rclamp(dot(nn, aa), -1.0f, 1.0f);
        float nb = rclamp(dot(nn, bb), -1.0f, 1.0f);
        float ha = rclamp(dot(hh, aa), -1.0f, 1.0f);
        float hb = rclamp(dot(hh, bb), -1.0f, 1.0f);
        float nh = rclamp(dot(nn, hh), -1.0f, 1.0f);
        V4 ss = sc.spec;
        float rev_fresnel = 1.0f - cl.fresnel;
        V4 col_f;
#pragma unroll
        for (int i = 0; i < 4; ++i) col_f.v[i] = lerpf(schlick_fresnel_from_fac(ss.v[i], hb), ss.v[i], rev_fresnel);
        if (cl.roughness == 0.0f) {
            filter = col_f;
            pdf = 0.0f;
            return;
        }
        float dist = ggx_d(nh, cl.roughness) / na;
        float g1 = ggx_g(ha, na, cl.roughness);
        float g2 = ggx_g(hb, nb, cl.roughness);
        filter = col_f * (dist * g1 * g2) * PSY_INV_PI;
        pdf = dist * PSY_INV_PI;
    } else {
        filter = sc.spec;
        pdf = 1.0f;
    }
}
template <uint32_t KIND>
PSY_DEV void closure_sample(const Closure& cl, const ShadeCtx& sc, V3 inc, float u, float v, V3& dir, V4& filter,
                            float& pdf) {
    if (PSY_KIND_IS(PSY_SHADER_LAMBERT)) {
        V3 d = cosine_sample_hemisphere(u, v);
        float p = d.z * PSY_INV_PI;
        V3 out = zup_to_vec(d, sc.nn);
        dir = out;
        if (dot(sc.fng, out) >= 0.0f) {
            filter = sc.spec * p;
            pdf = p;
        } else {
            filter = splat4(0.0f);
            pdf = 0.0f;
        }
    } else if (PSY_KIND_IS(PSY_SHADER_GGX)) {
        float theta_cos = half_theta_sample(u, cl.roughness);
        float theta_sin = sqrtf(1.0f - (theta_cos * theta_cos));
        float angle = v * PSY_PI_32 * 2.0f;
        V3 half_dir = mk3(PSY_COSF(angle) * theta_sin, PSY_SINF(angle) * theta_sin, theta_cos);
        half_dir = normalized(zup_to_vec(half_dir, sc.nn));
        V3 out = inc - (half_dir * 2.0f * dot(inc, half_dir));
        dir = out;
        if (dot(sc.fng, out) >= 0.0f) {
            closure_evaluate<KIND>(cl, sc, out, filter, pdf);
        } else {
            filter = splat4(0.0f);
            pdf = 0.0f;
        }
    } else {
        dir = mk3(0, 0, 0);
        filter = sc.spec;
        pdf = 1.0f;
    }
}
// sphere_lambert (surface_closure.rs:346-377).  The reference evaluates g and h before branching; they
// only feed the two partially-occluded cases, so they are evaluated there (pure functions: same bits).
PSY_NOINLINE float sphere_lambert(float nlcos, float rcos) {
    float rsin2 = 1.0f - (rcos * rcos);
    // The two acos calls below only pick the case.  acos is monotone, so `nl < pi/2 - r` is the same as
    // `nlcos > sin(r)`, and sin(r) = sqrt(1 - rcos^2) exactly.  With an upper bound of sin(r) that covers the
    // rounding of rsin2 (<= 6e-8 absolute: one rounded product, the subtraction is exact) and a margin far above
    // what rounding in acosf (<= 2 ulp) and in `H_PI - r` can move the comparison (2e-6 against < 5e-7), the outer
    // two cases are decided without the calls — same branch, same arithmetic in it, same bits.  The light-tree walk
    // calls this 8 times per level and most children are small and far away: the first case.
    {
        const float rsin_hi = sqrtf(rsin2 + 2.0e-7f) + 2.0e-6f;
        if (nlcos > rsin_hi) return nlcos * rsin2;
        if (nlcos < -rsin_hi) return 0.0f;
    }
    float nl = PSY_ACOSF(nlcos);
    float r = PSY_ACOSF(rcos);
    if (nl < (PSY_H_PI - r)) return nlcos * rsin2;
    if (!(nl < PSY_H_PI) && !(nl < (PSY_H_PI + r))) return 0.0f;
    float nlsin = sqrtf(1.0f - (nlcos * nlcos));
    float rsin = sqrtf(rsin2);
    float ysin = rcos / nlsin;
    float ycos2 = 1.0f - (ysin * ysin);
    float ycos = sqrtf(ycos2);
    float g = (-2.0f * nlsin * rcos * ycos) + PSY_H_PI - PSY_ASINF(ysin) + (ysin * ycos);
    float h = nlcos * ((ycos * sqrtf(rsin2 - ycos2)) + (rsin2 * PSY_ASINF(ycos / rsin)));
    if (nl < PSY_H_PI) return (nlcos * rsin2) + g - h;
    return (g + h) * PSY_INV_PI;
}
// The part of estimate_eval_over_sphere_light that does not depend on the light: evaluated once per
// light-tree walk instead of once per child (same inputs, same bits).
struct EstimateCtx {
    V3 nn;  // flipped, normalised shading normal (lambert: surface_closure.rs:382-389, ggx: :561-567)
    V3 aa;  // ggx: -inc.normalized()
};
template <uint32_t KIND>
PSY_DEV EstimateCtx closure_estimate_ctx(const Closure& cl, V3 inc, V3 nor, V3 nor_g) {
    EstimateCtx e;
    e.nn = mk3(0, 0, 0), e.aa = mk3(0, 0, 0);
    if (PSY_KIND_IS(PSY_SHADER_LAMBERT)) {
        e.nn = dot(nor_g, inc) <= 0.0f ? normalized(nor) : -normalized(nor);
    } else if (PSY_KIND_IS(PSY_SHADER_GGX)) {
        e.nn = dot(nor, inc) < 0.0f ? normalized(nor) : -normalized(nor);
        e.aa = -normalized(inc);
    }
    return e;
}
template <uint32_t KIND>
PSY_DEV float closure_estimate_sphere(const Closure& cl, const EstimateCtx& ec, V3 to_light_center,
                                      float light_radius_squared) {
    if (PSY_KIND_IS(PSY_SHADER_LAMBERT)) {
        float dist2 = length2(to_light_center);
        if (dist2 <= light_radius_squared) return rmin(light_radius_squared / dist2, 4.0f);
        float sin_theta_max2 = rmin(light_radius_squared / dist2, 1.0f);
        float cos_theta_max = sqrtf(1.0f - sin_theta_max2);
        V3 v = normalized(to_light_center);
        V3 nn = ec.nn;
        float cos_nv = rmin(rmax(dot(nn, v), -1.0f), 1.0f);
        return sphere_lambert(cos_nv, cos_theta_max);
    } else if (PSY_KIND_IS(PSY_SHADER_GGX)) {
        float dist2 = length2(to_light_center);
        float sin_theta_max2 = rmin(light_radius_squared / dist2, 1.0f);
        float cos_theta_max = sqrtf(1.0f - sin_theta_max2);
        V3 nn = ec.nn;
        V3 aa = ec.aa;
        V3 bb = normalized(to_light_center);
        float theta = PSY_ACOSF(cos_theta_max);
        V3 hh = normalized(aa + bb);
        float nh = rclamp(dot(nn, hh), -1.0f, 1.0f);
        float fac = ggx_d(nh, rmin(1.0f, sqrtf(cl.roughness) + (2.0f * theta / PSY_PI_32)));
        return fac * rmin(1.0f, 1.0f - cos_theta_max) * PSY_INV_PI;
    }
    return 0.0f;
}

//---- intersection data (finalisation of a HitRec) -------------------------------------------------
struct Isect {
    uint32_t kind;  // PSY_HIT_*
    V3 incoming, pos, nor, nor_g;
    float pos_err, t, sample_pdf;
    Closure closure;
};

PSY_DEV Closure shader_closure(const DScene& S, uint32_t shader_index) {
    Closure c;
    if (shader_index == PSY_NONE) {
        // unassigned_shader: Emit(rec709_to_xyz(1,0,1)) (tracer.rs:161-164)
        c.kind = PSY_SHADER_EMIT;
        c.color.kind = PSY_COLOR_XYZ;
        c.color.a = (1.0f * 0.4123907993f) + (0.0f * 0.3575843394f) + (1.0f * 0.1804807884f);
        c.color.b = (1.0f * 0.2126390059f) + (0.0f * 0.7151686788f) + (1.0f * 0.0721923154f);
        c.color.c = (1.0f * 0.0193308187f) + (0.0f * 0.1191947798f) + (1.0f * 0.9505321522f);
        c.roughness = 0.0f, c.fresnel = 0.0f;
        return c;
    }
    const PsyShader s = S.shaders[shader_index];
    c.kind = s.kind;
    c.color = s.color;
    c.roughness = s.roughness;
    c.fresnel = s.fresnel;
    return c;
}
// surface_point (triangle.rs:155-169)
PSY_DEV void surface_point(V3 t0, V3 t1, V3 t2, float b0, float b1, float b2, V3& pos, float& pos_err) {
    pos = ((t0 * b0) + (t1 * b1)) + (t2 * b2);
    V3 e = (((vabs(t0) * b0) + (vabs(t1) * b1)) + (vabs(t2) * b2)) * fp_gamma(7);
    pos_err = max_element(e);
}
PSY_NOINLINE float sphere_sample_pdf(const DScene& S, const PsyLight& l, const Xf& space, V3 arr, float time) {
    V3 a = xf_point(space, arr);
    double radius = (double)lerp_slice_f(S.light_shape + l.shape_base, l.n_shape, time);
    double d2 = (double)length2(mk3(0, 0, 0) - a);
    double dd = sqrt(d2);
    if (dd > radius) {
        double sin_theta_max2 = rmin((radius * radius) / d2, 1.0);
        double cos_theta_max = sqrt(1.0 - sin_theta_max2);
        return (float)uniform_sample_cone_pdf(cos_theta_max);
    }
    return (float)(1.0 / (4.0 * PSY_PI_64));
}
PSY_NOINLINE float rect_sample_pdf(const Xf& space_inv, float w, float h, V3 arr, V3 sample_dir, V3 hit_point) {
    V3 p1, p2, p3, p4;
    rect_corners(space_inv, w, h, p1, p2, p3, p4);
    V3 sp1 = normalized(p1 - arr), sp2 = normalized(p2 - arr), sp3 = normalized(p3 - arr), sp4 = normalized(p4 - arr);
    float area_1 = spherical_triangle_solid_angle(sp2, sp1, sp3);
    float area_2 = spherical_triangle_solid_angle(sp4, sp1, sp3);
    V3 normal = xf_normal(space_inv, mk3(0, 0, 1));
    if ((area_1 + area_2) < 0.01f) {
        float area = triangle_surface_area(p2, p1, p3) + triangle_surface_area(p4, p1, p3);
        return length2(hit_point - arr) / fabsf(dot(normalized(sample_dir), normalized(normal))) / area;
    }
    return 1.0f / (area_1 + area_2);
}

// Rebuilds SurfaceIntersectionData + closure for the hit the trace kernel recorded
// (triangle_mesh.rs:264-318, sphere_light.rs:281-334, rectangle_light.rs:296-330).
__device__ __noinline__ void finalize_hit(const DScene& S, const HitRec& h, V3 orig, V3 dir, float time, Isect& is) {
    is.kind = h.kind;
    is.incoming = dir;
    is.t = h.t;
    is.sample_pdf = 0.0f;
    if (h.kind != PSY_HIT_TRIANGLE && h.kind != PSY_HIT_SPHERE_LIGHT && h.kind != PSY_HIT_RECT_LIGHT) return;
    const DPath P = S.paths[h.path];
    const PsyInstance inst = S.instances[P.inst];
    const PsyObject obj = S.objects[inst.data_index];
    if (h.kind == PSY_HIT_TRIANGLE) {
        const DMesh m = S.meshes[obj.index];
        V3 p0, p1, p2;
        load_tri(S, m, h.prim, time, p0, p1, p2);
        Xf mat_space = xf_identity();
        if (P.xf_count > 0) {
            mat_space = path_mat_space(S, P, h.path, time);
            p0 = xf_point(mat_space, p0);
            p1 = xf_point(mat_space, p1);
            p2 = xf_point(mat_space, p2);
        }
        surface_point(p0, p1, p2, h.b0, h.b1, h.b2, is.pos, is.pos_err);
        V3 geo_normal = cross(p0 - p1, p0 - p2);
        V3 shading_normal = geo_normal;
        if (m.has_normals) {
            const uint4 idx = __ldg(S.tris + h.prim);
            uint32_t k = m.n_keys;
            uint32_t i1;
            float a2;
            bool blend = lerp_slice_keys(k, time, i1, a2);
            const float4* nb = S.normals + m.vert_base;
            V3 n0 = f4_xyz(__ldg(nb + idx.x * k + i1));
            V3 n1 = f4_xyz(__ldg(nb + idx.y * k + i1));
            V3 n2 = f4_xyz(__ldg(nb + idx.z * k + i1));
            if (blend) {
                n0 = lerp3(n0, f4_xyz(__ldg(nb + idx.x * k + i1 + 1)), a2);
                n1 = lerp3(n1, f4_xyz(__ldg(nb + idx.y * k + i1 + 1)), a2);
                n2 = lerp3(n2, f4_xyz(__ldg(nb + idx.z * k + i1 + 1)), a2);
            }
            n0 = normalized(n0), n1 = normalized(n1), n2 = normalized(n2);
            V3 s_nor = xf_normal(mat_space, ((n0 * h.b0) + (n1 * h.b1)) + (n2 * h.b2));
            shading_normal = dot(s_nor, geo_normal) >= 0.0f ? s_nor : -s_nor;
        }
        is.nor = shading_normal;
        is.nor_g = geo_normal;
        is.closure = shader_closure(S, inst.shader_index);
    } else if (h.kind == PSY_HIT_SPHERE_LIGHT) {
        const PsyLight l = S.lights[obj.index];
        Xf xform = path_xform(S, P, time);
        float radius = lerp_slice_f(S.light_shape + l.shape_base, l.n_shape, time);
        V3 orig_local = xf_point(xform, orig);
        V3 ldir = xf_vector(xform, dir);
        Xf inv_xform = xf_inverse(xform);
        V3 t_pos = orig_local + (ldir * h.t);
        V3 unit_pos = normalized(t_pos);
        is.pos = xf_vector(inv_xform, unit_pos * radius);
        is.pos_err = 0.001f;
        V3 normal = xf_normal(inv_xform, unit_pos);
        is.nor = normal;
        is.nor_g = normal;
        is.sample_pdf = sphere_sample_pdf(S, l, xform, orig, time);
        float inv_surface_area = (float)(1.0 / (4.0 * PSY_PI_64 * (double)radius * (double)radius));
        is.closure.kind = PSY_SHADER_EMIT;
        is.closure.color = color_scale(lerp_slice_color(S.light_colors + l.color_base, l.n_colors, time), inv_surface_area);
        is.closure.roughness = 0.0f, is.closure.fresnel = 0.0f;
    } else {
        const PsyLight l = S.lights[obj.index];
        Xf xform = path_xform(S, P, time);
        float w, hh;
        rect_dims(S.light_shape, l, time, w, hh);
        Xf space_inv = xf_inverse(xform);
        V3 p1, p2, p3, p4;
        rect_corners(space_inv, w, hh, p1, p2, p3, p4);
        V3 a = h.prim == 0 ? p1 : p3, b = h.prim == 0 ? p2 : p4, c = h.prim == 0 ? p3 : p1;
        surface_point(a, b, c, h.b0, h.b1, h.b2, is.pos, is.pos_err);
        V3 normal = cross(a - b, a - c);
        is.nor = normal;
        is.nor_g = normal;
        is.sample_pdf = rect_sample_pdf(space_inv, w, hh, orig, dir, is.pos);
        float inv_surface_area = (float)(1.0 / ((double)w * (double)hh));
        is.closure.kind = PSY_SHADER_EMIT;
        is.closure.color = color_scale(lerp_slice_color(S.light_colors + l.color_base, l.n_colors, time), inv_surface_area);
        is.closure.roughness = 0.0f, is.closure.fresnel = 0.0f;
    }
}

//---- light sampling ------------------------------------------------------------------------------
struct LightSample {
    int kind;  // 0 none, 1 distant, 2 surface
    V4 color;
    V3 direction;
    V3 pos, nor;
    float err, pdf, selection_pdf;
};

PSY_DEV V3 bbv(const float* p) { return mk3(p[0], p[1], p[2]); }

// LightTree::select (light_tree.rs:135-201)
template <uint32_t KIND>
PSY_DEV bool light_tree_select(const DScene& S, const PsyAssembly& a, V3 inc, V3 pos, V3 nor, V3 nor_g,
                               const Closure& sc, float time, float n, uint32_t& light_i, float& sel_pdf,
                               float& whittled_n) {
    if (a.light_root == PSY_NONE) return false;
    PsyLightNode node = S.light_nodes[a.light_root];
    float tot_prob = 1.0f;
    const EstimateCtx ec = closure_estimate_ctx<KIND>(sc, inc, nor, nor_g);
    while (node.child_count > 0) {
        float ps[8];
        float total = 0.0f;
        uint32_t nc = node.child_count;
#pragma unroll 1
        for (uint32_t i = 0; i < 8; ++i) {
            float p = 0.0f;
            if (i < nc) {
                const PsyLightNode nd = S.light_nodes[node.child_base + i];
                const PsyBBox* bb = S.lt_bounds + nd.bounds_base;
                uint32_t i1;
                float a2;
                V3 bmin, bmax;
                if (!lerp_slice_keys(nd.bounds_count, time, i1, a2)) {
                    bmin = bbv(bb[i1].min), bmax = bbv(bb[i1].max);
                } else {
                    bmin = lerp3(bbv(bb[i1].min), bbv(bb[i1 + 1].min), a2);
                    bmax = lerp3(bbv(bb[i1].max), bbv(bb[i1 + 1].max), a2);
                }
                V3 center = lerp3(bmin, bmax, 0.5f);
                V3 dd = center - pos;
                float r2 = length2(bmax - bmin) * 0.25f;
                float inv_surface_area = 1.0f / r2;
                float approx_contrib = closure_estimate_sphere<KIND>(sc, ec, dd, r2);
                p = nd.energy * inv_surface_area * approx_contrib;
                total += p;
            }
            ps[i] = p;
        }
        if (total <= 0.0f) {
            float p = 1.0f / (float)nc;
#pragma unroll
            for (int i = 0; i < 8; ++i) ps[i] = p;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ps[i] /= total;
        }
        float base = 0.0f;
        uint32_t pick = nc - 1;
        float pick_p = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            float p = ps[i];
            if ((n <= base + p) || (i == nc - 1)) {
                pick = i;
                pick_p = p;
                break;
            } else {
                base += p;
            }
        }
        tot_prob *= pick_p;
        n = (n - base) / pick_p;
        node = S.light_nodes[node.child_base + pick];
    }
    light_i = node.light_index;
    sel_pdf = tot_prob;
    whittled_n = n;
    return true;
}

// SphereLight::sample_from_point (sphere_light.rs:85-192)
__device__ __noinline__ void sample_sphere_light(const DScene& S, const PsyLight& l, const Xf& space, V3 arr_w, float u,
                                                 float v, float wl, float time, LightSample& out) {
    V3 arr = xf_point(space, arr_w);
    V3 pos = mk3(0, 0, 0);
    Xf inv_space = xf_inverse(space);
    double radius = (double)lerp_slice_f(S.light_shape + l.shape_base, l.n_shape, time);
    PsyColor col = lerp_slice_color(S.light_colors + l.color_base, l.n_colors, time);
    double surface_area_inv = 1.0 / (4.0 * PSY_PI_64 * radius * radius);
    V3 z = pos - arr;
    double d2 = (double)length2(z);
    double dd = sqrt(d2);
    V3 x, y;
    coordinate_system_from_vector(z, x, y);
    x = normalized(x), y = normalized(y), z = normalized(z);
    {
        V3 vr = mk3((float)radius, (float)radius, (float)radius);
        V3 v2 = xf_vector(inv_space, vr);
        out.err = length(v2) * 0.001f;
    }
    V3 normal;
    if (dd > radius) {
        double sin_theta_max2 = rmin((radius * radius) / d2, 1.0);
        double cos_theta_max2 = 1.0 - sin_theta_max2;
        double sin_theta_max = sqrt(sin_theta_max2);
        double cos_theta_max = sqrt(cos_theta_max2);
        V3 sample = normalized(uniform_sample_cone(u, v, cos_theta_max));
        double cos_theta = (double)sample.z;
        double cos_theta2 = cos_theta * cos_theta;
        double sin_theta2 = rmax(1.0 - cos_theta2, 0.0);
        double sin_theta = sqrt(sin_theta2);
        double ddd = 1.0 - (d2 * sin_theta * sin_theta / (radius * radius));
        double cos_a = ddd <= 0.0 ? sin_theta_max : ((dd / radius) * sin_theta2) + (cos_theta * sqrt(ddd));
        double sin_a = sqrt(rmax(1.0 - (cos_a * cos_a), 0.0));
        double phi = (double)v * 2.0 * PSY_PI_64;
        V3 s2 = mk3((float)(PSY_COS(phi) * sin_a * radius), (float)(PSY_SIN(phi) * sin_a * radius), (float)(dd - (cos_a * radius)));
        V3 sample_vec = ((x * s2.x) + (y * s2.y)) + (z * s2.z);
        normal = normalized(arr + sample_vec);
        out.pdf = (float)uniform_sample_cone_pdf(cos_theta_max);
    } else {
        V3 sample_vec = uniform_sample_sphere(u, v);
        normal = normalized(arr + sample_vec);
        out.pdf = (float)(1.0 / (4.0 * PSY_PI_64));
    }
    V3 point = normal * (float)radius;
    out.pos = xf_point(inv_space, point);
    out.nor = xf_normal(inv_space, normal);
    out.color = color_to_spectral(S, col, wl) * (float)surface_area_inv;
}

// RectangleLight::sample_from_point (rectangle_light.rs:121-243)
__device__ __noinline__ void sample_rect_light(const DScene& S, const PsyLight& l, const Xf& space, V3 arr, float u,
                                               float v, float wl, float time, LightSample& out) {
    float w, h;
    rect_dims(S.light_shape, l, time, w, h);
    PsyColor col = lerp_slice_color(S.light_colors + l.color_base, l.n_colors, time);
    double surface_area = (double)w * (double)h;
    double surface_area_inv = 1.0 / surface_area;
    Xf space_inv = xf_inverse(space);
    V3 p1, p2, p3, p4;
    rect_corners(space_inv, w, h, p1, p2, p3, p4);
    V3 sp1 = normalized(p1 - arr), sp2 = normalized(p2 - arr), sp3 = normalized(p3 - arr), sp4 = normalized(p4 - arr);
    float area_1 = spherical_triangle_solid_angle(sp2, sp1, sp3);
    float area_2 = spherical_triangle_solid_angle(sp4, sp1, sp3);
    V3 normal = xf_normal(space_inv, mk3(0, 0, 1));
    out.nor = normal;
    out.err = 0.0001f;
    if ((area_1 + area_2) < 0.01f) {
        float sa1 = triangle_surface_area(p2, p1, p3);
        float sa2 = triangle_surface_area(p4, p1, p3);
        float threshhold = sa1 / (sa1 + sa2);
        V3 sample_point;
        if (u < threshhold)
            sample_point = uniform_sample_triangle(p2, p1, p3, v, u / threshhold);
        else
            sample_point = uniform_sample_triangle(p4, p1, p3, v, (u - threshhold) / (1.0f - threshhold));
        V3 shadow_vec = sample_point - arr;
        out.color = color_to_spectral(S, col, wl) * (float)surface_area_inv * 0.5f;
        out.pdf = length2(sample_point - arr) / fabsf(dot(normalized(shadow_vec), normalized(normal))) / (sa1 + sa2);
        out.pos = sample_point;
    } else {
        float prob_1 = isinf(area_1) ? 1.0f : (isinf(area_2) ? 0.0f : area_1 / (area_1 + area_2));
        float prob_2 = 1.0f - prob_1;
        V3 shadow_vec = u < prob_1 ? uniform_sample_spherical_triangle(sp2, sp1, sp3, v, u / prob_1)
                                   : uniform_sample_spherical_triangle(sp4, sp1, sp3, v, 1.0f - ((u - prob_1) / prob_2));
        V3 arr_local = xf_point(space, arr);
        V3 svl = xf_vector(space, shadow_vec);
        svl = svl * (-arr_local.z / svl.z);
        V3 spl = arr_local + svl;
        spl.x = rmin(rmax(spl.x, w * -0.5f), w * 0.5f);
        spl.y = rmin(rmax(spl.y, h * -0.5f), h * 0.5f);
        spl.z = 0.0f;
        out.pos = xf_point(space_inv, spl);
        out.pdf = 1.0f / (area_1 + area_2);
        out.color = color_to_spectral(S, col, wl) * (float)surface_area_inv * 0.5f;
    }
}

PSY_DEV float distant_energy(const DScene& S, const PsyDistantLight& l) {
    float a = 0.0f;
    for (uint32_t i = 0; i < l.n_colors; ++i) {
        const PsyColor c = S.light_colors[l.color_base + i];
        float e = c.kind == PSY_COLOR_BLACKBODY ? (c.a * c.a) * (c.a * c.a) * c.b : c.b;
        a = a + e;
    }
    return a / (float)l.n_colors;
}

// Scene::sample_lights + Assembly::sample_lights (scene/mod.rs:28-99, scene/assembly.rs:46-140) in two
// halves so that the light-tree walk and the (f64-heavy) light sampling can live in different kernels:
//   select_light        world/local choice, weighted_choice, LightTree::select through nested assemblies
//   sample_picked_light DistantDiskLight / SphereLight / RectangleLight::sample_from_point
struct LightPick {
    uint32_t kind;         // 0 none, 1 distant light, 2 surface light
    uint32_t index;        // distant: index into distant_lights; surface: global light_instances index
    uint32_t parent_path;  // path of the assembly instance that contains the light (PSY_NONE = root)
    float selection_pdf;
};
template <uint32_t KIND>
__device__ __noinline__ void select_light(const DScene& S, float n, float time, const Isect& intr, LightPick& out) {
    out.kind = 0;
    out.index = 0;
    out.parent_path = PSY_NONE;
    out.selection_pdf = 0.0f;
    float wl_energy = S.world_light_energy_sum <= 0.0f ? 0.0f : 1.0f;
    float ll_energy = S.assemblies[S.root_assembly].light_energy <= 0.0f ? 0.0f : 1.0f;
    float tot_energy = wl_energy + ll_energy;
    if (tot_energy <= 0.0f) return;
    float wl_prob = wl_energy / tot_energy;
    if (n < wl_prob) {
        float nn = n / wl_prob;
        // weighted_choice (algorithm.rs:17-36)
        float total_weight = 0.0f;
        for (uint32_t i = 0; i < S.n_distant_lights; ++i) total_weight = total_weight + distant_energy(S, S.distant_lights[i]);
        float target = nn * total_weight;
        float x = 0.0f, p = 0.0f;
        uint32_t pick = PSY_NONE;
        for (uint32_t i = 0; i < S.n_distant_lights; ++i) {
            float wgt = distant_energy(S, S.distant_lights[i]);
            x += wgt;
            if (x > target) {
                pick = i;
                p = wgt / total_weight;
                break;
            }
        }
        if (pick == PSY_NONE) return;
        out.kind = 1;
        out.index = pick;
        out.selection_pdf = p * wl_prob;
        return;
    }
    float nn = (n - wl_prob) / (1.0f - wl_prob);
    if (intr.kind == PSY_HIT_MISS || intr.kind == PSY_HIT_OCCLUDED) return;

    // descend through nested assemblies
    uint32_t asm_index = S.root_assembly;
    uint32_t path_base = 0;
    uint32_t cur_path = PSY_NONE;  // path of the assembly instance we are inside (PSY_NONE = root)
    float sel_chain[TRACE_MAX_NEST];
    int depth = 0;
    for (;;) {
        const PsyAssembly a = S.assemblies[asm_index];
        // xform_stack.top() is the composed transform of the enclosing assembly instance
        bool have_parent = cur_path != PSY_NONE && S.paths[cur_path].xf_count > 0;
        Xf sel_xform = have_parent ? path_xform(S, S.paths[cur_path], time) : xf_identity();
        uint32_t light_i;
        float sel_pdf, whittled_n;
        if (!light_tree_select<KIND>(S, a, xf_vector(sel_xform, intr.incoming), xf_point(sel_xform, intr.pos),
                               xf_normal(sel_xform, intr.nor), xf_normal(sel_xform, intr.nor_g), intr.closure, time, nn,
                               light_i, sel_pdf, whittled_n))
            return;
        const PsyInstance inst = S.light_instances[a.light_instance_base + light_i];
        if (inst.kind == PSY_INSTANCE_OBJECT) {
            out.kind = 2;
            out.index = a.light_instance_base + light_i;
            out.parent_path = cur_path;
            float sp = sel_pdf;
            // sample.map(|..spdf| spdf * sel_pdf) applied from the innermost level outwards
            for (int k = depth - 1; k >= 0; --k) sp = sp * sel_chain[k];
            out.selection_pdf = sp * (1.0f - wl_prob);
            return;
        }
        if (depth >= TRACE_MAX_NEST) return;
        sel_chain[depth++] = sel_pdf;
        uint32_t local = S.light_inst_local[a.light_instance_base + light_i];
        if (local == PSY_NONE) return;
        cur_path = path_base + local;
        path_base = S.paths[cur_path].child_base;
        asm_index = inst.data_index;
        nn = whittled_n;
    }
}

__device__ __noinline__ void sample_picked_light(const DScene& S, const LightPick& pick, float u, float v, float wl,
                                                 float time, V3 pos, LightSample& out) {
    out.kind = 0;
    out.pdf = 0.0f;
    out.selection_pdf = pick.selection_pdf;
    if (pick.kind == 1u) {
        // DistantDiskLight::sample_from_point (distant_disk_light.rs:59-85)
        const PsyDistantLight l = S.distant_lights[pick.index];
        double radius = (double)lerp_slice_f(S.world_floats + l.radii_base, l.n_radii, time);
        V3 direction;
        {
            uint32_t i1;
            float a2;
            const float* dp = S.world_floats + l.dir_base;
            if (!lerp_slice_keys(l.n_dirs, time, i1, a2))
                direction = bbv(dp + 3 * i1);
            else
                direction = lerp3(bbv(dp + 3 * i1), bbv(dp + 3 * i1 + 3), a2);
        }
        PsyColor col = lerp_slice_color(S.light_colors + l.color_base, l.n_colors, time);
        double solid_angle_inv = 1.0 / (2.0 * PSY_PI_64 * (1.0 - PSY_COS(radius)));
        V3 z = -normalized(direction), x2, y2;
        coordinate_system_from_vector(z, x2, y2);
        double cos_theta_max = PSY_COS(radius);
        V3 sample = normalized(uniform_sample_cone(u, v, cos_theta_max));
        out.kind = 1;
        out.color = color_to_spectral(S, col, wl) * (float)solid_angle_inv;
        out.direction = ((x2 * sample.x) + (y2 * sample.y)) + (z * sample.z);
        out.pdf = (float)uniform_sample_cone_pdf(cos_theta_max);
    } else if (pick.kind == 2u) {
        // the light instance's world->object transform (assembly.rs:77-97)
        const PsyInstance inst = S.light_instances[pick.index];
        const PsyObject o = S.objects[inst.data_index];
        const PsyLight l = S.lights[o.index];
        bool have_parent = pick.parent_path != PSY_NONE && S.paths[pick.parent_path].xf_count > 0;
        Xf sel_xform = have_parent ? path_xform(S, S.paths[pick.parent_path], time) : xf_identity();
        Xf xform;
        if (inst.xform_count > 0) {
            Xf own = lerp_slice_xf(S.xforms + 3 * (size_t)inst.xform_base, inst.xform_count, time);
            xform = have_parent ? xf_mul(sel_xform, own) : own;
        } else {
            xform = sel_xform;
        }
        out.kind = 2;
        if (l.kind == PSY_LIGHT_SPHERE)
            sample_sphere_light(S, l, xform, pos, u, v, wl, time, out);
        else
            sample_rect_light(S, l, xform, pos, u, v, wl, time, out);
    }
}

}  // namespace psyd
