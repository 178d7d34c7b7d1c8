// oracle/osample.h — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Sampler, hashing, Monte-Carlo warps and fastapprox, restated from:
//   src/hash.rs:1-29, src/renderer.rs:698-737 (get_sample_4d, golden_ratio_sample),
//   src/math.rs:67-132 (probit), src/sampling/monte_carlo.rs:10-199, src/mis.rs:12-21.
//
// THIRD-PARTY, NOT UNDER /root/reference (parity with upstream binaries is UNPINNED for these):
//   * sobol_burley 0.3.0 (Cargo.lock) — `sample_4d(i, dimension_set, seed)`: Owen-scrambled,
//     index-shuffled Sobol.  Restated from the crate's published design: shuffle the bit-reversed
//     index with a Laine-Karras style hash seeded by hash(seed); XOR the bit-reversed Joe-Kuo
//     direction vectors selected by the top 16 bits; Owen-scramble each lane with a per-lane
//     hashed seed; reverse the bits back; put the top 23 bits into the mantissa of a float in [1,2) and subtract 1.  The CUDA path
//     uses exactly the same integer recipe, so "identical sample sequences" holds between oracle
//     and GPU; whether every constant equals upstream's cannot be checked offline.
//   * fastapprox 0.3.0 (Cargo.lock) — fast::pow2 / fast::exp, restated from the published
//     fastapprox formulas (P. Mineiro).  Unpinned likewise.
#pragma once
#include "omath.h"

#define SOBOL_TABLE_QUAL static const
#include "../tables/sobol_tables.inc"

namespace orc {

inline uint32_t hash_u32(uint32_t n, uint32_t seed) {
    uint32_t h = n;
    for (int i = 0; i < 3; ++i) {
        h *= 0x736caf6fu;
        h ^= h >> 16;
        h ^= seed;
    }
    return h;
}
inline float hash_u32_to_f32(uint32_t n, uint32_t seed) {
    const float INV_MAX = 1.0f / 4294967295.0f;
    return (float)hash_u32(n, seed) * INV_MAX;
}

inline uint32_t reverse_bits(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}

// sobol_burley: seed hash
inline uint32_t sb_hash(uint32_t n) {
    uint32_t h = n ^ 0x79c68e4au;
    h ^= h >> 16;
    h *= 0x7feb352du;
    h ^= h >> 15;
    h *= 0x846ca68bu;
    h ^= h >> 16;
    return h;
}
// sobol_burley: Laine-Karras style scramble of a bit-reversed integer
inline uint32_t sb_lk_scramble(uint32_t n, uint32_t hashed_seed) {
    n ^= n * 0x3d20adeau;
    n += hashed_seed;
    n *= (hashed_seed >> 16) | 1u;
    n ^= n * 0x05526c56u;
    n ^= n * 0x53a22864u;
    return n;
}
inline void sobol_sample_4d(uint32_t index, uint32_t dimension_set, uint32_t seed, float out[4]) {
    // index shuffle
    uint32_t shuffled_rev = sb_lk_scramble(reverse_bits(index), sb_hash(seed));
    uint32_t sob[4] = {0, 0, 0, 0};
    uint32_t top = shuffled_rev >> 16;  // only the top 16 bits are used
    for (int bit = 0; bit < 16; ++bit) {
        if (top & (0x8000u >> bit)) {
            for (int l = 0; l < 4; ++l) sob[l] ^= SOBOL_REV_VECTORS[dimension_set][bit][l];
        }
    }
    static const uint32_t LANE_SALT[4] = {0x912f69bau, 0x174f18abu, 0x691e72cau, 0xb40cc1b8u};
    for (int l = 0; l < 4; ++l) {
        uint32_t lane_seed = (dimension_set ^ seed) ^ LANE_SALT[l];
        for (int r = 0; r < 2; ++r) {
            lane_seed *= 0x736caf6fu;
            lane_seed ^= lane_seed >> 16;
        }
        uint32_t v = reverse_bits(sb_lk_scramble(sob[l], lane_seed));
        // sobol_burley's to_norm_floats: the top 23 bits become the mantissa of a float in [1, 2), minus 1
        const uint32_t fb = (v >> 9) | 0x3f800000u;
        float f;
        std::memcpy(&f, &fb, 4);
        out[l] = f - 1.0f;
    }
}

// get_sample_4d (renderer.rs:698-725)
inline void get_sample_4d(uint32_t i, uint32_t dimension_set, uint32_t px, uint32_t py, uint32_t seed, float out[4]) {
    uint32_t s = px ^ (py << 16) ^ (seed * 0x736caf6fu);
    if (dimension_set < 2) {
        sobol_sample_4d(i, dimension_set, s, out);
    } else {
        for (uint32_t k = 0; k < 4; ++k) out[k] = hash_u32_to_f32((dimension_set * 4 + k) ^ (i << 16), s);
    }
}

// golden_ratio_sample (renderer.rs:728-737)
inline float golden_ratio_sample(uint32_t i, uint32_t scramble) {
    uint32_t n = (i + hash_u32(scramble, 0)) * 2654435769u;
    return (float)n * (1.0f / 4294967296.0f);
}

// probit (math.rs:67-132)
inline float probit(float nf, float width) {
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
        double q = std::sqrt(-2.0 * m_log(n));
        x = (((((C1 * q + C2) * q + C3) * q + C4) * q + C5) * q + C6) / ((((D1 * q + D2) * q + D3) * q + D4) * q + 1.0);
    } else if (n <= N_HIGH) {
        // also taken for n <= 0: the reference's match arms are tried in this order
        double q = n - 0.5;
        double r = q * q;
        x = (((((A1 * r + A2) * r + A3) * r + A4) * r + A5) * r + A6) * q /
            (((((B1 * r + B2) * r + B3) * r + B4) * r + B5) * r + 1.0);
    } else if (n < 1.0) {
        double q = std::sqrt(-2.0 * m_log(1.0 - n));
        x = -(((((C1 * q + C2) * q + C3) * q + C4) * q + C5) * q + C6) / ((((D1 * q + D2) * q + D3) * q + D4) * q + 1.0);
    } else if (n == 1.0) {
        x = std::numeric_limits<double>::infinity();
    } else {
        x = std::numeric_limits<double>::quiet_NaN();
    }
    return (float)x * width;
}

// power_heuristic (mis.rs:12-21)
inline float power_heuristic(float a, float b) {
    if (std::isinf(a)) return a;
    float a2 = a * a, b2 = b * b;
    float mis_fac = a2 / (a2 + b2);
    return a / mis_fac;
}

// square_to_circle (monte_carlo.rs:10-32)
inline void square_to_circle(float x, float y, float& ox, float& oy) {
    const float QPI = 0.78539816339744830962f;
    if (x == 0.0f && y == 0.0f) {
        ox = 0.0f;
        oy = 0.0f;
        return;
    }
    float radius, angle;
    if (x > std::fabs(y)) {
        radius = x;
        angle = QPI * (y / x);
    } else if (y > std::fabs(x)) {
        radius = y;
        angle = QPI * (2.0f - (x / y));
    } else if (x < -std::fabs(y)) {
        radius = -x;
        angle = QPI * (4.0f + (y / x));
    } else {
        radius = -y;
        angle = QPI * (6.0f - (x / y));
    }
    ox = radius * m_cosf(angle);
    oy = radius * m_sinf(angle);
}
inline V3 cosine_sample_hemisphere(float u, float v) {
    float a, b;
    square_to_circle((u * 2.0f) - 1.0f, (v * 2.0f) - 1.0f, a, b);
    float z = std::sqrt(rmax(1.0f - ((a * a) + (b * b)), 0.0f));
    return mk(a, b, z);
}
inline V3 uniform_sample_sphere(float u, float v) {
    const float PI_32 = 3.14159265358979323846f;
    float z = 1.0f - (2.0f * u);
    float r = std::sqrt(rmax(1.0f - (z * z), 0.0f));
    float phi = 2.0f * PI_32 * v;
    return mk(r * m_cosf(phi), r * m_sinf(phi), z);
}
const double PI_64 = 3.14159265358979323846;
inline V3 uniform_sample_cone(float u, float v, double cos_theta_max) {
    double cos_theta = (1.0 - (double)u) + ((double)u * cos_theta_max);
    double sin_theta = std::sqrt(1.0 - (cos_theta * cos_theta));
    double phi = (double)v * 2.0 * PI_64;
    return mk((float)(m_cos(phi) * sin_theta), (float)(m_sin(phi) * sin_theta), (float)cos_theta);
}
inline double uniform_sample_cone_pdf(double cos_theta_max) { return 1.0 / (2.0 * PI_64 * (1.0 - cos_theta_max)); }
inline V3 uniform_sample_triangle(V3 va, V3 vb, V3 vc, float i, float j) {
    float isqrt = std::sqrt(i);
    float a = 1.0f - isqrt;
    float b = isqrt * (1.0f - j);
    float c = j * isqrt;
    return (va * a) + (vb * b) + (vc * c);
}
inline float triangle_surface_area(V3 p0, V3 p1, V3 p2) { return 0.5f * length(cross(p1 - p0, p2 - p0)); }

// spherical_triangle_solid_angle (monte_carlo.rs:97-131)
inline float spherical_triangle_solid_angle(V3 va, V3 vb, V3 vc) {
    double cos_a = (double)rmin(rmax(dot(vb, vc), -1.0f), 1.0f);
    double cos_b = (double)rmin(rmax(dot(vc, va), -1.0f), 1.0f);
    double cos_c = (double)rmin(rmax(dot(va, vb), -1.0f), 1.0f);
    double sin_a = std::sqrt(1.0 - (cos_a * cos_a));
    double sin_b = std::sqrt(1.0 - (cos_b * cos_b));
    double sin_c = std::sqrt(1.0 - (cos_c * cos_c));
    if (cos_a == 1.0 || cos_b == 1.0 || cos_c == 1.0) return 0.0f;
    double cos_va = rmin(rmax((cos_a - (cos_b * cos_c)) / (sin_b * sin_c), -1.0), 1.0);
    double cos_vb = rmin(rmax((cos_b - (cos_c * cos_a)) / (sin_c * sin_a), -1.0), 1.0);
    double cos_vc = rmin(rmax((cos_c - (cos_a * cos_b)) / (sin_a * sin_b), -1.0), 1.0);
    double ang_va = m_acos(cos_va), ang_vb = m_acos(cos_vb), ang_vc = m_acos(cos_vc);
    return (float)(ang_va + ang_vb + ang_vc - PI_64);
}

// uniform_sample_spherical_triangle (monte_carlo.rs:133-199)
inline V3 uniform_sample_spherical_triangle(V3 va, V3 vb, V3 vc, float i, float j) {
    double cos_a = (double)rmin(rmax(dot(vb, vc), -1.0f), 1.0f);
    double cos_b = (double)rmin(rmax(dot(vc, va), -1.0f), 1.0f);
    double cos_c = (double)rmin(rmax(dot(va, vb), -1.0f), 1.0f);
    double sin_a = std::sqrt(1.0 - (cos_a * cos_a));
    double sin_b = std::sqrt(1.0 - (cos_b * cos_b));
    double sin_c = std::sqrt(1.0 - (cos_c * cos_c));
    if (cos_a == 1.0 || cos_b == 1.0 || cos_c == 1.0) return va;
    double cos_va = rmin(rmax((cos_a - (cos_b * cos_c)) / (sin_b * sin_c), -1.0), 1.0);
    double cos_vb = rmin(rmax((cos_b - (cos_c * cos_a)) / (sin_c * sin_a), -1.0), 1.0);
    double cos_vc = rmin(rmax((cos_c - (cos_a * cos_b)) / (sin_a * sin_b), -1.0), 1.0);
    double sin_va = std::sqrt(1.0 - (cos_va * cos_va));
    double ang_va = m_acos(cos_va), ang_vb = m_acos(cos_vb), ang_vc = m_acos(cos_vc);
    double area = ang_va + ang_vb + ang_vc - PI_64;
    double area_2 = area * (double)i;
    double s = m_sin(area_2 - ang_va);
    double t = m_cos(area_2 - ang_va);
    double u = t - cos_va;
    double v = s + (sin_va * cos_c);
    double q_top = (((v * t) - (u * s)) * cos_va) - v;
    double q_bottom = ((v * s) + (u * t)) * sin_va;
    double q = q_top / q_bottom;
    V3 vc_2 = (va * (float)q) + (normalized(vc - (va * dot(vc, va))) * (float)std::sqrt(1.0 - (q * q)));
    float z = 1.0f - (j * (1.0f - dot(vc_2, vb)));
    return (vb * z) + (normalized(vc_2 - (vb * dot(vc_2, vb))) * std::sqrt(1.0f - (z * z)));
}

// fastapprox::fast::pow2 / exp (src/math.rs:134-160 wrappers)
inline float fast_pow2(float p) {
    float offset = (p < 0.0f) ? 1.0f : 0.0f;
    float clipp = (p < -126.0f) ? -126.0f : p;
    int w = (int)clipp;
    float z = clipp - (float)w + offset;
    uint32_t bits = (uint32_t)((float)(1 << 23) * (clipp + 121.2740575f + 27.7280233f / (4.84252568f - z) - 1.49012907f * z));
    return u2f(bits);
}
inline float fast_exp(float p) { return fast_pow2(1.442695040f * p); }

}  // namespace orc
