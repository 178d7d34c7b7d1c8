// oracle/oshade.h — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Hero-wavelength spectral colour, surface closures and the light sampling/intersection math,
// restated from:
//   src/color.rs:15-43,84-142,336-383,520-526,602-628
//   sub_crates/spectral_upsampling/src/meng.rs:177-328 (spectrum_xyz_to_p_4; tables/meng_tables.inc)
//   src/shading/surface_closure.rs:57-155,281-633,714-719
//   src/light/sphere_light.rs:51-192, rectangle_light.rs:52-243, distant_disk_light.rs:59-85
#pragma once
#include <cstdio>
#include "osample.h"

#define MENG_TABLE_QUAL static const
#include "../tables/meng_tables.inc"

namespace orc {

struct V4 {
    float v[4];
};
inline V4 splat4(float a) { return V4{{a, a, a, a}}; }
inline V4 operator*(V4 a, float s) { return V4{{a.v[0] * s, a.v[1] * s, a.v[2] * s, a.v[3] * s}}; }
inline V4 operator*(V4 a, V4 b) { return V4{{a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]}}; }
inline V4 operator/(V4 a, float s) { return V4{{a.v[0] / s, a.v[1] / s, a.v[2] / s, a.v[3] / s}}; }
inline V4 operator+(V4 a, V4 b) { return V4{{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2], a.v[3] + b.v[3]}}; }
// Vec4::max_element (horizontal maxps)
inline float max_element(V4 a) { return sse_max(sse_max(a.v[0], a.v[2]), sse_max(a.v[1], a.v[3])); }

const float WL_MIN = 380.0f, WL_MAX = 700.0f, WL_RANGE = WL_MAX - WL_MIN, WL_RANGE_Q = WL_RANGE / 4.0f;
inline float map_0_1_to_wavelength(float n) { return n * WL_RANGE + WL_MIN; }
inline float nth_wavelength(float hero, int n) {
    float wl = hero + (WL_RANGE_Q * (float)n);
    return wl > WL_MAX ? wl - WL_RANGE : wl;
}

// spectrum_xyz_to_p_4 (meng.rs:177-328)
inline V4 spectrum_xyz_to_p_4(const float lambdas[4], float x, float y, float z) {
    float inv_norm = x + y + z;
    float norm = 1.0f / inv_norm;
    if (!(norm < std::numeric_limits<float>::max())) return splat4(0.0f);
    float xyy0 = x * norm, xyy1 = y * norm;
    float uv0 = MENG_MAT_XY_TO_UV[0] * xyy0 + MENG_MAT_XY_TO_UV[1] * xyy1 + MENG_MAT_XY_TO_UV[2];
    float uv1 = MENG_MAT_XY_TO_UV[3] * xyy0 + MENG_MAT_XY_TO_UV[4] * xyy1 + MENG_MAT_XY_TO_UV[5];
    if (uv0 < 0.0f || uv0 >= (float)MENG_GRID_WIDTH || uv1 < 0.0f || uv1 >= (float)MENG_GRID_HEIGHT) return splat4(0.0f);
    int uvi0 = (int)uv0, uvi1 = (int)uv1;
    int cell_idx = uvi0 + MENG_GRID_WIDTH * uvi1;
    const short* cell = MENG_GRID[cell_idx];
    bool inside = cell[0] != 0;
    int num = cell[1];
    const short* idx = cell + 2;
    if (num == 0) return splat4(0.0f);

    float sb[4], sbf[4];
    int sb0[4], sb1[4];
    for (int k = 0; k < 4; ++k) {
        sb[k] = (lambdas[k] - MENG_SAMPLE_MIN) / (MENG_SAMPLE_MAX - MENG_SAMPLE_MIN) * ((float)MENG_NUM_SAMPLES - 1.0f);
        sb0[k] = (int)sb[k];
        sb1[k] = sb0[k] + 1 < MENG_NUM_SAMPLES - 1 ? sb0[k] + 1 : MENG_NUM_SAMPLES - 1;
        sbf[k] = sb[k] - (float)sb0[k];
    }
    V4 p[6];
    for (int i = 0; i < 6; ++i) p[i] = splat4(0.0f);
    for (int i = 0; i < num; ++i) {
        const float* spectrum = MENG_SPECTRA[idx[i]];
        for (int k = 0; k < 4; ++k) p[i].v[k] = spectrum[sb0[k]] * (1.0f - sbf[k]) + spectrum[sb1[k]] * sbf[k];
    }
    V4 interpolated = splat4(0.0f);
    if (inside) {
        float u2 = uv0 - (float)uvi0, v2 = uv1 - (float)uvi1;
        interpolated = p[0] * ((1.0f - u2) * (1.0f - v2)) + p[2] * ((1.0f - u2) * v2) + p[3] * (u2 * v2) +
                       p[1] * (u2 * (1.0f - v2));
    } else {
        float ex = uv0 - MENG_UV[idx[0]][0];
        float ey = uv1 - MENG_UV[idx[0]][1];
        float e0x = MENG_UV[idx[1]][0] - MENG_UV[idx[0]][0];
        float e0y = MENG_UV[idx[1]][1] - MENG_UV[idx[0]][1];
        float uu = e0x * ey - ex * e0y;
        for (int i = 0; i < num - 1; ++i) {
            float e1x, e1y;
            if (i == num - 2) {
                e1x = MENG_UV[idx[1]][0] - MENG_UV[idx[0]][0];
                e1y = MENG_UV[idx[1]][1] - MENG_UV[idx[0]][1];
            } else {
                e1x = MENG_UV[idx[i + 2]][0] - MENG_UV[idx[0]][0];
                e1y = MENG_UV[idx[i + 2]][1] - MENG_UV[idx[0]][1];
            }
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
            interpolated = p[0] * w + p[i + 1] * v + p[(i == num - 2) ? 1 : i + 2] * u;
            break;
        }
    }
    return interpolated * inv_norm;
}

// plancks_law (color.rs:336-375)
inline float plancks_law(float temperature, float wavelength) {
    const float C = 299792458.0f, H = 6.62607015e-34f, KB = 1.38064852e-23f;
    if (temperature < 200.0f) return 0.0f;
    float wl = wavelength * 1.0e-9f;
    const float TMP1 = (float)(2.0 * (double)H * (double)C * (double)C);
    const float TMP2 = (float)((double)H * (double)C / (double)KB);
    float wl2 = wl * wl;
    float wl5 = wl2 * wl2 * wl;
    float tmp3 = wl5 * (fast_exp(TMP2 / (wl * temperature)) - 1.0f);
    float energy = TMP1 / tmp3;
    return rmax(energy * 1.0e-6f, 0.0f);
}
inline float plancks_law_normalized(float temperature, float wavelength) {
    float t2 = temperature * temperature;
    return plancks_law(temperature, wavelength) * 4.0e7f / (t2 * t2);
}

// Color::to_spectral_sample (color.rs:84-122)
inline V4 color_to_spectral(const PsyColor& c, float hero) {
    float wls[4];
    for (int i = 0; i < 4; ++i) wls[i] = nth_wavelength(hero, i);
    if (c.kind == PSY_COLOR_XYZ) {
        return spectrum_xyz_to_p_4(wls, c.a, c.b, c.c) * splat4(1.0f / MENG_EQUAL_ENERGY_REFLECTANCE);
    } else if (c.kind == PSY_COLOR_BLACKBODY) {
        return V4{{plancks_law(c.a, wls[0]) * c.b, plancks_law(c.a, wls[1]) * c.b, plancks_law(c.a, wls[2]) * c.b,
                   plancks_law(c.a, wls[3]) * c.b}};
    }
    return V4{{plancks_law_normalized(c.a, wls[0]) * c.b, plancks_law_normalized(c.a, wls[1]) * c.b,
               plancks_law_normalized(c.a, wls[2]) * c.b, plancks_law_normalized(c.a, wls[3]) * c.b}};
}
// Color * f32 (color.rs:277-303)
inline PsyColor color_scale(const PsyColor& c, float s) {
    if (c.kind == PSY_COLOR_XYZ) return PsyColor{c.kind, c.a * s, c.b * s, c.c * s};
    return PsyColor{c.kind, c.a, c.b * s, c.c};
}
// Lerp for Color (color.rs:311-360); kinds are assumed equal (the reference panics otherwise)
inline PsyColor lerp(const PsyColor& x, const PsyColor& y, float alpha) {
    float inv = 1.0f - alpha;
    if (x.kind == PSY_COLOR_XYZ)
        return PsyColor{x.kind, (x.a * inv) + (y.a * alpha), (x.b * inv) + (y.b * alpha), (x.c * inv) + (y.c * alpha)};
    return PsyColor{x.kind, (x.a * inv) + (y.a * alpha), (x.b * inv) + (y.b * alpha), 0.0f};
}

// CIE 1931 analytic fits (color.rs:610-628)
inline float x_1931(float wl) {
    float t1 = (wl - 442.0f) * ((wl < 442.0f) ? 0.0624f : 0.0374f);
    float t2 = (wl - 599.8f) * ((wl < 599.8f) ? 0.0264f : 0.0323f);
    float t3 = (wl - 501.1f) * ((wl < 501.1f) ? 0.0490f : 0.0382f);
    return (0.362f * fast_exp(-0.5f * t1 * t1)) + (1.056f * fast_exp(-0.5f * t2 * t2)) - (0.065f * fast_exp(-0.5f * t3 * t3));
}
inline float y_1931(float wl) {
    float t1 = (wl - 568.8f) * ((wl < 568.8f) ? 0.0213f : 0.0247f);
    float t2 = (wl - 530.9f) * ((wl < 530.9f) ? 0.0613f : 0.0322f);
    return (0.821f * fast_exp(-0.5f * t1 * t1)) + (0.286f * fast_exp(-0.5f * t2 * t2));
}
inline float z_1931(float wl) {
    float t1 = (wl - 437.0f) * ((wl < 437.0f) ? 0.0845f : 0.0278f);
    float t2 = (wl - 459.0f) * ((wl < 459.0f) ? 0.0385f : 0.0725f);
    return (1.217f * fast_exp(-0.5f * t1 * t1)) + (0.681f * fast_exp(-0.5f * t2 * t2));
}
// XYZ::from_spectral_sample (color.rs:520-526)
inline void xyz_from_spectral(V4 e, float hero, float out[3]) {
    float X[4], Y[4], Z[4];
    for (int i = 0; i < 4; ++i) {
        float wl = nth_wavelength(hero, i);
        X[i] = x_1931(wl) * e.v[i];
        Y[i] = y_1931(wl) * e.v[i];
        Z[i] = z_1931(wl) * e.v[i];
    }
    out[0] = (((X[0] + X[1]) + X[2]) + X[3]) * 0.75f;
    out[1] = (((Y[0] + Y[1]) + Y[2]) + Y[3]) * 0.75f;
    out[2] = (((Z[0] + Z[1]) + Z[2]) + Z[3]) * 0.75f;
}

//------------------------------------------------------------------------------------------------
// Surface closures (surface_closure.rs)

struct Closure {
    uint32_t kind;  // PSY_SHADER_*
    PsyColor color;
    float roughness, fresnel;
};

const float PI_32 = 3.14159265358979323846f;
const float INV_PI = 1.0f / PI_32;
const float H_PI = PI_32 / 2.0f;

inline float ggx_d(float nh, float rough) {
    if (nh <= 0.0f) return 0.0f;
    float rough2 = rough * rough;
    float tmp = 1.0f + ((rough2 - 1.0f) * (nh * nh));
    return rough2 / (PI_32 * tmp * tmp);
}
inline float ggx_g(float vh, float vn, float rough) {
    if ((vh * vn) <= 0.0f) return 0.0f;
    return 2.0f / (1.0f + std::sqrt(1.0f + rough * rough * (1.0f - vn * vn) / (vn * vn)));
}
inline float half_theta_sample(float u, float rough) {
    float rough2 = rough * rough;
    float top = 1.0f - u;
    float bottom = 1.0f + ((rough2 - 1.0f) * u);
    return std::sqrt(top / bottom);
}
inline float schlick_fresnel_from_fac(float fac, float c) {
    float c1 = 1.0f - c;
    float c2 = c1 * c1;
    return fac + ((1.0f - fac) * c1 * c2 * c2);
}

inline void flip_normals(V3 inc, V3 nor, V3 nor_g, V3& nn, V3& flipped_nor_g) {
    if (dot(nor_g, inc) <= 0.0f) {
        nn = normalized(nor);
        flipped_nor_g = nor_g;
    } else {
        nn = -normalized(nor);
        flipped_nor_g = -nor_g;
    }
}

// closure.evaluate -> (filter, pdf)
inline void closure_evaluate(const Closure& cl, V3 inc, V3 out, V3 nor, V3 nor_g, float wl, V4& filter, float& pdf) {
    if (cl.kind == PSY_SHADER_LAMBERT) {
        V3 nn, fng;
        flip_normals(inc, nor, nor_g, nn, fng);
        if (dot(fng, out) >= 0.0f) {
            float fac = rmax(dot(nn, normalized(out)), 0.0f) * INV_PI;
            filter = color_to_spectral(cl.color, wl) * fac;
            pdf = fac;
        } else {
            filter = splat4(0.0f);
            pdf = 0.0f;
        }
    } else if (cl.kind == PSY_SHADER_GGX) {
        V3 aa = -normalized(inc);
        V3 bb = normalized(out);
        V3 hh = normalized(aa + bb);
        V3 nn, fng;
        flip_normals(inc, nor, nor_g, nn, fng);
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
        V4 ss = color_to_spectral(cl.color, wl);
        float rev_fresnel = 1.0f - cl.fresnel;
        V4 col_f;
        for (int i = 0; i < 4; ++i) col_f.v[i] = lerp(schlick_fresnel_from_fac(ss.v[i], hb), ss.v[i], rev_fresnel);
        if (cl.roughness == 0.0f) {
            filter = col_f;
            pdf = 0.0f;
            return;
        }
        float dist = ggx_d(nh, cl.roughness) / na;
        float g1 = ggx_g(ha, na, cl.roughness);
        float g2 = ggx_g(hb, nb, cl.roughness);
        filter = col_f * (dist * g1 * g2) * INV_PI;
        pdf = dist * INV_PI;
    } else {
        filter = color_to_spectral(cl.color, wl);
        pdf = 1.0f;
    }
}

// closure.sample -> (dir, filter, pdf)
inline void closure_sample(const Closure& cl, V3 inc, V3 nor, V3 nor_g, float u, float v, float wl, V3& dir, V4& filter,
                           float& pdf) {
    if (cl.kind == PSY_SHADER_LAMBERT) {
        V3 nn, fng;
        flip_normals(inc, nor, nor_g, nn, fng);
        V3 d = cosine_sample_hemisphere(u, v);
        float p = d.z * INV_PI;
        V3 out = zup_to_vec(d, nn);
        dir = out;
        if (dot(fng, out) >= 0.0f) {
            filter = color_to_spectral(cl.color, wl) * p;
            pdf = p;
        } else {
            filter = splat4(0.0f);
            pdf = 0.0f;
        }
    } else if (cl.kind == PSY_SHADER_GGX) {
        V3 nn, fng;
        flip_normals(inc, nor, nor_g, nn, fng);
        float theta_cos = half_theta_sample(u, cl.roughness);
        float theta_sin = std::sqrt(1.0f - (theta_cos * theta_cos));
        float angle = v * PI_32 * 2.0f;
        V3 half_dir = mk(m_cosf(angle) * theta_sin, m_sinf(angle) * theta_sin, theta_cos);
        half_dir = normalized(zup_to_vec(half_dir, nn));
        V3 out = inc - (half_dir * 2.0f * dot(inc, half_dir));
        dir = out;
        if (dot(fng, out) >= 0.0f) {
            closure_evaluate(cl, inc, out, nor, nor_g, wl, filter, pdf);
        } else {
            filter = splat4(0.0f);
            pdf = 0.0f;
        }
    } else {
        dir = mk(0, 0, 0);
        filter = color_to_spectral(cl.color, wl);
        pdf = 1.0f;
    }
}

inline bool g_sphere_lambert_debug = false;
inline bool& sphere_lambert_debug() { return g_sphere_lambert_debug; }
inline float sphere_lambert(float nlcos, float rcos) {
    float nlsin = std::sqrt(1.0f - (nlcos * nlcos));
    float rsin2 = 1.0f - (rcos * rcos);
    float rsin = std::sqrt(rsin2);
    float ysin = rcos / nlsin;
    float ycos2 = 1.0f - (ysin * ysin);
    float ycos = std::sqrt(ycos2);
    float g = (-2.0f * nlsin * rcos * ycos) + H_PI - m_asinf(ysin) + (ysin * ycos);
    float h = nlcos * ((ycos * std::sqrt(rsin2 - ycos2)) + (rsin2 * m_asinf(ycos / rsin)));
    float nl = m_acosf(nlcos);
    float r = m_acosf(rcos);
    if (ORC_UNLIKELY(g_sphere_lambert_debug))
        std::fprintf(stderr, "[oracle dbg]       sphere_lambert nlcos %.9g rcos %.9g nl %.9g r %.9g  nl-(H_PI-r) %.3g  ysin %.9g ycos2 %.3g g %.9g h %.9g\n",
                     (double)nlcos, (double)rcos, (double)nl, (double)r, (double)nl - ((double)H_PI - (double)r), (double)ysin,
                     (double)ycos2, (double)g, (double)h);
    if (nl < (H_PI - r)) return nlcos * rsin2;
    if (nl < H_PI) return (nlcos * rsin2) + g - h;
    if (nl < (H_PI + r)) return (g + h) * INV_PI;
    return 0.0f;
}

// estimate_eval_over_sphere_light (surface_closure.rs:115-155, 334-408, 535-591)
inline float closure_estimate_sphere(const Closure& cl, V3 inc, V3 to_light_center, float light_radius_squared, V3 nor,
                                     V3 nor_g) {
    if (cl.kind == PSY_SHADER_LAMBERT) {
        float dist2 = length2(to_light_center);
        if (dist2 <= light_radius_squared) return rmin(light_radius_squared / dist2, 4.0f);
        float sin_theta_max2 = rmin(light_radius_squared / dist2, 1.0f);
        float cos_theta_max = std::sqrt(1.0f - sin_theta_max2);
        V3 v = normalized(to_light_center);
        V3 nn = dot(nor_g, inc) <= 0.0f ? normalized(nor) : -normalized(nor);
        float cos_nv = rmin(rmax(dot(nn, v), -1.0f), 1.0f);
        return sphere_lambert(cos_nv, cos_theta_max);
    } else if (cl.kind == PSY_SHADER_GGX) {
        float dist2 = length2(to_light_center);
        float sin_theta_max2 = rmin(light_radius_squared / dist2, 1.0f);
        float cos_theta_max = std::sqrt(1.0f - sin_theta_max2);
        V3 nn = dot(nor, inc) < 0.0f ? normalized(nor) : -normalized(nor);
        V3 aa = -normalized(inc);
        V3 bb = normalized(to_light_center);
        float theta = m_acosf(cos_theta_max);
        V3 hh = normalized(aa + bb);
        float nh = rclamp(dot(nn, hh), -1.0f, 1.0f);
        float fac = ggx_d(nh, rmin(1.0f, std::sqrt(cl.roughness) + (2.0f * theta / PI_32)));
        return fac * rmin(1.0f, 1.0f - cos_theta_max) * INV_PI;
    }
    return 0.0f;  // Emit: unimplemented!() in the reference, never reached
}

}  // namespace orc
