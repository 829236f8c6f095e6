// Device functions of the physically based shading stage: BRDF, light-index permutation,
// CDF resampling walks, secondary-ray tracing.
//
// Replaces (reference file:line):
//   MultiLobe.eval / Lambertian / GGX      lib/torch_pbr/bxdf.py:111-146, 217-265, 321-330
//   warp_utils helpers                      lib/torch_pbr/utils/warp_utils.py:62-101, 693-702, 730-747, 782-794
//   cdf_resampling_merge_kernel             lib/nerfacc/cuda/csrc/cdf.cu:217-334
//   cdf_resampling_sdf_fine_kernel          lib/nerfacc/cuda/csrc/cdf.cu:536-638
//   light-index shuffle                     models/intrinsic_avatar.py:1355-1378 (stateless permutation)
#pragma once
#include "ia_device.cuh"

// ------------------------------------------------------------------------------------------------
// MultiLobe.eval with attenuation 0.  wi: towards viewer, wo: towards light, n: unit normal.
// diff [1] and spec [3] include the cosine factor.
__device__ __forceinline__ void ia_brdf_multilobe(const float wi[3], const float n[3], const float wo[3],
                                                  float rough, const float albedo[3], float metal, float& diff,
                                                  float spec[3]) {
    const float eps = 1e-6f;
    const float PI = 3.14159265358979323846f;
    float cos_o = wo[0] * n[0] + wo[1] * n[1] + wo[2] * n[2];
    diff = fmaxf(cos_o, 0.f) / PI;
    // local frame (coordinate_system): c = tangent, b = cross(c, n)
    float c[3];
    if (fabsf(n[0]) > fabsf(n[1])) {
        float inv = 1.0f / sqrtf(n[0] * n[0] + n[2] * n[2]);
        c[0] = n[2] * inv; c[1] = 0.f; c[2] = -n[0] * inv;
    } else {
        float inv = 1.0f / sqrtf(n[1] * n[1] + n[2] * n[2]);
        c[0] = 0.f; c[1] = n[2] * inv; c[2] = -n[1] * inv;
    }
    float b[3] = {c[1] * n[2] - c[2] * n[1], c[2] * n[0] - c[0] * n[2], c[0] * n[1] - c[1] * n[0]};
    float wol[3] = {wo[0] * b[0] + wo[1] * b[1] + wo[2] * b[2], wo[0] * c[0] + wo[1] * c[1] + wo[2] * c[2], cos_o};
    float wil[3] = {wi[0] * b[0] + wi[1] * b[1] + wi[2] * b[2], wi[0] * c[0] + wi[1] * c[1] + wi[2] * c[2],
                    wi[0] * n[0] + wi[1] * n[1] + wi[2] * n[2]};
    float wh[3] = {wil[0] + wol[0], wil[1] + wol[1], wil[2] + wol[2]};
    float nh = fmaxf(sqrtf(wh[0] * wh[0] + wh[1] * wh[1] + wh[2] * wh[2]), 1e-12f);  // F.normalize eps
    wh[0] /= nh; wh[1] /= nh; wh[2] /= nh;
    spec[0] = spec[1] = spec[2] = 0.f;
    if (wil[2] > eps && wol[2] > eps) {
        float alpha = rough;
        float k = (alpha * alpha + 2.f * alpha + 1.f) / 8.0f;
        float cos2 = wh[2] * wh[2];
        float alpha2 = alpha * alpha;
        float t = cos2 * (alpha2 - 1.f) + 1.f;
        float ndf = alpha2 * (1.0f / (PI * t * t + eps));
        float den_i = wil[2] * (1.0f - k) + k, den_o = wol[2] * (1.0f - k) + k;
        float g1i = den_i > eps ? wil[2] / (den_i + eps) : 0.f;
        float g1o = den_o > eps ? wol[2] / (den_o + eps) : 0.f;
        float cih = fabsf(wil[0] * wh[0] + wil[1] * wh[1] + wil[2] * wh[2]);
        float fr = exp2f((-5.55473f * cih - 6.98316f) * cih);
        float common = ndf * g1i * g1o;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float F0 = 0.04f * (1.0f - metal) + albedo[ch] * metal;
            float F = F0 + (1.0f - F0) * fr;
            spec[ch] = common * F / (4.f * wil[2] + eps);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// BSDF importance sampling and its pdf (render_mode = mats | mis): MultiLobe.sample / MultiLobe.pdf in eval
// mode (lib/torch_pbr/bxdf.py:290-388), GGX.pdf / GGX.sample (:222-236, 267-281), Lambertian (:117-139),
// sample_GGX_VNDF, eval_GGX_VNDF, sample_Lambertian_surface, sample_uniform_disk_concentric
// (lib/torch_pbr/utils/warp_utils.py:632-690, 797-815, 599-616, 139-172).
__device__ __forceinline__ void ia_coord_frame(const float n[3], float t[3], float b[3]) {
    // coordinate_system (warp_utils.py:62-101): b = the axis built from n, t = cross(b, n)
    if (fabsf(n[0]) > fabsf(n[1])) {
        float inv = 1.0f / sqrtf(n[0] * n[0] + n[2] * n[2]);
        b[0] = n[2] * inv; b[1] = 0.f; b[2] = -n[0] * inv;
    } else {
        float inv = 1.0f / sqrtf(n[1] * n[1] + n[2] * n[2]);
        b[0] = 0.f; b[1] = n[2] * inv; b[2] = -n[1] * inv;
    }
    t[0] = b[1] * n[2] - b[2] * n[1]; t[1] = b[2] * n[0] - b[0] * n[2]; t[2] = b[0] * n[1] - b[1] * n[0];
}
__device__ __forceinline__ float ia_dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void ia_normalize12(float v[3]) {  // F.normalize(dim=-1), eps 1e-12
    float nrm = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12f);
    v[0] /= nrm; v[1] /= nrm; v[2] /= nrm;
}
__device__ __forceinline__ float ia_luminance(const float x[3]) { return x[0] * 0.212671f + x[1] * 0.715160f + x[2] * 0.072169f; }

// lobe selection weights shared by pdf and sample (bxdf.py:297-312, 343-358); Fresnel with F0 = albedo
__device__ __forceinline__ void ia_lobe_weights(const float wi[3], const float n[3], const float albedo[3], float metal,
                                                float& wd, float& ws) {
    wd = (1.0f - metal) * ia_luminance(albedo);
    const float cos_t = ia_dot3(wi, n);
    const float fr = exp2f((-5.55473f * cos_t - 6.98316f) * cos_t);
    const float fres[3] = {albedo[0] + (1.0f - albedo[0]) * fr, albedo[1] + (1.0f - albedo[1]) * fr,
                           albedo[2] + (1.0f - albedo[2]) * fr};
    ws = cos_t > 0.f ? ia_luminance(fres) : 0.f;
}

__device__ __forceinline__ float ia_multilobe_pdf(const float wi[3], const float n[3], const float wo[3], float rough,
                                                  const float albedo[3], float metal) {
    const float eps = 1e-6f;
    const float PI = 3.14159265358979323846f;
    float wd, ws;
    ia_lobe_weights(wi, n, albedo, metal, wd, ws);
    const float p_d = (wd + ws > eps) ? wd / (wd + ws + eps) : 1.0f;
    const float pdf_d = fmaxf(ia_dot3(n, wo), 0.f) / PI;
    float t[3], b[3];
    ia_coord_frame(n, t, b);
    const float wol[3] = {ia_dot3(wo, t), ia_dot3(wo, b), ia_dot3(wo, n)};
    const float wil[3] = {ia_dot3(wi, t), ia_dot3(wi, b), ia_dot3(wi, n)};
    float wh[3] = {wil[0] + wol[0], wil[1] + wol[1], wil[2] + wol[2]};
    ia_normalize12(wh);
    const float alpha = rough;
    const float k = (alpha * alpha + 2.f * alpha + 1.f) / 8.0f;
    const float den = wil[2] * (1.0f - k) + k;
    const float g1 = den > eps ? wil[2] / (den + eps) : 0.f;
    const float alpha2 = alpha * alpha;
    const float tt = wh[2] * wh[2] * (alpha2 - 1.f) + 1.f;
    const float ndf = alpha2 * (1.0f / (PI * tt * tt + eps));
    const float vndf = (wh[2] > eps && wil[2] > eps) ? g1 * fmaxf(ia_dot3(wh, wil), 0.f) * ndf / (wil[2] + eps) : 0.f;
    const float pdf_s = (4.f * fabsf(ia_dot3(wil, wh)) > eps) ? vndf / (4.f * fabsf(ia_dot3(wol, wh)) + eps) : 0.f;
    return p_d * pdf_d + (1.0f - p_d) * pdf_s;
}

// u0, u1: the two uniforms of `sample`; u0 picks the lobe and is rescaled (bxdf.py:360-369)
__device__ __forceinline__ void ia_multilobe_sample(const float n[3], const float wi[3], float rough, const float albedo[3],
                                                    float metal, float u0, float u1, float wo[3]) {
    const float eps = 1e-6f;
    const float PI = 3.14159265358979323846f;
    float wd, ws;
    ia_lobe_weights(wi, n, albedo, metal, wd, ws);
    const float p_s = (wd + ws > eps) ? ws / (wd + ws + eps) : 0.f;
    float t[3], b[3];
    ia_coord_frame(n, t, b);
    float v[3];
    if (p_s > u0) {
        // GGX.sample: visible-normal sampling in the local frame, then reflect wi about the sampled normal
        const float s0 = u0 / p_s;
        const float wil[3] = {ia_dot3(wi, t), ia_dot3(wi, b), ia_dot3(wi, n)};
        float vh[3] = {rough * wil[0], rough * wil[1], wil[2]};
        ia_normalize12(vh);
        const float lensq = vh[0] * vh[0] + vh[1] * vh[1];
        float T1[3] = {1.f, 0.f, 0.f};
        if (lensq > eps) { T1[0] = -vh[1] / sqrtf(lensq + eps); T1[1] = vh[0] / sqrtf(lensq + eps); T1[2] = 0.f; }
        const float T2[3] = {vh[1] * T1[2] - vh[2] * T1[1], vh[2] * T1[0] - vh[0] * T1[2], vh[0] * T1[1] - vh[1] * T1[0]};
        const float r = sqrtf(s0);
        const float phi = 2.0f * PI * u1;
        const float t1 = r * cosf(phi);
        float t2 = r * sinf(phi);
        const float sv = 0.5f * (1.0f + vh[2]);
        t2 = (1.0f - sv) * sqrtf(fmaxf(1.0f - t1 * t1, 0.f)) + sv * t2;
        const float t3 = sqrtf(fmaxf(1.0f - t1 * t1 - t2 * t2, 0.f));
        const float nh[3] = {t1 * T1[0] + t2 * T2[0] + t3 * vh[0], t1 * T1[1] + t2 * T2[1] + t3 * vh[1],
                             t1 * T1[2] + t2 * T2[2] + t3 * vh[2]};
        float wh[3] = {rough * nh[0], rough * nh[1], fmaxf(nh[2], 0.f)};
        ia_normalize12(wh);
        const float dp = 2.f * ia_dot3(wil, wh);
        v[0] = dp * wh[0] - wil[0]; v[1] = dp * wh[1] - wil[1]; v[2] = dp * wh[2] - wil[2];
    } else {
        // Lambertian.sample (eval): concentric disk -> cosine-weighted hemisphere
        const float s0 = (u0 - p_s) / (1.0f - p_s);
        const float ox = 2.0f * s0 - 1.0f, oy = 2.0f * u1 - 1.0f;
        const bool big = fabsf(ox) > fabsf(oy);
        const float rr = big ? ox : oy;
        const float th = big ? PI / 4.0f * (oy / ox) : PI / 2.0f - PI / 4.0f * (ox / oy);
        v[0] = rr * cosf(th); v[1] = rr * sinf(th);
        v[2] = sqrtf(fmaxf(1.0f - v[0] * v[0] - v[1] * v[1], 0.f));
    }
    // to_world (normalised)
    wo[0] = v[0] * t[0] + v[1] * b[0] + v[2] * n[0];
    wo[1] = v[0] * t[1] + v[1] * b[1] + v[2] * n[1];
    wo[2] = v[0] * t[2] + v[1] * b[2] + v[2] * n[2];
    ia_normalize12(wo);
}

// Counter-based uniforms in [0,1) with 24 bits (the reference draws torch.rand on the device): stream `dim` of
// shading sample j of the pixel whose key is `key` (= ia_pixel_key(seed, ray index)).  Mirrored by oracle/pbr.py.
__device__ __forceinline__ uint32_t ia_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float ia_rng_uniform(uint32_t key, uint32_t j, uint32_t dim) {
    uint32_t x = ia_mix32(key ^ ia_mix32(j * 0x9E3779B9u + dim * 0x85EBCA6Bu + 0x6A09E667u));
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}

// ------------------------------------------------------------------------------------------------
// Environment light look-ups per direction (EnvironmentLightTensor.sample / pdf / eval,
// lib/torch_pbr/light.py:259-446; light_utils.py:6-63 with xyz2lonlat_mode = null).
struct IaEnv {
    const float* env;   // [H,W,3]
    const float* pdf;   // [H,W] normalised
    const float* cols;  // [H,W+1]
    const float* rows;  // [H+1]
    int H, W;
};
__device__ __forceinline__ int ia_searchsorted_right(const float* a, int n, float v) {  // first i with a[i] > v
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ void ia_env_uv(const float d[3], float& u, float& v, float& lat) {
    const float PI = 3.14159265358979323846f;
    float lon = atan2f(d[0], d[2]);
    float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    lat = asinf(d[1] / nrm);
    u = lon / (2 * PI) + 0.5f;
    v = lat / PI + 0.5f;
}
// sample(): inverse CDF with linear in-bin offset (light.py:341-412) -> unit WORLD direction
__device__ __forceinline__ void ia_env_sample(const IaEnv& E, float u1, float u2, float dn[3]) {
    const float PI = 3.14159265358979323846f;
    const int H = E.H, W = E.W;
    int ri = ia_searchsorted_right(E.rows, H + 1, u1);
    int below = max(ri - 1, 0), above = min(ri, H);
    float rfrac = (u1 - E.rows[below]) / (E.rows[above] - E.rows[below]);
    ri = below;
    const float* crow = E.cols + (size_t)ri * (W + 1);
    int ci = ia_searchsorted_right(crow, W + 1, u2);
    below = max(ci - 1, 0); above = min(ci, W);
    float cfrac = (u2 - crow[below]) / (crow[above] - crow[below]);
    ci = below;
    float uu = ((float)ci + cfrac) / (float)W, vv = ((float)ri + rfrac) / (float)H;
    float lon = (uu - 0.5f) * 2 * PI, lat = (vv - 0.5f) * PI;
    float d[3] = {cosf(lat) * sinf(lon), sinf(lat), cosf(lat) * cosf(lon)};
    ia_normalize(d, dn, 1e-12f);
}
// pdf() (light.py:259-296) of a WORLD direction
__device__ __forceinline__ float ia_env_pdf(const IaEnv& E, const float dw[3]) {
    const float PI = 3.14159265358979323846f;
    const int H = E.H, W = E.W;
    float u, v, la;
    ia_env_uv(dw, u, v, la);
    int col = (int)fminf(fmaxf(floorf(u * (float)W), 0.f), (float)(W - 1));
    int row = (int)fminf(fmaxf(floorf(v * (float)H), 0.f), (float)(H - 1));
    float sin_theta = sinf(PI / 2.0f - la);
    float pdf_scale = (float)((double)H * (double)W / (2.0 * 3.14159265358979323846 * 3.14159265358979323846));
    return sin_theta > 0 ? E.pdf[(size_t)row * W + col] * pdf_scale / sin_theta : 0.f;
}
// eval(): bilinear grid_sample, align_corners=True, border (light.py:298-339) of a WORLD direction
__device__ __forceinline__ void ia_env_eval(const IaEnv& E, const float dw[3], float em[3]) {
    const int H = E.H, W = E.W;
    float u, v, la;
    ia_env_uv(dw, u, v, la);
    float fx = fminf(fmaxf(((u * 2 - 1) + 1.f) / 2 * (W - 1), 0.f), (float)(W - 1));
    float fy = fminf(fmaxf(((v * 2 - 1) + 1.f) / 2 * (H - 1), 0.f), (float)(H - 1));
    int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    float wx = fx - x0, wy = fy - y0;
    int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float* env = E.env;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float a = env[((size_t)y0 * W + x0) * 3 + ch], b = env[((size_t)y0 * W + x1) * 3 + ch];
        float c2 = env[((size_t)y1 * W + x0) * 3 + ch], d2 = env[((size_t)y1 * W + x1) * 3 + ch];
        em[ch] = a * (1 - wx) * (1 - wy) + b * wx * (1 - wy) + c2 * (1 - wx) * wy + d2 * wx * wy;
    }
}

// ------------------------------------------------------------------------------------------------
// Kensler's keyed permutation of [0, l) ("Correlated Multi-Jittered Sampling", 2013) and key mixing.
__device__ __forceinline__ uint32_t ia_permute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8; i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1; i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2; i *= 0x9e501cc3u;
        i ^= (i & w) >> 2; i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}
__device__ __forceinline__ uint32_t ia_pixel_key(uint32_t seed, uint64_t ray_index) {
    uint32_t x = (uint32_t)(ray_index * 0x9E3779B1ull + seed);
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

// ------------------------------------------------------------------------------------------------
// cdf_resampling_merge_kernel body for one ray (cdf.cu:217-334).  flags: bit0 = is_left, bit1 = is_right.
// Writes the merged edge list compactly (exactly the slots the reference marks is_fg_sample) and
// returns its length.  out_flags bit2 = is_resample.
__device__ __forceinline__ int ia_merge_resample(const float* vals, const uint8_t* flags, const float* weights,
                                                 int steps, int n_new, float* ovals, uint8_t* oflags,
                                                 float* odists) {
    float weights_sum = 0.0f;
    for (int j = 0; j < steps - 1; j++)
        weights_sum += ((flags[j] & 1) && (flags[j + 1] & 2)) ? weights[j] : 0.0f;
    weights_sum += fmaxf(1.0f - weights_sum, 0.0f);
    int num_bins = n_new;
    float cdf_step_size = (1.0f - 1.0 / num_bins) / (n_new - 1);
    int idx = 0, j = 0;
    float cdf_prev = 0.0f, cdf_next = weights[idx] / weights_sum;
    float cdf_u = 1.0 / (2 * num_bins);
    float start = vals[0], end = vals[1];
    ovals[0] = start;
    oflags[0] = 1;
    while (j < num_bins && idx < steps - 1) {
        if (cdf_u < cdf_next) {
            float scaling = (end - start) / (cdf_next - cdf_prev);
            float offset = (cdf_u - cdf_prev) * scaling;
            float t = offset + start;
            cdf_u += cdf_step_size;
            if (odists) odists[j + idx] = t - ovals[j + idx];
            j += 1;
            ovals[j + idx] = t;
            oflags[j + idx] = 1 | 2 | 4;
        } else {
            if (odists) odists[j + idx] = end - ovals[j + idx];
            idx += 1;
            ovals[j + idx] = end;
            oflags[j + idx] = (flags[idx] & 2);
            if (idx >= steps - 1) break;
            start = vals[idx];
            end = vals[idx + 1];
            if ((flags[idx] & 1) && (flags[idx + 1] & 2)) {
                cdf_prev = cdf_next;
                cdf_next += weights[idx] / weights_sum;
                oflags[j + idx] |= 1;
            }
        }
    }
    while (idx < steps - 1) {
        if (odists) odists[j + idx] = end - ovals[j + idx];
        idx += 1;
        ovals[j + idx] = end;
        oflags[j + idx] = (flags[idx] & 2);
        if (idx >= steps - 1) break;
        start = vals[idx];
        end = vals[idx + 1];
        if ((flags[idx] & 1) && (flags[idx + 1] & 2)) oflags[j + idx] |= 1;
    }
    return j + idx + 1;
}
