// Device functions of the physically based shading stage: BRDF, light-index permutation,
// CDF resampling walks, secondary-ray tracing.
//
// Replaces (reference file:line):
//   MultiLobe.eval / Lambertian / GGX      lib/torch_pbr/bxdf.py:111-146, 217-265, 321-330
//   warp_utils helpers                      lib/torch_pbr/utils/warp_utils.py:62-101, 693-702, 730-747, 782-794
//   cdf_resampling_merge_kernel             lib/nerfacc/cuda/csrc/cdf.cu:217-334
//   cdf_resampling_sdf_fine_kernel          lib/nerfacc/cuda/csrc/cdf.cu:536-638
//   compute_indirect_radiance               models/intrinsic_avatar.py:396-545 (lazy, per ray)
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

// ------------------------------------------------------------------------------------------------
// Secondary ray: compute_indirect_radiance for ONE ray, lazily (models/intrinsic_avatar.py:396-545):
// coarse samples are generated and queried in order only until the first +->- SDF crossing is found
// and the 5 fine points of cdf_resampling_sdf_fine_kernel (cdf.cu:536-638) are placed; the <=4 fine
// intervals are then rendered.  Results equal the reference's evaluate-everything formulation.
struct IaTraceCounters {
    unsigned q, qg, fetch, geo, rad, skin;
};

template <bool GI>
__device__ __forceinline__ void ia_team_trace(const Team& team, const IaFrame& p, const float* __restrict__ wgeo,
                                              const float* __restrict__ wmlp, const uint32_t* __restrict__ occ,
                                              const float o[3], const float d[3], float& T, float rgb[3],
                                              IaTraceCounters& cnt) {
    T = 1.0f;
    rgb[0] = rgb[1] = rgb[2] = 0.f;
    IaMarcher m;
    m.init(p, o, d, p.sec_near, p.sec_far, p.sec_step);
    float ts, te;
    bool cont;
    IaQuery q;
    auto query_sdf = [&](float t) {
        float x[3] = {o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t};
        ia_team_query<false>(team, p, wgeo, x, q);
        cnt.q++; cnt.fetch += q.n_fetch; cnt.geo += q.n_valid;
        return q.sdf;
    };
    if (!m.next(occ, p.occ_res, ts, te, cont)) return;
    float sdf_prev = query_sdf(ts);
    float cs = ts, ce = te;  // current interval of the CDF walk (the crossing interval once found)
    float sdf_cur = 0.f;
    bool found = false;
    while (m.next(occ, p.occ_res, ts, te, cont)) {
        sdf_cur = query_sdf(ts);
        if (sdf_prev >= 0 && sdf_cur < 0) { found = true; break; }
        sdf_prev = sdf_cur;
        cs = ts; ce = te;
    }
    if (!found) return;
    bool pending = true;  // (ts, te, sdf_cur) is the already-queried interval after the crossing one
    const int num_bins = 5;
    float cdf_step_size = (1.0f - 1.0 / num_bins) / 4;
    float tpl[5];
    int j = 0;
    float trans = 1.0f;
    float a = ia_alpha(sdf_prev, ce - cs, p.beta);
    float weight = a;
    trans *= (1.0f - a);
    float cdf_prev = 0.0f, cdf_next = weight;
    float cdf_u = 1.0 / (2 * num_bins);
    while (j < num_bins) {
        if (cdf_u < cdf_next) {
            float scaling = (ce - cs) / (cdf_next - cdf_prev);
            float t = (cdf_u - cdf_prev) * scaling + cs;
            tpl[j] = t;
            cdf_u += cdf_step_size;
            j += 1;
        } else {
            float s;
            if (pending) {
                cs = ts; ce = te; s = sdf_cur;
                pending = false;
            } else {
                if (!m.next(occ, p.occ_res, cs, ce, cont)) break;
                s = query_sdf(cs);
            }
            a = ia_alpha(s, ce - cs, p.beta);
            weight = trans * a;
            trans *= (1.0f - a);
            cdf_prev = cdf_next;
            cdf_next += weight;
        }
    }
    // rendering() over the fine intervals (models/volrend.py:135-187)
    float Tacc = 1.0f, acc = 0.f;
    float view_w[3];
    if (GI) ia_dir_s2w(p, d, view_w);
    for (int i = 0; i + 1 < j; i++) {
        float s0 = tpl[i], e0 = tpl[i + 1];
        float mid = (s0 + e0) / 2.0f;
        float x[3] = {o[0] + d[0] * mid, o[1] + d[1] * mid, o[2] + d[2] * mid};
        float w;
        if (GI) {
            ia_team_query<true>(team, p, wgeo, x, q);
            cnt.qg++; cnt.fetch += q.n_fetch; cnt.geo += q.n_valid + (q.valid ? 1 : 0); cnt.skin += q.valid ? 1 : 0;
            float al = ia_alpha(q.sdf, e0 - s0, p.beta);
            w = Tacc * al;
            Tacc *= (1.0f - al);
            if (q.valid) {
                float nw[3], c[3];
                ia_dir_s2w(p, q.grad, nw);
                ia_team_radiance<false>(team, p, wmlp, q.xc, q.feat, view_w, nw, c, nullptr);
                cnt.rad++;
                rgb[0] += w * c[0]; rgb[1] += w * c[1]; rgb[2] += w * c[2];
            }
        } else {
            float sd = query_sdf(mid);
            (void)x;
            float al = ia_alpha(sd, e0 - s0, p.beta);
            w = Tacc * al;
            Tacc *= (1.0f - al);
        }
        acc += w;
    }
    T = 1.0f - acc;
}
