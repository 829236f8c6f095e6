// Training-mode building blocks of the physically based branch.
//
// In training the reference traces the secondary rays under torch.no_grad() -- directions, inverse pdf, transmittance and
// indirect radiance are constants of the graph (models/intrinsic_avatar.py:673-706 for uniform_light, the training default
// configs/config.yaml:46; :575-640, :763-800, :880-896 for the siblings) -- and differentiates only
//     diff, spec = MultiLobe.eval(wi, n, wo, roughness, albedo, metallic)         lib/torch_pbr/bxdf.py:111-146, 217-265, 321-330
//     em_li      = EnvironmentLightTensor.eval(dirs_world)                        lib/torch_pbr/light.py:298-339
//     Li         = em_li * tr + rgb_map
//     Lo_diff    = Li * diff * inv_pdf,  Lo_spec = Li * spec * inv_pdf            models/intrinsic_avatar.py:736-738
//     Lo         = (1 - metallic) * albedo * Lo_diff + Lo_spec                    :741-751
// with autograd.  k_pbr_shade is that combine for one shading sample per thread (the forward values come from the render
// path's own ia_brdf_multilobe, so they are the eval frame's bit for bit); k_pbr_shade_backward is its hand-derived backward
// to the normal, the three material channels and Li; k_env_backward scatters a gradient on em_li into the environment map's
// texels (the bilinear weights of ia_env_eval).
//
// The backward differentiates the frame-free form of the lobe: with a unit normal the local frame is orthonormal, so spec
// depends on n only through ci = wi.n, co = wo.n and ch = (wi+wo).n / |wi+wo| (|wi.wh| does not depend on n at all).  The
// gradient it returns for n is therefore the reference's up to a component along n, which the backward of the F.normalize
// that produced n (models/intrinsic_avatar.py:1093) removes.
#pragma once
#include "ia_pbr.cuh"

__device__ __forceinline__ void ia_load3(const float* __restrict__ p, long long i, float v[3]) {
    v[0] = p[i * 3]; v[1] = p[i * 3 + 1]; v[2] = p[i * 3 + 2];
}

__global__ void __launch_bounds__(256) k_pbr_shade(const float* __restrict__ wi_, const float* __restrict__ n_,
                                                   const float* __restrict__ wo_, const float* __restrict__ rough,
                                                   const float* __restrict__ albedo, const float* __restrict__ metal,
                                                   const float* __restrict__ Li_, const float* __restrict__ inv_pdf, long long n,
                                                   float* __restrict__ Lo, float* __restrict__ Lo_diff,
                                                   float* __restrict__ Lo_spec) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float wi[3], nn[3], wo[3], al[3], Li[3];
    ia_load3(wi_, i, wi); ia_load3(n_, i, nn); ia_load3(wo_, i, wo); ia_load3(albedo, i, al); ia_load3(Li_, i, Li);
    const float m = metal[i], ip = inv_pdf[i];
    float diff = 0.f, spec[3] = {0.f, 0.f, 0.f};
    if (ia_dot3(nn, wo) > 1e-6f) ia_brdf_multilobe(wi, nn, wo, rough[i], al, m, diff, spec);      // the cosine mask (:690)
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float ld = Li[c] * diff * ip, ls = Li[c] * spec[c] * ip;
        if (Lo_diff) Lo_diff[i * 3 + c] = ld;
        if (Lo_spec) Lo_spec[i * 3 + c] = ls;
        Lo[i * 3 + c] = (1.0f - m) * al[c] * ld + ls;
    }
}

__global__ void __launch_bounds__(256) k_pbr_shade_backward(
    const float* __restrict__ wi_, const float* __restrict__ n_, const float* __restrict__ wo_, const float* __restrict__ rough,
    const float* __restrict__ albedo, const float* __restrict__ metal, const float* __restrict__ Li_,
    const float* __restrict__ inv_pdf, const float* __restrict__ dLo_, const float* __restrict__ dLod_,
    const float* __restrict__ dLos_, long long n, float* __restrict__ g_n, float* __restrict__ g_rough,
    float* __restrict__ g_albedo, float* __restrict__ g_metal, float* __restrict__ g_Li) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float eps = 1e-6f, PI = 3.14159265358979323846f;
    float wi[3], nn[3], wo[3], al[3], Li[3], dLo[3], dLod[3] = {0.f, 0.f, 0.f}, dLos[3] = {0.f, 0.f, 0.f};
    ia_load3(wi_, i, wi); ia_load3(n_, i, nn); ia_load3(wo_, i, wo); ia_load3(albedo, i, al); ia_load3(Li_, i, Li);
    ia_load3(dLo_, i, dLo);
    if (dLod_) ia_load3(dLod_, i, dLod);
    if (dLos_) ia_load3(dLos_, i, dLos);
    const float a = rough[i], m = metal[i], ip = inv_pdf[i];
    const float ci = ia_dot3(wi, nn), co = ia_dot3(wo, nn);
    float gn[3] = {0.f, 0.f, 0.f}, ga = 0.f, gal[3] = {0.f, 0.f, 0.f}, gm = 0.f, gL[3] = {0.f, 0.f, 0.f};
    if (co > 1e-6f) {
        const float diff = co / PI;
        float Gd[3], Gs[3], d_diff = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float ld = Li[c] * diff * ip;
            Gd[c] = dLod[c] + (1.0f - m) * al[c] * dLo[c];
            Gs[c] = dLos[c] + dLo[c];
            gal[c] = (1.0f - m) * ld * dLo[c];
            gm -= al[c] * ld * dLo[c];
            gL[c] = Gd[c] * diff * ip;
            d_diff += Gd[c] * Li[c];
        }
        d_diff *= ip;
        float g_co = d_diff / PI, g_ci = 0.f;
        if (ci > eps) {                                   // the lobe's own support (co > eps holds)
            const float h[3] = {wi[0] + wo[0], wi[1] + wo[1], wi[2] + wo[2]};
            const float nh = fmaxf(sqrtf(ia_dot3(h, h)), 1e-12f);
            const float ch = ia_dot3(h, nn) / nh, cih = fabsf(ia_dot3(wi, h) / nh);
            const float k = (a * a + 2.f * a + 1.f) / 8.0f, a2 = a * a, cos2 = ch * ch;
            const float t = cos2 * (a2 - 1.f) + 1.f, D = PI * t * t + eps, ndf = a2 / D;
            const float den_i = ci * (1.0f - k) + k, den_o = co * (1.0f - k) + k;
            const bool oki = den_i > eps, oko = den_o > eps;
            const float g1i = oki ? ci / (den_i + eps) : 0.f, g1o = oko ? co / (den_o + eps) : 0.f;
            const float fr = exp2f((-5.55473f * cih - 6.98316f) * cih);
            const float common = ndf * g1i * g1o, den4 = 4.f * ci + eps;
            float sF = 0.f;                               // sum_c d_spec_c F_c
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float F0 = 0.04f * (1.0f - m) + al[c] * m, Fc = F0 + (1.0f - F0) * fr;
                const float d_spec = Gs[c] * Li[c] * ip;
                gL[c] += Gs[c] * (common * Fc / den4) * ip;
                sF += d_spec * Fc;
                const float d_F0 = d_spec * common / den4 * (1.0f - fr);
                gm += d_F0 * (al[c] - 0.04f);
                gal[c] += d_F0 * m;
            }
            const float d_common = sF / den4;
            g_ci -= sF * common * 4.f / (den4 * den4);
            const float d_ndf = d_common * g1i * g1o, d_g1i = d_common * ndf * g1o, d_g1o = d_common * ndf * g1i;
            ga += d_ndf * (2.f * a / D - a2 * (2.f * PI * t * (2.f * a * cos2)) / (D * D));
            const float g_ch = d_ndf * (-a2 * 2.f * PI * t * (2.f * ch * (a2 - 1.f)) / (D * D));
            float dk = 0.f;
            if (oki) {
                const float q = den_i + eps;
                g_ci += d_g1i * (1.0f / q - ci * (1.0f - k) / (q * q));
                dk -= d_g1i * ci * (1.0f - ci) / (q * q);
            }
            if (oko) {
                const float q = den_o + eps;
                g_co += d_g1o * (1.0f / q - co * (1.0f - k) / (q * q));
                dk -= d_g1o * co * (1.0f - co) / (q * q);
            }
            ga += dk * (a + 1.f) / 4.f;
#pragma unroll
            for (int c = 0; c < 3; c++) gn[c] += g_ch / nh * h[c];
        }
#pragma unroll
        for (int c = 0; c < 3; c++) gn[c] += g_ci * wi[c] + g_co * wo[c];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        g_n[i * 3 + c] = gn[c];
        g_albedo[i * 3 + c] = gal[c];
        g_Li[i * 3 + c] = gL[c];
    }
    g_rough[i] = ga;
    g_metal[i] = gm;
}

// Backward of ia_env_eval: g_env [H][W][3] += bilinear weights x d_em (float atomics; border clamp as the forward).
__global__ void __launch_bounds__(256) k_env_backward(const IaEnv E, const float* __restrict__ dirs, const float* __restrict__ d_em,
                                                      long long n, float* __restrict__ g_env) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int H = E.H, W = E.W;
    float d[3], g[3], u, v, la;
    ia_load3(dirs, i, d); ia_load3(d_em, i, g);
    if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f) return;
    ia_env_uv(d, u, v, la);
    const float fx = fminf(fmaxf(((u * 2 - 1) + 1.f) / 2 * (W - 1), 0.f), (float)(W - 1));
    const float fy = fminf(fmaxf(((v * 2 - 1) + 1.f) / 2 * (H - 1), 0.f), (float)(H - 1));
    const int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    const float wx = fx - x0, wy = fy - y0;
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float w00 = (1 - wx) * (1 - wy), w01 = wx * (1 - wy), w10 = (1 - wx) * wy, w11 = wx * wy;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        atomicAdd(g_env + ((size_t)y0 * W + x0) * 3 + ch, w00 * g[ch]);
        atomicAdd(g_env + ((size_t)y0 * W + x1) * 3 + ch, w01 * g[ch]);
        atomicAdd(g_env + ((size_t)y1 * W + x0) * 3 + ch, w10 * g[ch]);
        atomicAdd(g_env + ((size_t)y1 * W + x1) * 3 + ch, w11 * g[ch]);
    }
}

extern "C" int ia_op_pbr_shade(ia_ctx* c, const float* d_wi, const float* d_n, const float* d_wo, const float* d_rough,
                               const float* d_albedo, const float* d_metal, const float* d_Li, const float* d_inv_pdf, int64_t n,
                               float* d_Lo, float* d_Lo_diff, float* d_Lo_spec, void* stream) {
    IA_REQUIRE(c && n >= 0, IA_EINVAL, "ia_op_pbr_shade: bad argument");
    if (n == 0) return IA_OK;
    IA_REQUIRE(d_wi && d_n && d_wo && d_rough && d_albedo && d_metal && d_Li && d_inv_pdf && d_Lo, IA_EINVAL,
               "ia_op_pbr_shade: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    k_pbr_shade<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_wi, d_n, d_wo, d_rough, d_albedo, d_metal, d_Li,
                                                                             d_inv_pdf, n, d_Lo, d_Lo_diff, d_Lo_spec);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_pbr_shade_backward(ia_ctx* c, const float* d_wi, const float* d_n, const float* d_wo, const float* d_rough,
                                        const float* d_albedo, const float* d_metal, const float* d_Li, const float* d_inv_pdf,
                                        const float* d_dLo, const float* d_dLo_diff, const float* d_dLo_spec, int64_t n,
                                        float* d_g_n, float* d_g_rough, float* d_g_albedo, float* d_g_metal, float* d_g_Li,
                                        void* stream) {
    IA_REQUIRE(c && n >= 0, IA_EINVAL, "ia_op_pbr_shade_backward: bad argument");
    if (n == 0) return IA_OK;
    IA_REQUIRE(d_wi && d_n && d_wo && d_rough && d_albedo && d_metal && d_Li && d_inv_pdf && d_dLo && d_g_n && d_g_rough &&
                   d_g_albedo && d_g_metal && d_g_Li,
               IA_EINVAL, "ia_op_pbr_shade_backward: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    k_pbr_shade_backward<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_wi, d_n, d_wo, d_rough, d_albedo, d_metal, d_Li, d_inv_pdf, d_dLo, d_dLo_diff, d_dLo_spec, n, d_g_n, d_g_rough,
        d_g_albedo, d_g_metal, d_g_Li);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_env_backward(ia_ctx* c, const float* d_dirs_world, const float* d_dem, int64_t n, float* d_g_env,
                                  void* stream) {
    IA_REQUIRE(c && n >= 0 && d_g_env, IA_EINVAL, "ia_op_env_backward: bad argument");
    IA_REQUIRE(c->have_light, IA_ESTATE, "ia_op_env_backward: call ia_set_light first");
    if (n == 0) return IA_OK;
    IA_REQUIRE(d_dirs_world && d_dem, IA_EINVAL, "ia_op_env_backward: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    k_env_backward<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c->env, d_dirs_world, d_dem, n, d_g_env);
    IA_LAUNCH_CHECK();
    return IA_OK;
}
