// libia_b200: kernels + C ABI of the B200-native IntrinsicAvatar render path (sm_100a).
// See include/ia_b200.h for the ABI and the reference interfaces each entry point replaces.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>

#include "../../include/ia_b200.h"
#include "ia_pbr.cuh"

// ================================================================================================
// error handling
static thread_local std::string g_err;
extern "C" const char* ia_last_error(void) { return g_err.c_str(); }
extern "C" int ia_version(void) { return 200; }
// 1: voxel_J is stored as 32-byte voxels (IA_VOXEL32: Broyden roots agree with the reference kernel to ~1e-5);
// 0: 48-byte fp32 voxels (bit-compatible arithmetic, roots within 2e-6)
extern "C" int ia_voxel_format(void) { return IA_VOXEL32; }

#define IA_CHECK_CUDA(expr)                                                                     \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            g_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                         \
            return IA_ECUDA;                                                                    \
        }                                                                                       \
    } while (0)
#define IA_REQUIRE(cond, code, msg)                                                             \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            g_err = msg;                                                                        \
            return code;                                                                        \
        }                                                                                       \
    } while (0)
#define IA_LAUNCH_CHECK() IA_CHECK_CUDA(cudaGetLastError())
// stage timing: events are recorded only when enabled with ia_set_timing
#define IA_STAGE_BEGIN(c, id, st)                                   \
    do {                                                            \
        if ((c)->timing) { cudaEventRecord((c)->ev0[id], st); }     \
    } while (0)
#define IA_STAGE_END(c, id, st, n_kernels)                          \
    do {                                                            \
        (c)->n_launches += (n_kernels);                             \
        if ((c)->timing) { cudaEventRecord((c)->ev1[id], st); (c)->ev_used[id] = true; } \
    } while (0)

// ================================================================================================
// context
struct ia_ctx {
    int device = 0;
    int n_sm = 148;
    IaFrame f;  // host copy of the per-frame constants
    bool have_fields = false, have_lbs = false, have_pose = false, have_occ = false, have_light = false,
         have_cfg = false;
    // owned device memory
    float* d_mlp = nullptr;
    float4* d_lbs_w = nullptr;
    float4* d_voxel_J = nullptr;
    uint32_t* d_occ_bits = nullptr;
    int occ_res_alloc = 0;
    float occ_thre = 0.001f;
    // light
    int spp = 0;
    int env_H = 0, env_W = 0;
    float *d_light_dir_w = nullptr, *d_light_dir_s = nullptr, *d_light_em = nullptr, *d_light_pdf = nullptr;
    float* d_u_table = nullptr;
    float *d_env_pdf = nullptr, *d_env_cols = nullptr, *d_env_rows = nullptr, *d_env_rowsum = nullptr;
    double* d_env_total = nullptr;
    IaEnv env = {};              // tables of the last ia_set_light* call (env itself is the caller's buffer)
    bool light_uniform = false;  // light tables hold the stratified sphere of render_mode = uniform_light
    float* d_vis = nullptr;      // [n_rays] visibility accumulator (uniform_light)
    float* d_bg = nullptr;       // [n_rays][3] background radiance per ray (background colour / add_emitter)
    // workspace for ia_render
    int64_t ws_rays = 0, ws_samples = 0, ws_resamples = 0;
    int* d_hit_rays = nullptr;       // [n_rays] ray index of every hit slot, in ray order
    uint8_t* d_hit_flag = nullptr;   // [n_rays] ray enters an occupied cell
    int* d_blk_cnt = nullptr;        // [ceil(n_rays / 256)] hits per setup block -> exclusive offsets
    float* d_hit_od = nullptr;       // [n_rays][8]: o(3), d(3), far, opacity
    int* d_hit_info = nullptr;       // [n_rays][2]: sample offset, count
    IaSample* d_samples = nullptr;   // [ws_samples]
    struct IaSampleAux* d_samples_aux = nullptr;  // [ws_samples] rgb, world normal, ray slot
    float* d_rs_t = nullptr;         // [ws_resamples]
    float* d_rs_w = nullptr;
    int* d_rs_src = nullptr;
    float* d_acc = nullptr;          // [n_rays][6] rgb_phys, demod_phys accumulators
    unsigned long long* d_counters = nullptr;  // IA_N_COUNTERS
    int* d_work = nullptr;           // [8] work-stealing counters / n_hit / n_samples
    unsigned char* d_wf_scratch = nullptr;  // wavefront integrator: per-CTA candidate / task scratch
    // stage timing (CUDA events on the launching stream) and launch accounting
    bool timing = false;
    cudaEvent_t ev0[IA_N_STAGES] = {}, ev1[IA_N_STAGES] = {};
    bool ev_used[IA_N_STAGES] = {};
    unsigned long long n_launches = 0;
    // occupancy scratch
    float *d_occ_a = nullptr, *d_occ_b = nullptr;
    int* d_occ_hist = nullptr;
    double* d_occ_sum = nullptr;
};

template <typename T>
static int ia_realloc(T** p, size_t n) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    if (n == 0) return IA_OK;
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) {
        g_err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return IA_ECUDA;
    }
    return IA_OK;
}

extern "C" int ia_create(ia_ctx** out, int device) {
    IA_REQUIRE(out, IA_EINVAL, "ia_create: out is NULL");
    IA_CHECK_CUDA(cudaSetDevice(device));
    ia_ctx* c = new ia_ctx();
    c->device = device;
    memset(&c->f, 0, sizeof(IaFrame));
    cudaDeviceProp prop;
    IA_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    static const int bones[IA_N_INIT] = {0, 1, 2, 4, 5, 10, 11, 12, 15, 16, 17, 18, 19};  // deformer_torch.py:27
    for (int i = 0; i < IA_N_INIT; i++) c->f.init_bones[i] = bones[i];
    for (int i = 0; i < 3; i++) { c->f.albedo_ratio[i] = 1.f; c->f.background[i] = 1.f; }
    if (ia_realloc(&c->d_counters, IA_N_COUNTERS) || ia_realloc(&c->d_work, 8)) return IA_ECUDA;
    IA_CHECK_CUDA(cudaMemset(c->d_counters, 0, IA_N_COUNTERS * sizeof(unsigned long long)));
    IA_CHECK_CUDA(cudaMemset(c->d_work, 0, 8 * sizeof(int)));
    *out = c;
    return IA_OK;
}

extern "C" int ia_set_timing(ia_ctx* c, int enable) {
    IA_REQUIRE(c, IA_EINVAL, "ia_set_timing: NULL context");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    if (enable && !c->ev0[0]) {
        for (int i = 0; i < IA_N_STAGES; i++) {
            IA_CHECK_CUDA(cudaEventCreate(&c->ev0[i]));
            IA_CHECK_CUDA(cudaEventCreate(&c->ev1[i]));
        }
    }
    c->timing = enable != 0;
    for (int i = 0; i < IA_N_STAGES; i++) c->ev_used[i] = false;
    return IA_OK;
}

extern "C" int ia_get_timings(ia_ctx* c, float* h_ms, uint64_t* h_launches, void* stream) {
    IA_REQUIRE(c && h_ms, IA_EINVAL, "ia_get_timings: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    IA_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    for (int i = 0; i < IA_N_STAGES; i++) {
        h_ms[i] = -1.f;
        if (c->timing && c->ev_used[i]) IA_CHECK_CUDA(cudaEventElapsedTime(&h_ms[i], c->ev0[i], c->ev1[i]));
    }
    if (h_launches) *h_launches = c->n_launches;
    return IA_OK;
}

extern "C" int ia_destroy(ia_ctx* c) {
    if (!c) return IA_OK;
    cudaSetDevice(c->device);
    for (int i = 0; i < IA_N_STAGES; i++) {
        if (c->ev0[i]) cudaEventDestroy(c->ev0[i]);
        if (c->ev1[i]) cudaEventDestroy(c->ev1[i]);
    }
    void* ptrs[] = {c->d_mlp, c->d_lbs_w, c->d_voxel_J, c->d_occ_bits, c->d_light_dir_w, c->d_light_dir_s,
                    c->d_light_em, c->d_light_pdf, c->d_u_table, c->d_env_pdf, c->d_env_cols, c->d_env_rows,
                    c->d_env_rowsum, c->d_env_total, c->d_hit_rays, c->d_hit_od, c->d_hit_info, c->d_samples,
                    c->d_rs_t, c->d_rs_w, c->d_rs_src, c->d_acc, c->d_counters, c->d_work, c->d_occ_a, c->d_occ_b,
                    c->d_occ_hist, c->d_occ_sum, c->d_wf_scratch, c->d_samples_aux, c->d_vis, c->d_bg, c->d_hit_flag, c->d_blk_cnt};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete c;
    return IA_OK;
}

// ================================================================================================
// model state
extern "C" int ia_set_fields(ia_ctx* c, const float* d_geo_hash, const float* d_rad_hash, int64_t n_entries,
                             const float* lvl_scale, const int32_t* lvl_res, const int32_t* lvl_size,
                             const int32_t* lvl_off, const float* geo_w1, const float* geo_b1, const float* geo_w2,
                             const float* geo_b2, const float* rad_w1, const float* rad_b1, const float* rad_w2,
                             const float* rad_b2, const float* rad_w3, const float* rad_b3, const float* mat_w1,
                             const float* mat_b1, const float* mat_w2, const float* mat_b2, const float* mat_w3,
                             const float* mat_b3, const float* mat_scale, const float* mat_bias, const float* bbox,
                             float beta, void* stream) {
    IA_REQUIRE(c && d_geo_hash && d_rad_hash, IA_EINVAL, "ia_set_fields: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    int64_t need = 0;
    for (int l = 0; l < IA_N_LEVELS; l++) {
        c->f.lvl_scale[l] = lvl_scale[l];
        c->f.lvl_res[l] = (uint32_t)lvl_res[l];
        c->f.lvl_size[l] = (uint32_t)lvl_size[l];
        c->f.lvl_off[l] = (uint32_t)lvl_off[l];
        need = std::max<int64_t>(need, (int64_t)lvl_off[l] + lvl_size[l]);
    }
    IA_REQUIRE(need <= n_entries, IA_EINVAL, "ia_set_fields: hash table smaller than the level layout");
    c->f.geo_hash = reinterpret_cast<const float2*>(d_geo_hash);
    c->f.rad_hash = reinterpret_cast<const float2*>(d_rad_hash);
    std::vector<float> blob(IA_MLP_END, 0.f);
    auto put_T = [&](int off, const float* w, int n_out, int n_in) {  // [out][in] -> [in][out]
        for (int o = 0; o < n_out; o++)
            for (int i = 0; i < n_in; i++) blob[off + i * n_out + o] = w[o * n_in + i];
    };
    auto put = [&](int off, const float* w, int n) { memcpy(&blob[off], w, n * sizeof(float)); };
    put_T(IA_GEO_W1T, geo_w1, 64, 35); put(IA_GEO_B1, geo_b1, 64); put(IA_GEO_W2, geo_w2, 13 * 64); put(IA_GEO_B2, geo_b2, 13);
    put_T(IA_RAD_W1T, rad_w1, 64, 67); put(IA_RAD_B1, rad_b1, 64); put_T(IA_RAD_W2T, rad_w2, 64, 64);
    put(IA_RAD_B2, rad_b2, 64); put(IA_RAD_W3, rad_w3, 3 * 64); put(IA_RAD_B3, rad_b3, 3);
    put_T(IA_MAT_W1T, mat_w1, 64, 48); put(IA_MAT_B1, mat_b1, 64); put_T(IA_MAT_W2T, mat_w2, 64, 64);
    put(IA_MAT_B2, mat_b2, 64); put(IA_MAT_W3, mat_w3, 5 * 64); put(IA_MAT_B3, mat_b3, 5);
    // mma.sync B fragments (layout: ia_mma.cuh): frag(f0, k_steps, n_tiles, W) with W(k, n) = weight between input slot k and
    // output n of the layer; layers fed from C fragments take their inputs in ia_kperm order
    blob.resize(IA_BLOB_FLOATS, 0.f);
    auto rna_tf32 = [](float x) {   // cvt.rna.tf32.f32: round the magnitude to 10 mantissa bits, ties away from zero
        uint32_t u;
        memcpy(&u, &x, 4);
        u = (u + 0x1000u) & 0xffffe000u;
        float r;
        memcpy(&r, &u, 4);
        return r;
    };
    auto frag = [&](int f0, int k_steps, int n_tiles, auto W) {
        for (int s = 0; s < k_steps; s++)
            for (int nt = 0; nt < n_tiles; nt++)
                for (int lane = 0; lane < 32; lane++) {
                    const int g = lane >> 2, t = lane & 3;
                    float* d = &blob[IA_MLP_END + ((size_t)(f0 + s * n_tiles + nt) * 32 + lane) * 4];
                    const float b0 = W(8 * s + t, 8 * nt + g), b1 = W(8 * s + t + 4, 8 * nt + g);
                    d[0] = rna_tf32(b0); d[1] = rna_tf32(b0 - d[0]);
                    d[2] = rna_tf32(b1); d[3] = rna_tf32(b1 - d[2]);
                }
    };
    const float* B = blob.data();
    auto geo_in = [](int k) { return k < 32 ? 3 + k : (k < 35 ? k - 32 : -1); };   // tile column -> geometry input
    frag(IA_FRAG_FEAT, 8, 2, [&](int k, int n) { return n < 13 ? B[IA_GEO_W2 + n * 64 + ia_kperm(k)] : 0.f; });
    frag(IA_FRAG_BWD, 8, 5, [&](int k, int n) { return geo_in(n) < 0 ? 0.f : B[IA_GEO_W1T + geo_in(n) * 64 + ia_kperm(k)]; });
    frag(IA_FRAG_RAD1, 9, 8, [&](int k, int n) { return ia_rad_in_of(k) < 0 ? 0.f : B[IA_RAD_W1T + ia_rad_in_of(k) * 64 + n]; });
    frag(IA_FRAG_RAD2, 8, 8, [&](int k, int n) { return B[IA_RAD_W2T + ia_kperm(k) * 64 + n]; });
    frag(IA_FRAG_RAD3, 8, 1, [&](int k, int n) { return n < 3 ? B[IA_RAD_W3 + n * 64 + ia_kperm(k)] : 0.f; });
    frag(IA_FRAG_MAT1, 6, 8, [&](int k, int n) { return ia_mat_in_of(k) < 0 ? 0.f : B[IA_MAT_W1T + ia_mat_in_of(k) * 64 + n]; });
    frag(IA_FRAG_MAT2, 8, 8, [&](int k, int n) { return B[IA_MAT_W2T + ia_kperm(k) * 64 + n]; });
    frag(IA_FRAG_MAT3, 8, 1, [&](int k, int n) { return n < 5 ? B[IA_MAT_W3 + n * 64 + ia_kperm(k)] : 0.f; });
    frag(IA_FRAG_GEO1, 5, 8, [&](int k, int n) { return geo_in(k) < 0 ? 0.f : B[IA_GEO_W1T + geo_in(k) * 64 + n]; });
    if (ia_realloc(&c->d_mlp, IA_BLOB_FLOATS)) return IA_ECUDA;
    IA_CHECK_CUDA(cudaMemcpyAsync(c->d_mlp, blob.data(), (size_t)IA_BLOB_FLOATS * sizeof(float), cudaMemcpyHostToDevice,
                                  (cudaStream_t)stream));
    IA_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // blob is a stack-scoped host buffer
    c->f.mlp = c->d_mlp;
    for (int i = 0; i < 5; i++) { c->f.mat_scale[i] = mat_scale[i]; c->f.mat_bias[i] = mat_bias[i]; }
    for (int i = 0; i < 3; i++) {
        c->f.center[i] = (bbox[i] + bbox[3 + i]) / 2;
        c->f.scale[i] = bbox[3 + i] - bbox[i];
    }
    c->f.beta = beta;
    c->have_fields = true;
    return IA_OK;
}

__global__ void k_repack_lbs(const float* __restrict__ src, float* __restrict__ dst, int nvox) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
#pragma unroll
    for (int j = 0; j < IA_N_BONES; j++) dst[(size_t)v * IA_N_BONES + j] = src[(size_t)j * nvox + v];
}

extern "C" int ia_set_lbs_voxels(ia_ctx* c, const float* d_lbs_voxel, int D, int H, int W, const float* off,
                                 const float* scl, void* stream) {
    IA_REQUIRE(c && d_lbs_voxel && off && scl, IA_EINVAL, "ia_set_lbs_voxels: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    size_t nvox = (size_t)D * H * W;
    if (ia_realloc(&c->d_lbs_w, nvox * 6) || ia_realloc(&c->d_voxel_J, nvox * IA_VOXEL_F4)) return IA_ECUDA;
    k_repack_lbs<<<(unsigned)((nvox + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_lbs_voxel, (float*)c->d_lbs_w,
                                                                                 (int)nvox);
    IA_LAUNCH_CHECK();
    c->f.D = D; c->f.H = H; c->f.W = W;
    for (int i = 0; i < 3; i++) { c->f.off[i] = off[i]; c->f.scl[i] = scl[i]; }
    c->f.lbs_w = c->d_lbs_w;
    c->f.voxel_J = c->d_voxel_J;
    {
        // voxel (xi, yi, zi) sits at g = 2 i / (n - 1) - 1 of the normalised cube, i.e. at x = g / scl - off
        const int n[3] = {W, H, D};
        for (int i = 0; i < 3; i++) {
            c->f.vox_h[i] = (float)(2.0 / (double)(n[i] - 1) / (double)scl[i]);
            c->f.vox_b[i] = (float)(-1.0 / (double)scl[i] - (double)off[i]);
        }
    }
    c->have_lbs = true;
    return IA_OK;
}

// precompute_kernel (precompute.cu:22-71): blended 3x4 per voxel from the channels-last skinning weights.
__device__ __forceinline__ void ia_blend_J(const IaFrame& p, int v, float J[12]) {
    float w[IA_N_BONES];
    const float4* src = p.lbs_w + (size_t)v * 6;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        float4 a = __ldg(src + q);
        w[q * 4 + 0] = a.x; w[q * 4 + 1] = a.y; w[q * 4 + 2] = a.z; w[q * 4 + 3] = a.w;
    }
#pragma unroll
    for (int k = 0; k < 12; k++) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < IA_N_BONES; j++) s += w[j] * p.tfs[j][k];
        J[k] = s;
    }
}

__global__ void k_precompute(const __grid_constant__ IaFrame p, float4* __restrict__ voxel_J, int nvox) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    float J[12];
    ia_blend_J(p, v, J);
    float4* o = voxel_J + (size_t)v * IA_VOXEL_F4;
#if IA_VOXEL32
    {
        // deformed voxel centre in fp32 (evaluated in double, with the same fp32 lattice constants the fetch uses),
        // rotation in fp16
        const int xi = v % p.W, yi = (v / p.W) % p.H, zi = v / (p.W * p.H);
        const double cx = (double)fmaf((float)xi, p.vox_h[0], p.vox_b[0]);
        const double cy = (double)fmaf((float)yi, p.vox_h[1], p.vox_b[1]);
        const double cz = (double)fmaf((float)zi, p.vox_h[2], p.vox_b[2]);
        __half2 h[5] = {__floats2half2_rn(J[0], J[1]), __floats2half2_rn(J[2], J[4]), __floats2half2_rn(J[5], J[6]),
                        __floats2half2_rn(J[8], J[9]), __floats2half2_rn(J[10], 0.f)};
        // y_c is formed with the UNROUNDED rotation: the stored pair then gives y_c + R~ (x - c_c) =
        // J x + t + (R~ - R)(x - c_c), i.e. the fp16 rounding of R only acts on the offset from the voxel centre
        const float y0 = (float)((double)J[0] * cx + (double)J[1] * cy + (double)J[2] * cz + (double)J[3]);
        const float y1 = (float)((double)J[4] * cx + (double)J[5] * cy + (double)J[6] * cz + (double)J[7]);
        const float y2 = (float)((double)J[8] * cx + (double)J[9] * cy + (double)J[10] * cz + (double)J[11]);
        const float* hf = reinterpret_cast<const float*>(h);
        o[0] = make_float4(y0, y1, y2, hf[0]);
        o[1] = make_float4(hf[1], hf[2], hf[3], hf[4]);
    }
#else
    o[0] = make_float4(J[0], J[1], J[2], J[3]);
    o[1] = make_float4(J[4], J[5], J[6], J[7]);
    o[2] = make_float4(J[8], J[9], J[10], J[11]);
#endif
}

extern "C" int ia_set_pose(ia_ctx* c, const float* tfs, const float* w2s, void* stream) {
    IA_REQUIRE(c && tfs && w2s, IA_EINVAL, "ia_set_pose: NULL argument");
    IA_REQUIRE(c->have_lbs, IA_ESTATE, "ia_set_pose: call ia_set_lbs_voxels first");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    for (int j = 0; j < IA_N_BONES; j++)
        for (int k = 0; k < 12; k++) c->f.tfs[j][k] = tfs[j * 16 + k];
    for (int k = 0; k < 12; k++) c->f.w2s[k] = w2s[k];
    int nvox = c->f.D * c->f.H * c->f.W;
    IA_STAGE_BEGIN(c, IA_STAGE_PRECOMPUTE, (cudaStream_t)stream);
    k_precompute<<<(nvox + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c->f, c->d_voxel_J, nvox);
    IA_STAGE_END(c, IA_STAGE_PRECOMPUTE, (cudaStream_t)stream, 1);
    IA_LAUNCH_CHECK();
    c->have_pose = true;
    return IA_OK;
}

extern "C" int ia_set_render_config(ia_ctx* c, const float* aabb, int n_per_ray, int n_per_sec, float sec_near,
                                    float sec_far, float occ_thre, const float* bg, const float* ratio) {
    IA_REQUIRE(c && aabb, IA_EINVAL, "ia_set_render_config: NULL argument");
    IA_REQUIRE(n_per_ray > 0 && n_per_sec > 1, IA_EINVAL, "ia_set_render_config: bad sample counts");
    float dx = aabb[3] - aabb[0], dy = aabb[4] - aabb[1], dz = aabb[5] - aabb[2];
    c->f.step_primary = sqrtf(dx * dx + dy * dy + dz * dz) / (float)n_per_ray;  // models/intrinsic_avatar.py:197-211
    c->f.sec_near = sec_near;
    c->f.sec_far = sec_far;
    c->f.sec_step = (sec_far - sec_near) / (float)(n_per_sec - 1);              // :458-460
    c->occ_thre = occ_thre;
    for (int i = 0; i < 3; i++) {
        c->f.background[i] = bg ? bg[i] : 1.f;
        c->f.albedo_ratio[i] = ratio ? ratio[i] : 1.f;
    }
    c->have_cfg = true;
    return IA_OK;
}

extern "C" int ia_set_secondary_sampling(ia_ctx* c, int importance_sample, int zero_crossing_search) {
    IA_REQUIRE(c, IA_EINVAL, "ia_set_secondary_sampling: NULL context");
    c->f.sec_mode = !importance_sample ? IA_SEC_PLAIN : (zero_crossing_search ? IA_SEC_ZERO_CROSSING : IA_SEC_IMPORTANCE);
    return IA_OK;
}

// ================================================================================================
// op-level kernels
// the blended fp32 transform in the reference's channel-major layout [12][nvox], recomputed from the context's skinning
// weights and bone transforms (independent of how voxel_J is stored)
__global__ void k_op_precompute_out(const __grid_constant__ IaFrame p, float* __restrict__ out, int nvox) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    float J[12];
    ia_blend_J(p, v, J);
#pragma unroll
    for (int k = 0; k < 12; k++) out[(size_t)k * nvox + v] = J[k];
}

extern "C" int ia_op_precompute(ia_ctx* c, float* d_out, void* stream) {
    IA_REQUIRE(c && d_out, IA_EINVAL, "ia_op_precompute: NULL argument");
    IA_REQUIRE(c->have_pose, IA_ESTATE, "ia_op_precompute: call ia_set_pose first");
    int nvox = c->f.D * c->f.H * c->f.W;
    k_op_precompute_out<<<(nvox + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c->f, d_out, nvox);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

__global__ void __launch_bounds__(256) k_op_broyden(const __grid_constant__ IaFrame p, const float* __restrict__ xd,
                                                    long long n, float* __restrict__ x_out, float* __restrict__ Jinv_out,
                                                    uint8_t* __restrict__ valid_raw, uint8_t* __restrict__ valid_out) {
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int lane = team.thread_rank();
    long long teams = (long long)gridDim.x * (blockDim.x / IA_TEAM);
    for (long long i = (long long)blockIdx.x * (blockDim.x / IA_TEAM) + threadIdx.x / IA_TEAM; i < n; i += teams) {
        float pt[3] = {xd[i * 3 + 0], xd[i * 3 + 1], xd[i * 3 + 2]};
        float x[3] = {0, 0, 0}, Ji[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool ok = false;
        if (lane < IA_N_INIT) ok = ia_broyden_chain(p, p.init_bones[lane], pt, x, Ji, nullptr);
        bool keep = ok;
#pragma unroll
        for (int j = 1; j < IA_N_INIT; j++) {
            float xj0 = team.shfl(x[0], j), xj1 = team.shfl(x[1], j), xj2 = team.shfl(x[2], j);
            bool vj = team.shfl((int)ok, j) != 0;
            float e0 = x[0] - xj0, e1 = x[1] - xj1, e2 = x[2] - xj2;
            if (vj && j > lane && (e0 * e0 + e1 * e1 + e2 * e2) < 0.0001f * 0.0001f) keep = false;
        }
        if (lane < IA_N_INIT) {
            size_t o = (size_t)i * IA_N_INIT + lane;
            x_out[o * 3 + 0] = x[0]; x_out[o * 3 + 1] = x[1]; x_out[o * 3 + 2] = x[2];
            if (Jinv_out)
                for (int k = 0; k < 9; k++) Jinv_out[o * 9 + k] = ok ? Ji[k] : 0.f;
            if (valid_raw) valid_raw[o] = ok;
            valid_out[o] = keep;
        }
    }
}

extern "C" int ia_op_broyden(ia_ctx* c, const float* d_xd, int64_t n, float* d_x, float* d_Jinv, uint8_t* d_valid_raw,
                             uint8_t* d_valid, void* stream) {
    IA_REQUIRE(c && d_xd && d_x && d_valid, IA_EINVAL, "ia_op_broyden: NULL argument");
    IA_REQUIRE(c->have_pose, IA_ESTATE, "ia_op_broyden: call ia_set_pose first");
    if (n == 0) return IA_OK;
    int blocks = (int)std::min<int64_t>((n + 15) / 16, (int64_t)c->n_sm * 16);
    k_op_broyden<<<blocks, 256, 0, (cudaStream_t)stream>>>(c->f, d_xd, n, d_x, d_Jinv, d_valid_raw, d_valid);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// ------------------------------------------------------------------------------------------------
// stage a weight range of the blob in shared memory
__device__ __forceinline__ void ia_stage(float* dst, const float* __restrict__ src, int n) {
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
}

template <bool GRAD>
__global__ void __launch_bounds__(256) k_op_query(const __grid_constant__ IaFrame p, const float* __restrict__ xd,
                                                  long long n, float* __restrict__ sdf, float* __restrict__ xc,
                                                  uint8_t* __restrict__ valid, float* __restrict__ grad,
                                                  float* __restrict__ grad_cano, float* __restrict__ feat,
                                                  unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) float smem[];
    ia_stage(smem, p.mlp, IA_GEO_END);
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int lane = team.thread_rank();
    long long teams = (long long)gridDim.x * (blockDim.x / IA_TEAM);
    unsigned nq = 0, nfetch = 0, ngeo = 0;
    for (long long i = (long long)blockIdx.x * (blockDim.x / IA_TEAM) + threadIdx.x / IA_TEAM; i < n; i += teams) {
        float pt[3] = {xd[i * 3 + 0], xd[i * 3 + 1], xd[i * 3 + 2]};
        IaQuery q;
        ia_team_query<GRAD>(team, p, smem, pt, q);
        nq++; nfetch += q.n_fetch; ngeo += (lane == 0) ? q.n_valid : 0;
        if (lane == 0) {
            if (sdf) sdf[i] = q.sdf;
            if (valid) valid[i] = q.valid;
            if (xc) { xc[i * 3 + 0] = q.xc[0]; xc[i * 3 + 1] = q.xc[1]; xc[i * 3 + 2] = q.xc[2]; }
            if (GRAD) {
                if (grad) { grad[i * 3 + 0] = q.grad[0]; grad[i * 3 + 1] = q.grad[1]; grad[i * 3 + 2] = q.grad[2]; }
                if (grad_cano) {
                    grad_cano[i * 3 + 0] = q.grad_cano[0]; grad_cano[i * 3 + 1] = q.grad_cano[1];
                    grad_cano[i * 3 + 2] = q.grad_cano[2];
                }
                if (feat)
                    for (int o = 0; o < 13; o++) feat[i * 13 + o] = q.feat[o];
            }
        }
    }
    if (lane == 0 && nq) atomicAdd(&counters[GRAD ? IA_CNT_QUERIES_GRAD : IA_CNT_QUERIES], nq);
    if (nfetch) atomicAdd(&counters[IA_CNT_BROYDEN_FETCH], nfetch);
    if (ngeo) atomicAdd(&counters[IA_CNT_GEO_EVAL], ngeo);
}

static int ia_query_blocks(const ia_ctx* c, int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + 15) / 16, (int64_t)c->n_sm * 8));
}

extern "C" int ia_op_query(ia_ctx* c, const float* d_xd, int64_t n, int with_grad, float* d_sdf, float* d_xc,
                           uint8_t* d_valid, float* d_grad, float* d_grad_cano, float* d_feature, void* stream) {
    IA_REQUIRE(c && d_xd, IA_EINVAL, "ia_op_query: NULL argument");
    IA_REQUIRE(c->have_pose && c->have_fields, IA_ESTATE, "ia_op_query: fields and pose must be set");
    if (n == 0) return IA_OK;
    size_t sm = IA_GEO_END * sizeof(float);
    if (with_grad)
        k_op_query<true><<<ia_query_blocks(c, n), 256, sm, (cudaStream_t)stream>>>(
            c->f, d_xd, n, d_sdf, d_xc, d_valid, d_grad, d_grad_cano, d_feature, c->d_counters);
    else
        k_op_query<false><<<ia_query_blocks(c, n), 256, sm, (cudaStream_t)stream>>>(
            c->f, d_xd, n, d_sdf, d_xc, d_valid, nullptr, nullptr, nullptr, c->d_counters);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

__global__ void __launch_bounds__(256) k_op_shade_fields(const __grid_constant__ IaFrame p, const float* __restrict__ xc,
                                                         const float* __restrict__ feat, const float* __restrict__ view,
                                                         const float* __restrict__ nrm, long long n,
                                                         float* __restrict__ rgb, float* __restrict__ mat) {
    extern __shared__ __align__(16) float smem[];
    ia_stage(smem, p.mlp, IA_MLP_END);
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    long long teams = (long long)gridDim.x * (blockDim.x / IA_TEAM);
    for (long long i = (long long)blockIdx.x * (blockDim.x / IA_TEAM) + threadIdx.x / IA_TEAM; i < n; i += teams) {
        float x[3], f[13], v[3], nw[3], c[3], m[5];
        for (int d = 0; d < 3; d++) { x[d] = xc[i * 3 + d]; v[d] = view[i * 3 + d]; nw[d] = nrm[i * 3 + d]; }
        for (int o = 0; o < 13; o++) f[o] = feat[i * 13 + o];
        ia_team_radiance<true>(team, p, smem, x, f, v, nw, c, m);
        if (team.thread_rank() == 0) {
            for (int d = 0; d < 3; d++) rgb[i * 3 + d] = c[d];
            for (int o = 0; o < 5; o++) mat[i * 5 + o] = m[o];
        }
    }
}

extern "C" int ia_op_shade_fields(ia_ctx* c, const float* d_xc, const float* d_feature, const float* d_view,
                                  const float* d_normal, int64_t n, float* d_rgb, float* d_mat, void* stream) {
    IA_REQUIRE(c && d_xc && d_feature && d_view && d_normal && d_rgb && d_mat, IA_EINVAL, "ia_op_shade_fields: NULL");
    IA_REQUIRE(c->have_fields, IA_ESTATE, "ia_op_shade_fields: fields must be set");
    if (n == 0) return IA_OK;
    size_t sm = IA_MLP_END * sizeof(float);
    IA_CHECK_CUDA(cudaFuncSetAttribute(k_op_shade_fields, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_op_shade_fields<<<ia_query_blocks(c, n), 256, sm, (cudaStream_t)stream>>>(c->f, d_xc, d_feature, d_view, d_normal,
                                                                                n, d_rgb, d_mat);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_op_traverse(const __grid_constant__ IaFrame p, const float* __restrict__ ro,
                              const float* __restrict__ rd, long long n, float near_plane, float far_plane, float step,
                              int* __restrict__ n_edges, int* __restrict__ n_samples, const int* __restrict__ edge_base,
                              const int* __restrict__ sample_base, float* __restrict__ vals, uint8_t* __restrict__ is_left,
                              uint8_t* __restrict__ is_right, float* __restrict__ t_starts, float* __restrict__ t_ends) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float o[3] = {ro[i * 3], ro[i * 3 + 1], ro[i * 3 + 2]}, d[3] = {rd[i * 3], rd[i * 3 + 1], rd[i * 3 + 2]};
    IaMarcher m;
    m.init(p, o, d, near_plane, far_plane, step);
    int ne = 0, ns = 0;
    float ts, te;
    bool cont;
    const bool write = vals != nullptr;
    int eb = write ? edge_base[i] : 0, sb = write ? sample_base[i] : 0;
    while (m.next(p.occ_bits, p.occ_res, ts, te, cont)) {
        if (!cont) {
            if (write) { vals[eb + ne] = ts; is_left[eb + ne] = 1; }
            ne++;
            if (write) { vals[eb + ne] = te; is_right[eb + ne] = 1; }
            ne++;
        } else {
            if (write) { vals[eb + ne] = te; is_left[eb + ne - 1] = 1; is_right[eb + ne] = 1; }
            ne++;
        }
        if (write) { t_starts[sb + ns] = ts; t_ends[sb + ns] = te; }
        ns++;
    }
    if (!write) { n_edges[i] = ne; n_samples[i] = ns; }
}

extern "C" int ia_op_traverse(ia_ctx* c, const float* d_o, const float* d_d, int64_t n, float near_plane, float far_plane,
                              float step, int32_t* d_ne, int32_t* d_ns, const int32_t* d_eb, const int32_t* d_sb,
                              float* d_vals, uint8_t* d_il, uint8_t* d_ir, float* d_ts, float* d_te, void* stream) {
    IA_REQUIRE(c && d_o && d_d, IA_EINVAL, "ia_op_traverse: NULL argument");
    IA_REQUIRE(c->have_occ, IA_ESTATE, "ia_op_traverse: occupancy grid not set");
    if (n == 0) return IA_OK;
    k_op_traverse<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(c->f, d_o, d_d, n, near_plane, far_plane,
                                                                                step, d_ne, d_ns, d_eb, d_sb, d_vals, d_il,
                                                                                d_ir, d_ts, d_te);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// ================================================================================================
// occupancy grid build (models/intrinsic_avatar.py:307-381)
// N_S jittered points per cell (3: test-time grid, max over them; 1: the training-time update).  ema != NULL: the EMA state of
// the level, occs = max(ema * decay, alpha), written back (OccGridEstimator._update, temporal_occ_grid.py:392-394).
template <int N_S>
__global__ void __launch_bounds__(256) k_occ_eval(const __grid_constant__ IaFrame p, int res, float aabb0, float aabb1,
                                                  float aabb2, float ext0, float ext1, float ext2,
                                                  const float* __restrict__ jitter, float* __restrict__ occs,
                                                  unsigned long long* __restrict__ counters, float* __restrict__ ema = nullptr,
                                                  float decay = 0.f) {
    extern __shared__ __align__(16) float smem[];
    ia_stage(smem, p.mlp, IA_GEO_END);
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int nvox = res * res * res;
    int teams = gridDim.x * (blockDim.x / IA_TEAM);
    unsigned nq = 0, nfetch = 0, ngeo = 0;
    for (int v = blockIdx.x * (blockDim.x / IA_TEAM) + threadIdx.x / IA_TEAM; v < nvox; v += teams) {
        int cz = v % res, cy = (v / res) % res, cx = v / (res * res);
        float best = 0.f;
        for (int s = 0; s < N_S; s++) {
            const float* j = jitter + ((size_t)v * N_S + s) * 3;
            float x[3];
            x[0] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)cx, j[0]), (float)res), ext0), aabb0);
            x[1] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)cy, j[1]), (float)res), ext1), aabb1);
            x[2] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)cz, j[2]), (float)res), ext2), aabb2);
            IaQuery q;
            ia_team_query<false>(team, p, smem, x, q);
            nq++; nfetch += q.n_fetch; ngeo += team.thread_rank() == 0 ? q.n_valid : 0;
            float a = ia_alpha(q.sdf, p.step_primary, p.beta);
            best = s == 0 ? a : fmaxf(best, a);
        }
        if (team.thread_rank() == 0) {
            if (ema) { best = fmaxf(ema[v] * decay, best); ema[v] = best; }
            occs[v] = best;
        }
    }
    if (team.thread_rank() == 0 && nq) atomicAdd(&counters[IA_CNT_QUERIES], nq);
    if (nfetch) atomicAdd(&counters[IA_CNT_BROYDEN_FETCH], nfetch);
    if (ngeo) atomicAdd(&counters[IA_CNT_GEO_EVAL], ngeo);
}

__global__ void k_occ_maxpool(const float* __restrict__ occs, float* __restrict__ pooled, int res, double* __restrict__ sum) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    int nvox = res * res * res;
    float m = -INFINITY;
    if (v < nvox) {
        int z = v % res, y = (v / res) % res, x = v / (res * res);
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx >= 0 && xx < res && yy >= 0 && yy < res && zz >= 0 && zz < res)
                        m = fmaxf(m, occs[(xx * res + yy) * res + zz]);
                }
        pooled[v] = m;
    }
    // block sum of the entries >= 0 (reference: occs_[occs_ >= 0].mean(); alphas are always >= 0)
    double s = (v < nvox && m >= 0) ? (double)m : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < (blockDim.x + 31) / 32; i++) t += ws[i];
        atomicAdd(sum, t);
    }
}

__global__ void k_occ_thresh(const float* __restrict__ pooled, float* __restrict__ comp, int nvox, const double* __restrict__ sum,
                             float occ_thre) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    float thre = fminf((float)(*sum / (double)nvox), occ_thre);
    comp[v] = pooled[v] > thre ? (float)(v + 1) : 0.f;
}

// one max-propagation pass of max_connected_component (models/utils.py:152-163)
__global__ void k_cc_iter(const float* __restrict__ in, float* __restrict__ out, int res) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    int nvox = res * res * res;
    if (v >= nvox) return;
    float self = in[v];
    if (self == 0.f) { out[v] = 0.f; return; }
    int z = v % res, y = (v / res) % res, x = v / (res * res);
    float m = self;
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dz = -1; dz <= 1; dz++) {
                int xx = x + dx, yy = y + dy, zz = z + dz;
                if (xx >= 0 && xx < res && yy >= 0 && yy < res && zz >= 0 && zz < res)
                    m = fmaxf(m, in[(xx * res + yy) * res + zz]);
            }
    out[v] = m;
}

__global__ void k_cc_hist(const float* __restrict__ comp, int* __restrict__ hist, int nvox) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    float c = comp[v];
    if (c > 0.f) atomicAdd(&hist[(int)c - 1], 1);
}

// torch.mode: most frequent label (smallest on ties); result in hist[nvox]
__global__ void k_cc_pick(int* __restrict__ hist, int nvox) {
    __shared__ int best_cnt[256], best_lbl[256];
    int bc = 0, bl = 0;
    for (int i = threadIdx.x; i < nvox; i += blockDim.x) {
        int h = hist[i];
        if (h > bc) { bc = h; bl = i + 1; }
    }
    best_cnt[threadIdx.x] = bc; best_lbl[threadIdx.x] = bl;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < blockDim.x; i++)
            if (best_cnt[i] > bc || (best_cnt[i] == bc && best_cnt[i] > 0 && best_lbl[i] < bl)) { bc = best_cnt[i]; bl = best_lbl[i]; }
        hist[nvox] = bl;
    }
}

__global__ void k_cc_final(const float* __restrict__ comp, const int* __restrict__ hist, int nvox, uint32_t* __restrict__ bits,
                           uint8_t* __restrict__ bytes) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    bool on = false;
    if (v < nvox) {
        int lbl = hist[nvox];
        on = lbl > 0 && comp[v] == (float)lbl;
        if (bytes) bytes[v] = on;
    }
    unsigned b = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && v < nvox) bits[v >> 5] = b;
}

__global__ void k_occ_from_bytes(const uint8_t* __restrict__ bytes, int nvox, uint32_t* __restrict__ bits) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    bool on = v < nvox && bytes[v] != 0;
    unsigned b = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && v < nvox) bits[v >> 5] = b;
}

static int ia_occ_alloc(ia_ctx* c, int res) {
    int nvox = res * res * res;
    IA_REQUIRE(nvox % 32 == 0, IA_EINVAL, "occupancy resolution^3 must be a multiple of 32");
    if (c->occ_res_alloc != res) {
        if (ia_realloc(&c->d_occ_bits, (size_t)nvox / 32) || ia_realloc(&c->d_occ_a, (size_t)nvox) ||
            ia_realloc(&c->d_occ_b, (size_t)nvox) || ia_realloc(&c->d_occ_hist, (size_t)nvox + 1) ||
            ia_realloc(&c->d_occ_sum, 1))
            return IA_ECUDA;
        c->occ_res_alloc = res;
    }
    c->f.occ_bits = c->d_occ_bits;
    c->f.occ_res = res;
    return IA_OK;
}

static int ia_occupancy_pipeline(ia_ctx* c, const float* aabb, int res, const float* d_jitter, float* d_ema, float decay,
                                 float occ_thre, uint8_t* d_bin_out, cudaStream_t st) {
    if (int e = ia_occ_alloc(c, res)) return e;
    for (int i = 0; i < 6; i++) c->f.aabb[i] = aabb[i];
    int nvox = res * res * res;
    size_t sm = IA_GEO_END * sizeof(float);
    IA_STAGE_BEGIN(c, IA_STAGE_OCCUPANCY, st);
    if (d_ema)
        k_occ_eval<1><<<c->n_sm * 8, 256, sm, st>>>(c->f, res, aabb[0], aabb[1], aabb[2], aabb[3] - aabb[0], aabb[4] - aabb[1],
                                                    aabb[5] - aabb[2], d_jitter, c->d_occ_a, c->d_counters, d_ema, decay);
    else
        k_occ_eval<3><<<c->n_sm * 8, 256, sm, st>>>(c->f, res, aabb[0], aabb[1], aabb[2], aabb[3] - aabb[0], aabb[4] - aabb[1],
                                                    aabb[5] - aabb[2], d_jitter, c->d_occ_a, c->d_counters);
    IA_LAUNCH_CHECK();
    IA_CHECK_CUDA(cudaMemsetAsync(c->d_occ_sum, 0, sizeof(double), st));
    int nb = (nvox + 255) / 256;
    k_occ_maxpool<<<nb, 256, 0, st>>>(c->d_occ_a, c->d_occ_b, res, c->d_occ_sum);
    k_occ_thresh<<<nb, 256, 0, st>>>(c->d_occ_b, c->d_occ_a, nvox, c->d_occ_sum, occ_thre);
    float *a = c->d_occ_a, *b = c->d_occ_b;
    for (int it = 0; it < res * 3; it++) {
        k_cc_iter<<<nb, 256, 0, st>>>(a, b, res);
        std::swap(a, b);
    }
    IA_CHECK_CUDA(cudaMemsetAsync(c->d_occ_hist, 0, ((size_t)nvox + 1) * sizeof(int), st));
    k_cc_hist<<<nb, 256, 0, st>>>(a, c->d_occ_hist, nvox);
    k_cc_pick<<<1, 256, 0, st>>>(c->d_occ_hist, nvox);
    k_cc_final<<<nb, 256, 0, st>>>(a, c->d_occ_hist, nvox, c->d_occ_bits, d_bin_out);
    IA_STAGE_END(c, IA_STAGE_OCCUPANCY, st, 6 + res * 3);
    IA_LAUNCH_CHECK();
    c->have_occ = true;
    return IA_OK;
}

extern "C" int ia_build_occupancy(ia_ctx* c, const float* aabb, int res, const float* d_jitter, uint8_t* d_bin_out,
                                  void* stream) {
    IA_REQUIRE(c && aabb && d_jitter, IA_EINVAL, "ia_build_occupancy: NULL argument");
    IA_REQUIRE(c->have_pose && c->have_fields && c->have_cfg, IA_ESTATE, "ia_build_occupancy: fields/pose/config not set");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    return ia_occupancy_pipeline(c, aabb, res, d_jitter, nullptr, 0.f, c->occ_thre, d_bin_out, (cudaStream_t)stream);
}

extern "C" int ia_update_occupancy_ema(ia_ctx* c, const float* aabb, int res, const float* d_jitter, float* d_occs,
                                       float ema_decay, float occ_thre, uint8_t* d_bin_out, void* stream) {
    IA_REQUIRE(c && aabb && d_jitter && d_occs, IA_EINVAL, "ia_update_occupancy_ema: NULL argument");
    IA_REQUIRE(c->have_pose && c->have_fields && c->have_cfg, IA_ESTATE, "ia_update_occupancy_ema: fields/pose/config not set");
    IA_REQUIRE(ema_decay >= 0.f && ema_decay <= 1.f && occ_thre > 0.f, IA_EINVAL, "ia_update_occupancy_ema: bad decay / threshold");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    return ia_occupancy_pipeline(c, aabb, res, d_jitter, d_occs, ema_decay, occ_thre, d_bin_out, (cudaStream_t)stream);
}

extern "C" int ia_set_occupancy(ia_ctx* c, const float* aabb, int res, const uint8_t* d_bin, void* stream) {
    IA_REQUIRE(c && aabb && d_bin, IA_EINVAL, "ia_set_occupancy: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    if (int e = ia_occ_alloc(c, res)) return e;
    for (int i = 0; i < 6; i++) c->f.aabb[i] = aabb[i];
    int nvox = res * res * res;
    k_occ_from_bytes<<<(nvox + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_bin, nvox, c->d_occ_bits);
    IA_LAUNCH_CHECK();
    c->have_occ = true;
    return IA_OK;
}

// ================================================================================================
// environment light (lib/torch_pbr/light.py:259-446)
__global__ void k_env_pdf(const float* __restrict__ env, int H, int W, float* __restrict__ pdf, float* __restrict__ rowsum,
                          double* __restrict__ total) {
    int r = blockIdx.x;
    float sy = sinf(((float)r + 0.5f) / (float)H * 3.14159265358979323846f);
    double s = 0;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const float* px = env + ((size_t)r * W + x) * 3;
        float v = fmaxf(px[0], fmaxf(px[1], px[2])) * sy;
        if (v <= 0) v = 1e-6f;
        pdf[(size_t)r * W + x] = v;
        s += v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < (blockDim.x + 31) / 32; i++) t += ws[i];
        rowsum[r] = (float)t;
        atomicAdd(total, t);
    }
}

// per row: pdf /= total; cols = cumsum(pdf) / rowtotal with a leading 0  (one warp per row)
__global__ void k_env_cols(float* __restrict__ pdf, int H, int W, const double* __restrict__ total, float* __restrict__ cols,
                           float* __restrict__ rowtot) {
    int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (r >= H) return;
    int lane = threadIdx.x & 31;
    float inv_total = (float)(1.0 / *total);
    float* prow = pdf + (size_t)r * W;
    float* crow = cols + (size_t)r * (W + 1);
    float carry = 0.f;
    for (int x0 = 0; x0 < W; x0 += 32) {
        int x = x0 + lane;
        float v = x < W ? prow[x] * inv_total : 0.f;
        if (x < W) prow[x] = v;
        float s = v;
        for (int o = 1; o < 32; o <<= 1) {
            float t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        s += carry;
        if (x < W) crow[x + 1] = s;
        carry = __shfl_sync(0xffffffffu, s, 31);
    }
    float tot = carry;
    float den = tot > 0 ? tot : 1.f;
    for (int x = lane; x < W; x += 32) crow[x + 1] = crow[x + 1] / den;
    if (lane == 0) { crow[0] = 0.f; rowtot[r] = tot; }
}

// rows cdf (single block, serial over H by thread 0; H ~ 1k)
__global__ void k_env_rows(const float* __restrict__ rowtot, int H, float* __restrict__ rows) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float s = 0.f;
    rows[0] = 0.f;
    for (int r = 0; r < H; r++) { s += rowtot[r]; rows[r + 1] = s; }
    float den = s > 0 ? s : 1.f;
    for (int r = 0; r < H; r++) rows[r + 1] /= den;
}

__global__ void k_env_sample(const __grid_constant__ IaFrame p, const IaEnv E, const float* __restrict__ u1,
                             const float* __restrict__ u2, int spp, float* __restrict__ dir_w, float* __restrict__ dir_s,
                             float* __restrict__ em, float* __restrict__ pdf_out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= spp) return;
    float dn[3];
    ia_env_sample(E, u1[k], u2[k], dn);
    dir_w[k * 3 + 0] = dn[0]; dir_w[k * 3 + 1] = dn[1]; dir_w[k * 3 + 2] = dn[2];
    // direction used by the integrator: w2s, then back s2w for the lookups (pbr_light_forward :784-842)
    float ds[3], dw[3];
    ia_dir_w2s(p, dn, ds);
    dir_s[k * 3 + 0] = ds[0]; dir_s[k * 3 + 1] = ds[1]; dir_s[k * 3 + 2] = ds[2];
    ia_dir_s2w(p, ds, dw);
    pdf_out[k] = ia_env_pdf(E, dw);
    float e3[3];
    ia_env_eval(E, dw, e3);
    em[k * 3 + 0] = e3[0]; em[k * 3 + 1] = e3[1]; em[k * 3 + 2] = e3[2];
}

// render_mode = uniform_light: the n_rows x n_cols stratified sphere directions of
// sample_uniform_sphere_stratified in eval mode (lib/torch_pbr/light.py:161-217; sample_uniform_sphere,
// warp_utils.py:174-198).  They are used as SMPL-frame ray directions as they are
// (models/intrinsic_avatar.py:672-686); the envmap is looked up at transform_dirs_s2w(d) (:724-730).
__global__ void k_env_uniform(const __grid_constant__ IaFrame p, const IaEnv E, int n_rows, int n_cols,
                              float* __restrict__ dir_w, float* __restrict__ dir_s, float* __restrict__ em,
                              float* __restrict__ pdf_out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_rows * n_cols) return;
    const float PI = 3.14159265358979323846f;
    const int row = k / n_cols, col = k - row * n_cols;
    const float v = ((float)row + 0.5f) / (float)n_rows, u = ((float)col + 0.5f) / (float)n_cols;
    const float z = v * 2.0f - 1.0f;
    const float phi = 2.0f * PI * u;
    const float r = sqrtf(fmaxf(1.0f - z * z, 0.f));
    float d[3] = {cosf(phi) * r, sinf(phi) * r, z};
    ia_normalize12(d);
    dir_s[k * 3 + 0] = d[0]; dir_s[k * 3 + 1] = d[1]; dir_s[k * 3 + 2] = d[2];
    float dw[3];
    ia_dir_s2w(p, d, dw);
    dir_w[k * 3 + 0] = dw[0]; dir_w[k * 3 + 1] = dw[1]; dir_w[k * 3 + 2] = dw[2];
    float e3[3];
    ia_env_eval(E, dw, e3);
    em[k * 3 + 0] = e3[0]; em[k * 3 + 1] = e3[1]; em[k * 3 + 2] = e3[2];
    pdf_out[k] = 1.0f / (4.0f * PI);
}

static int ia_light_alloc(ia_ctx* c, int H, int W, int spp, cudaStream_t st) {
    if (c->env_H != H || c->env_W != W) {
        if (ia_realloc(&c->d_env_pdf, (size_t)H * W) || ia_realloc(&c->d_env_cols, (size_t)H * (W + 1)) ||
            ia_realloc(&c->d_env_rows, (size_t)H + 1) || ia_realloc(&c->d_env_rowsum, (size_t)H) ||
            ia_realloc(&c->d_env_total, 1))
            return IA_ECUDA;
        c->env_H = H; c->env_W = W;
    }
    if (c->spp != spp) {
        if (ia_realloc(&c->d_light_dir_w, (size_t)spp * 3) || ia_realloc(&c->d_light_dir_s, (size_t)spp * 3) ||
            ia_realloc(&c->d_light_em, (size_t)spp * 3) || ia_realloc(&c->d_light_pdf, (size_t)spp) ||
            ia_realloc(&c->d_u_table, (size_t)spp))
            return IA_ECUDA;
        // stratified CDF positions of cdf_resampling_kernel, same float recurrence (cdf.cu:53-58,105)
        std::vector<float> u(spp);
        float step = (1.0f - 1.0 / spp) / (spp - 1);
        float cu = 1.0 / (2 * spp);
        for (int j = 0; j < spp; j++) { u[j] = cu; cu += step; }
        IA_CHECK_CUDA(cudaMemcpyAsync(c->d_u_table, u.data(), spp * sizeof(float), cudaMemcpyHostToDevice, st));
        IA_CHECK_CUDA(cudaStreamSynchronize(st));
        c->spp = spp;
    }
    return IA_OK;
}

// pdf table and row / column CDFs of the envmap (update_pdf, light.py:417-446)
static int ia_env_tables(ia_ctx* c, const float* d_env, int H, int W, cudaStream_t st) {
    IA_CHECK_CUDA(cudaMemsetAsync(c->d_env_total, 0, sizeof(double), st));
    k_env_pdf<<<H, 256, 0, st>>>(d_env, H, W, c->d_env_pdf, c->d_env_rowsum, c->d_env_total);
    k_env_cols<<<(H + 7) / 8, 256, 0, st>>>(c->d_env_pdf, H, W, c->d_env_total, c->d_env_cols, c->d_env_rowsum);
    k_env_rows<<<1, 32, 0, st>>>(c->d_env_rowsum, H, c->d_env_rows);
    c->env.env = d_env; c->env.pdf = c->d_env_pdf; c->env.cols = c->d_env_cols; c->env.rows = c->d_env_rows;
    c->env.H = H; c->env.W = W;
    return IA_OK;
}

static int ia_light_outputs(ia_ctx* c, int spp, float* d_dirs_out, float* d_em_out, float* d_pdf_out, cudaStream_t st) {
    if (d_dirs_out) IA_CHECK_CUDA(cudaMemcpyAsync(d_dirs_out, c->d_light_dir_w, (size_t)spp * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (d_em_out) IA_CHECK_CUDA(cudaMemcpyAsync(d_em_out, c->d_light_em, (size_t)spp * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (d_pdf_out) IA_CHECK_CUDA(cudaMemcpyAsync(d_pdf_out, c->d_light_pdf, (size_t)spp * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return IA_OK;
}

extern "C" int ia_set_light(ia_ctx* c, const float* d_env, int H, int W, const float* d_u1, const float* d_u2, int spp,
                            float* d_dirs_out, float* d_em_out, float* d_pdf_out, void* stream) {
    IA_REQUIRE(c && d_env && d_u1 && d_u2, IA_EINVAL, "ia_set_light: NULL argument");
    IA_REQUIRE(spp > 1, IA_EINVAL, "ia_set_light: samples_per_pixel must be > 1 (lib/nerfacc/cdf.py:51)");
    IA_REQUIRE(spp < 65536, IA_EINVAL, "ia_set_light: samples_per_pixel must be < 65536");
    IA_REQUIRE(c->have_pose, IA_ESTATE, "ia_set_light: call ia_set_pose first");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = ia_light_alloc(c, H, W, spp, st)) return e;
    IA_STAGE_BEGIN(c, IA_STAGE_LIGHT, st);
    if (int e = ia_env_tables(c, d_env, H, W, st)) return e;
    k_env_sample<<<(spp + 127) / 128, 128, 0, st>>>(c->f, c->env, d_u1, d_u2, spp, c->d_light_dir_w, c->d_light_dir_s,
                                                    c->d_light_em, c->d_light_pdf);
    IA_STAGE_END(c, IA_STAGE_LIGHT, st, 4);
    IA_LAUNCH_CHECK();
    if (int e = ia_light_outputs(c, spp, d_dirs_out, d_em_out, d_pdf_out, st)) return e;
    c->have_light = true;
    c->light_uniform = false;
    return IA_OK;
}

extern "C" int ia_set_light_uniform(ia_ctx* c, const float* d_env, int H, int W, int n_rows, int n_cols,
                                    float* d_dirs_out, float* d_em_out, void* stream) {
    IA_REQUIRE(c && d_env, IA_EINVAL, "ia_set_light_uniform: NULL argument");
    IA_REQUIRE(n_rows > 0 && n_cols > 0 && n_rows * n_cols > 1 && n_rows * n_cols < 65536, IA_EINVAL,
               "ia_set_light_uniform: need 1 < n_rows * n_cols < 65536");
    IA_REQUIRE(c->have_pose, IA_ESTATE, "ia_set_light_uniform: call ia_set_pose first");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int spp = n_rows * n_cols;
    if (int e = ia_light_alloc(c, H, W, spp, st)) return e;
    IA_STAGE_BEGIN(c, IA_STAGE_LIGHT, st);
    if (int e = ia_env_tables(c, d_env, H, W, st)) return e;
    k_env_uniform<<<(spp + 127) / 128, 128, 0, st>>>(c->f, c->env, n_rows, n_cols, c->d_light_dir_w, c->d_light_dir_s,
                                                     c->d_light_em, c->d_light_pdf);
    IA_STAGE_END(c, IA_STAGE_LIGHT, st, 4);
    IA_LAUNCH_CHECK();
    if (int e = ia_light_outputs(c, spp, d_dirs_out, d_em_out, nullptr, st)) return e;
    c->have_light = true;
    c->light_uniform = true;
    return IA_OK;
}

// ================================================================================================
// subject set-up on the device: SMPL linear blend skinning, skinning-weight voxelisation
#include "ia_body.cuh"

extern "C" int ia_smpl_lbs(ia_ctx* c, const float* d_v_template, const float* d_shapedirs, const float* d_posedirs,
                           const float* d_J_regressor, const float* d_lbs_weights, const int32_t* h_parents, int V, int NB,
                           const float* h_betas, const float* h_pose72, const float* h_transl3, float* d_vertices,
                           float* d_joints, float* d_A, void* stream) {
    IA_REQUIRE(c && d_v_template && d_shapedirs && d_posedirs && d_J_regressor && d_lbs_weights && h_parents && h_betas && h_pose72 &&
                   h_transl3 && d_vertices && d_joints && d_A, IA_EINVAL, "ia_smpl_lbs: NULL argument");
    IA_REQUIRE(V > 0 && NB >= 0 && NB <= 300, IA_EINVAL, "ia_smpl_lbs: bad sizes");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    // scratch: v_shaped [V*3] | joints_rest [72] | pose_feature [207] | A_rel [384] | betas [NB] | pose [72] | transl [3] | parents [24]
    const size_t n_f = (size_t)V * 3 + 72 + 208 + 384 + NB + 72 + 4;
    float* w = nullptr;
    IA_CHECK_CUDA(cudaMallocAsync((void**)&w, n_f * sizeof(float) + 24 * sizeof(int), st));
    float *v_shaped = w, *joints_rest = w + (size_t)V * 3, *pose_feature = joints_rest + 72, *A_rel = pose_feature + 208,
          *betas = A_rel + 384, *pose = betas + NB, *transl = pose + 72;
    int* parents = reinterpret_cast<int*>(w + n_f);
    if (NB) IA_CHECK_CUDA(cudaMemcpyAsync(betas, h_betas, NB * sizeof(float), cudaMemcpyHostToDevice, st));
    IA_CHECK_CUDA(cudaMemcpyAsync(pose, h_pose72, 72 * sizeof(float), cudaMemcpyHostToDevice, st));
    IA_CHECK_CUDA(cudaMemcpyAsync(transl, h_transl3, 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    IA_CHECK_CUDA(cudaMemcpyAsync(parents, h_parents, 24 * sizeof(int), cudaMemcpyHostToDevice, st));
    k_lbs_shape<<<(V * 3 + 255) / 256, 256, 0, st>>>(d_v_template, d_shapedirs, betas, V, NB, v_shaped);
    k_lbs_joints<<<24, 256, 0, st>>>(d_J_regressor, v_shaped, V, joints_rest);
    k_lbs_chain<<<1, 32, 0, st>>>(pose, joints_rest, parents, transl, pose_feature, d_joints, d_A, A_rel);
    k_lbs_skin<<<(V + 127) / 128, 128, 0, st>>>(v_shaped, d_posedirs, pose_feature, d_lbs_weights, A_rel, transl, V, d_vertices);
    IA_LAUNCH_CHECK();
    IA_CHECK_CUDA(cudaStreamSynchronize(st));   // the small host arrays above must outlive the copies
    IA_CHECK_CUDA(cudaFreeAsync(w, st));
    return IA_OK;
}

extern "C" int ia_voxelize_lbs(ia_ctx* c, const float* d_verts, const float* d_weights, int V, int resolution,
                               float* d_lbs_voxel, float* h_offset_kernel3, float* h_scale_kernel3, void* stream) {
    IA_REQUIRE(c && d_verts && d_weights && d_lbs_voxel && h_offset_kernel3 && h_scale_kernel3, IA_EINVAL,
               "ia_voxelize_lbs: NULL argument");
    IA_REQUIRE(V > 0 && resolution >= 8 && resolution % 4 == 0, IA_EINVAL, "ia_voxelize_lbs: resolution must be a multiple of 4");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int D = resolution / 4, H = resolution, W = resolution, P = D * H * W;
    float* tmp = nullptr;
    IA_CHECK_CUDA(cudaMallocAsync((void**)&tmp, ((size_t)24 * P + 8) * sizeof(float), st));
    float* mnmx = tmp + (size_t)24 * P;
    k_vox_bbox<<<1, 256, 0, st>>>(d_verts, V, mnmx);
    float h[6];
    IA_CHECK_CUDA(cudaMemcpyAsync(h, mnmx, sizeof(h), cudaMemcpyDeviceToHost, st));
    IA_CHECK_CUDA(cudaStreamSynchronize(st));
    // deformer_torch.py:150-167 (fp32 like the reference's torch ops)
    const float off[3] = {(h[0] + h[3]) * 0.5f, (h[1] + h[4]) * 0.5f, (h[2] + h[5]) * 0.5f};
    const float scale = fmaxf(fmaxf(h[3] - h[0], h[4] - h[1]), h[5] - h[2]) / 2 * 1.2f;
    const float ratio = (float)H / (float)D;
    k_vox_knn<<<(P + 127) / 128, 128, 0, st>>>(d_verts, d_weights, V, D, H, W, off[0], off[1], off[2], scale, ratio, d_lbs_voxel);
    float *a = d_lbs_voxel, *b = tmp;
    for (int it = 0; it < 30; it++) {
        k_vox_smooth<<<(P + 255) / 256, 256, 0, st>>>(a, b, D, H, W);
        std::swap(a, b);
    }
    // 30 passes: the result is back in d_lbs_voxel
    IA_LAUNCH_CHECK();
    IA_CHECK_CUDA(cudaFreeAsync(tmp, st));
    for (int i = 0; i < 3; i++) {
        h_offset_kernel3[i] = -off[i];
        h_scale_kernel3[i] = 1.0f / scale;
    }
    h_scale_kernel3[2] *= ratio;
    return IA_OK;
}

// ================================================================================================
// training-mode building block: backward of the canonical geometry network
#include "ia_train.cuh"

extern "C" int ia_op_geometry_backward(ia_ctx* c, const float* d_xc, const float* d_dout, int64_t n, float* d_g_hash,
                                       float* d_g_mlp, float* d_g_x, void* stream) {
    IA_REQUIRE(c && d_xc && d_dout && d_g_hash && d_g_mlp, IA_EINVAL, "ia_op_geometry_backward: NULL argument");
    IA_REQUIRE(c->have_fields, IA_ESTATE, "ia_op_geometry_backward: call ia_set_fields first");
    if (n == 0) return IA_OK;
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    const size_t sm = 2 * IA_GEO_END * sizeof(float);
    IA_CHECK_CUDA(cudaFuncSetAttribute(k_geometry_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_geometry_backward<<<ia_query_blocks(c, n), 256, sm, (cudaStream_t)stream>>>(c->f, d_xc, d_dout, n, d_g_hash, d_g_mlp, d_g_x,
                                                                                  nullptr);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_shade_fields_backward(ia_ctx* c, const float* d_xc, const float* d_feature, const float* d_view,
                                           const float* d_normal, const float* d_drgb, const float* d_dmat, int64_t n,
                                           float* d_g_rad_hash, float* d_g_mlp, float* d_g_x, float* d_g_feature,
                                           float* d_g_normal, void* stream) {
    IA_REQUIRE(c && d_g_rad_hash && d_g_mlp && n >= 0, IA_EINVAL, "ia_op_shade_fields_backward: NULL argument");
    IA_REQUIRE(c->have_fields, IA_ESTATE, "ia_op_shade_fields_backward: call ia_set_fields first");
    if (n == 0) return IA_OK;
    IA_REQUIRE(d_xc && d_feature && d_view && d_normal && d_drgb && d_dmat, IA_EINVAL,
               "ia_op_shade_fields_backward: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    const size_t sm = ((IA_SHB_THREADS / IA_TEAM) * IA_SHB_TEAM_FLOATS + IA_SHADE_GRAD_FLOATS) * sizeof(float);
    IA_CHECK_CUDA(cudaFuncSetAttribute(k_shade_fields_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int per_cta = IA_SHB_THREADS / IA_TEAM;       // one CTA per SM: its threads own the weight-gradient accumulators
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + per_cta - 1) / per_cta, (int64_t)c->n_sm));
    k_shade_fields_backward<<<blocks, IA_SHB_THREADS, sm, (cudaStream_t)stream>>>(
        c->f, d_xc, d_feature, d_view, d_normal, d_drgb, d_dmat, n, d_g_rad_hash, d_g_mlp, d_g_x, d_g_feature, d_g_normal);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_volrend(ia_ctx* c, const int32_t* d_packed_info, const float* d_sdf, const float* d_dists,
                             const float* d_values, int n_channels, float beta, int64_t n_rays, float* d_weights,
                             float* d_comp, float* d_opacity, void* stream) {
    IA_REQUIRE(c && n_rays >= 0 && n_channels >= 0 && n_channels <= IA_VOLREND_MAX_C && beta > 0.f, IA_EINVAL,
               "ia_op_volrend: bad argument (channels <= 16, beta > 0)");
    if (n_rays == 0) return IA_OK;
    IA_REQUIRE(d_packed_info && d_sdf && d_dists && (d_values || n_channels == 0) && (d_comp || n_channels == 0) && d_opacity,
               IA_EINVAL, "ia_op_volrend: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    k_volrend<<<(unsigned)((n_rays + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_packed_info, d_sdf, d_dists, d_values,
                                                                                 n_channels, beta, n_rays, d_weights, d_comp,
                                                                                 d_opacity);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_volrend_backward(ia_ctx* c, const int32_t* d_packed_info, const float* d_sdf, const float* d_dists,
                                      const float* d_values, int n_channels, float beta, const float* d_dcomp,
                                      const float* d_dopacity, const float* d_dweights, int64_t n_rays, float* d_g_sdf,
                                      float* d_g_values, float* d_g_beta, void* stream) {
    IA_REQUIRE(c && n_rays >= 0 && n_channels >= 0 && n_channels <= IA_VOLREND_MAX_C && beta > 0.f && d_g_beta, IA_EINVAL,
               "ia_op_volrend_backward: bad argument (channels <= 16, beta > 0)");
    if (n_rays == 0) return IA_OK;
    IA_REQUIRE(d_packed_info && d_sdf && d_dists && (d_values || n_channels == 0) && (d_dcomp || n_channels == 0) && d_g_sdf,
               IA_EINVAL, "ia_op_volrend_backward: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    k_volrend_backward<<<(unsigned)((n_rays + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        d_packed_info, d_sdf, d_dists, d_values, n_channels, beta, d_dcomp, d_dopacity, d_dweights, n_rays, d_g_sdf, d_g_values,
        d_g_beta);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_query_train(ia_ctx* c, const float* d_xd, int64_t n, float* d_sdf, float* d_xc, uint8_t* d_valid,
                                 float* d_grad, float* d_grad_cano, float* d_feature, float* d_J_inv, int32_t* d_best,
                                 void* stream) {
    IA_REQUIRE(c && n >= 0, IA_EINVAL, "ia_op_query_train: NULL argument");
    IA_REQUIRE(c->have_pose && c->have_fields, IA_ESTATE, "ia_op_query_train: fields and pose must be set");
    if (n == 0) return IA_OK;
    IA_REQUIRE(d_xd && d_sdf && d_xc && d_valid && d_J_inv && d_best, IA_EINVAL, "ia_op_query_train: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    const size_t sm = IA_GEO_END * sizeof(float);
    k_query_train<<<ia_query_blocks(c, n), 256, sm, (cudaStream_t)stream>>>(c->f, d_xd, n, d_sdf, d_xc, d_valid, d_grad,
                                                                            d_grad_cano, d_feature, d_J_inv, d_best);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_query_backward(ia_ctx* c, const float* d_xc, const uint8_t* d_valid, const float* d_J_inv,
                                    const float* d_dout, int64_t n, float* d_g_hash, float* d_g_mlp, float* d_g_tfs,
                                    float* d_g_x, void* stream) {
    IA_REQUIRE(c && d_g_hash && d_g_mlp && d_g_tfs && n >= 0, IA_EINVAL, "ia_op_query_backward: NULL argument");
    IA_REQUIRE(c->have_pose && c->have_fields, IA_ESTATE, "ia_op_query_backward: fields and pose must be set");
    if (n == 0) return IA_OK;
    IA_REQUIRE(d_xc && d_valid && d_J_inv && d_dout && d_g_x, IA_EINVAL, "ia_op_query_backward: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    const size_t sm = 2 * IA_GEO_END * sizeof(float);
    IA_CHECK_CUDA(cudaFuncSetAttribute(k_geometry_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    // network backward at the arg-min roots (gradient of hash table and weights, and dL/dx_c), then dL/dx_c through the
    // implicit-differentiation correction to the bone transforms
    k_geometry_backward<<<ia_query_blocks(c, n), 256, sm, (cudaStream_t)stream>>>(c->f, d_xc, d_dout, n, d_g_hash, d_g_mlp, d_g_x,
                                                                                  d_valid);
    IA_LAUNCH_CHECK();
    k_deform_backward<<<ia_query_blocks(c, n), 256, 0, (cudaStream_t)stream>>>(c->f, d_xc, d_valid, d_J_inv, d_g_x, n, d_g_tfs);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_deform_backward(ia_ctx* c, const float* d_xc, const uint8_t* d_valid, const float* d_J_inv,
                                     const float* d_g_xc, int64_t n, float* d_g_tfs, void* stream) {
    IA_REQUIRE(c && d_g_tfs && n >= 0, IA_EINVAL, "ia_op_deform_backward: NULL argument");
    IA_REQUIRE(c->have_lbs && c->have_pose, IA_ESTATE, "ia_op_deform_backward: call ia_set_lbs_voxels and ia_set_pose first");
    if (n == 0) return IA_OK;     // (an empty batch has no buffers)
    IA_REQUIRE(d_xc && d_valid && d_J_inv && d_g_xc, IA_EINVAL, "ia_op_deform_backward: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    k_deform_backward<<<ia_query_blocks(c, n * IA_N_INIT), 256, 0, (cudaStream_t)stream>>>(c->f, d_xc, d_valid, d_J_inv, d_g_xc,
                                                                                         n * IA_N_INIT, d_g_tfs);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

#include "ia_render.cuh"
#include "ia_train_pbr.cuh"
