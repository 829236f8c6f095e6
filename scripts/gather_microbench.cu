// Microbenchmark (run by hand on the GPU box): what does the L1 charge for the trilinear voxel_J gather of the Broyden
// phase, as a function of the voxel FORMAT?  Persistent 148 x 512-thread CTAs (128 registers, as k_shade_wf), every
// thread runs dependent fetches (the next cell depends on the previous result, as a Broyden chain does); each fetch reads
// the 8 corners of a cell of a 32 x 128 x 128 grid and accumulates them with the trilinear weights.
//   format 0: 48-B voxel, 3 x LDG.128 per corner           (product, IA_FETCH_MODE 0)
//   format 1: 32-B voxel, 1 x LDG.256 per corner
//   format 2: 32-B voxel, 2 x LDG.128 per corner
//   format 3: 24-B voxel (12 x fp16), LDG.128 + LDG.64 per corner (8-B aligned: 3 x LDG.64)
//   format 4: 16-B voxel, 1 x LDG.128 per corner           (lower bound of one request per corner)
//   format 5: 64-B voxel, 2 x LDG.256 per corner
//   format 6: 48-B voxel, 3 x tex1Dfetch<float4> per corner (texture path: bit-exact point fetch from linear memory)
//   format 7: 48-B voxel, 2 x LDG.128 + 1 x tex1Dfetch per corner (do the LSU and TEX address stages run in parallel?)
//   format 8: 48-B voxel, 3 x plain ld.global (not .nc) per corner
// `spread`: the lanes of a warp pick cells within a cube of that side around a per-warp base cell (1 = all lanes the same
// cell, 128 = independent random cells); `smem_kb` of dynamic shared memory shrink the L1 like the kernel's own use does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_microbench scripts/gather_microbench.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define D_ 32
#define H_ 128
#define W_ 128

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float4 ld_plain(const float4* ptr) {   // ld.global (coherent path), not hoistable past stores
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr));
    return v;
}
__device__ __forceinline__ void ld256(const void* ptr, float r[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "l"(ptr));
}

template <int FMT>
__global__ void __launch_bounds__(512, 1) k_gather(const unsigned char* __restrict__ tab, int iters, int spread, float* out,
                                                   cudaTextureObject_t tex) {
    extern __shared__ float pad[];
    const int lane = threadIdx.x & 31;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = 0.f;
    uint32_t s_warp = hash32(warp_id * 2654435761u + 1u), s_lane = hash32(s_warp ^ (lane * 0x9E3779B1u));
    if (pad && iters < 0) pad[threadIdx.x] = 1.f;
    float dep = 0.f;
    for (int it = 0; it < iters; it++) {
        s_warp = hash32(s_warp + 0x68bc21ebu);
        s_lane = hash32(s_lane + 0x02e5be93u + (uint32_t)(dep != 12345.f ? 0 : 1));  // dependent on the previous fetch
        const int bx = s_warp % (W_ - 1), by = (s_warp >> 8) % (H_ - 1), bz = (s_warp >> 16) % (D_ - 1);
        int x0 = bx + (int)(s_lane % spread), y0 = by + (int)((s_lane >> 8) % spread), z0 = bz + (int)((s_lane >> 16) % spread);
        x0 = min(x0, W_ - 2); y0 = min(y0, H_ - 2); z0 = min(z0, D_ - 2);
        const float wx = (s_lane & 1023) * (1.f / 1024), wy = ((s_lane >> 10) & 1023) * (1.f / 1024), wz = ((s_lane >> 20) & 1023) * (1.f / 1024);
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int xi = x0 + (c & 1), yi = y0 + ((c >> 1) & 1), zi = z0 + (c >> 2);
            const float w = ((c & 1) ? wx : 1 - wx) * ((c & 2) ? wy : 1 - wy) * ((c & 4) ? wz : 1 - wz);
            const size_t v = (size_t)((zi * H_ + yi) * W_ + xi);
            if (FMT == 0) {
                const float4* p = reinterpret_cast<const float4*>(tab + v * 48);
                float4 a = __ldg(p), b = __ldg(p + 1), cc = __ldg(p + 2);
                acc[0] = fmaf(a.x, w, acc[0]); acc[1] = fmaf(a.y, w, acc[1]); acc[2] = fmaf(a.z, w, acc[2]); acc[3] = fmaf(a.w, w, acc[3]);
                acc[4] = fmaf(b.x, w, acc[4]); acc[5] = fmaf(b.y, w, acc[5]); acc[6] = fmaf(b.z, w, acc[6]); acc[7] = fmaf(b.w, w, acc[7]);
                acc[8] = fmaf(cc.x, w, acc[8]); acc[9] = fmaf(cc.y, w, acc[9]); acc[10] = fmaf(cc.z, w, acc[10]); acc[11] = fmaf(cc.w, w, acc[11]);
            } else if (FMT == 1) {
                float r[8];
                ld256(tab + v * 32, r);
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] = fmaf(r[k], w, acc[k]);
            } else if (FMT == 2) {
                const float4* p = reinterpret_cast<const float4*>(tab + v * 32);
                float4 a = __ldg(p), b = __ldg(p + 1);
                acc[0] = fmaf(a.x, w, acc[0]); acc[1] = fmaf(a.y, w, acc[1]); acc[2] = fmaf(a.z, w, acc[2]); acc[3] = fmaf(a.w, w, acc[3]);
                acc[4] = fmaf(b.x, w, acc[4]); acc[5] = fmaf(b.y, w, acc[5]); acc[6] = fmaf(b.z, w, acc[6]); acc[7] = fmaf(b.w, w, acc[7]);
            } else if (FMT == 3) {
                const uint2* p = reinterpret_cast<const uint2*>(tab + v * 24);
                uint2 a = __ldg(p), b = __ldg(p + 1), cc = __ldg(p + 2);
                const uint32_t u[6] = {a.x, a.y, b.x, b.y, cc.x, cc.y};
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[k]));
                    acc[2 * k] = fmaf(f.x, w, acc[2 * k]); acc[2 * k + 1] = fmaf(f.y, w, acc[2 * k + 1]);
                }
            } else if (FMT == 6 || FMT == 7 || FMT == 8) {
                const float4* p = reinterpret_cast<const float4*>(tab + v * 48);
                float4 a, b, cc;
                if (FMT == 6) { a = tex1Dfetch<float4>(tex, (int)(v * 3)); b = tex1Dfetch<float4>(tex, (int)(v * 3 + 1)); }
                else if (FMT == 7) { a = __ldg(p); b = __ldg(p + 1); }
                else { a = ld_plain(p); b = ld_plain(p + 1); }
                if (FMT == 8) cc = ld_plain(p + 2); else cc = tex1Dfetch<float4>(tex, (int)(v * 3 + 2));
                acc[0] = fmaf(a.x, w, acc[0]); acc[1] = fmaf(a.y, w, acc[1]); acc[2] = fmaf(a.z, w, acc[2]); acc[3] = fmaf(a.w, w, acc[3]);
                acc[4] = fmaf(b.x, w, acc[4]); acc[5] = fmaf(b.y, w, acc[5]); acc[6] = fmaf(b.z, w, acc[6]); acc[7] = fmaf(b.w, w, acc[7]);
                acc[8] = fmaf(cc.x, w, acc[8]); acc[9] = fmaf(cc.y, w, acc[9]); acc[10] = fmaf(cc.z, w, acc[10]); acc[11] = fmaf(cc.w, w, acc[11]);
            } else if (FMT == 4) {
                float4 a = __ldg(reinterpret_cast<const float4*>(tab + v * 16));
                acc[0] = fmaf(a.x, w, acc[0]); acc[1] = fmaf(a.y, w, acc[1]); acc[2] = fmaf(a.z, w, acc[2]); acc[3] = fmaf(a.w, w, acc[3]);
            } else {
                float r[8], q[8];
                ld256(tab + v * 64, r);
                ld256(tab + v * 64 + 32, q);
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] = fmaf(r[k], w, acc[k]);
#pragma unroll
                for (int k = 0; k < 4; k++) acc[8 + k] = fmaf(q[k], w, acc[8 + k]);
            }
        }
        dep = acc[0] + acc[5];
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 12; k++) s += acc[k];
    if (s == 123.456f) out[0] = s;
}

template <int FMT>
static void run(const unsigned char* tab, int smem_kb, int spread, float* out, cudaTextureObject_t tex) {
    const int iters = 2000;
    cudaFuncSetAttribute(k_gather<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<FMT><<<148, 512, smem_kb * 1024>>>(tab, 200, spread, out, tex);
    cudaEventRecord(e0);
    k_gather<FMT><<<148, 512, smem_kb * 1024>>>(tab, iters, spread, out, tex);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fetches = 148.0 * 512 * iters;
    printf("fmt %d smem %3d KB spread %3d : %7.2f ms  %6.2f Gfetch/s  %s\n", FMT, smem_kb, spread, ms, fetches / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t nvox = (size_t)D_ * H_ * W_;
    unsigned char* tab;
    cudaMalloc(&tab, nvox * 64);
    cudaMemset(tab, 0, nvox * 64);
    float* out;
    cudaMalloc(&out, 4);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = tab;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>();
    rd.res.linear.sizeInBytes = nvox * 48;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    cudaError_t te = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    printf("texture object: %s\n", cudaGetErrorString(te));
    const int smems[1] = {48};
    const int spreads[4] = {2, 4, 8, 128};
    for (int si = 0; si < 1; si++)
        for (int sp = 0; sp < 4; sp++) {
            run<0>(tab, smems[si], spreads[sp], out, tex);
            run<1>(tab, smems[si], spreads[sp], out, tex);
            run<4>(tab, smems[si], spreads[sp], out, tex);
            run<6>(tab, smems[si], spreads[sp], out, tex);
            run<7>(tab, smems[si], spreads[sp], out, tex);
            run<8>(tab, smems[si], spreads[sp], out, tex);
        }
    return 0;
}
