// Warp-level tensor-core evaluation of the small dense layers on the path (models/network_utils.py:201-244 VanillaMLP,
// :360-428 LipshitzMLP; call sites models/rf/geometry.py:147-172, models/rf/radiance.py:111-135, models/pbr/material.py:31-51).
//
// Why: the 16-lane team evaluation (ia_team_geometry / ia_team_dense64) broadcasts every input with a shuffle and re-reads
// every weight row from shared memory for every point: per geometry evaluation 70 LDS.128 + 32 SHFL per lane, all on the
// same LSU / shared-memory pipe the voxel and hash-grid gathers of the wavefront kernel are bound by.  Here a warp batches
// 16 points as the M dimension of mma.sync.m16n8k8 (TF32 inputs, fp32 accumulate): the weights are read once per 16 points
// as pre-split B fragments, the inputs once as A fragments.
//
// Accuracy: fp32 operands are split a = a_hi + a_lo (a_hi = tf32(a), a_lo = tf32(a - a_hi)) and each product is evaluated as
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi ("3xTF32"): the dropped term a_lo*b_lo and the rounding of the lo parts are ~2^-22
// relative, i.e. the layer output differs from an fp32 FMA chain by a few fp32 ulps of its largest addend -- the same size
// as the difference between two fp32 summation orders (measured against the fp32 team path in tests/test_gpu_ops.py).
//
// Fragment layouts of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 (g = lane >> 2, t = lane & 3):
//   A [16 x 8]: a0 = (g, t)  a1 = (g + 8, t)  a2 = (g, t + 4)  a3 = (g + 8, t + 4)
//   B [ 8 x 8]: b0 = (k = t, n = g)  b1 = (k = t + 4, n = g)
//   C [16 x 8]: c0 = (g, 2t)  c1 = (g, 2t + 1)  c2 = (g + 8, 2t)  c3 = (g + 8, 2t + 1)
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t ia_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
// x = hi + lo with hi = x rounded to TF32 (10 mantissa bits, ties away from zero -- what cvt.rna.tf32.f32 computes, without
// its Inf / NaN guard: three extra instructions per conversion on sm_100a) and lo = x - hi (exact in fp32) TRUNCATED to
// TF32: a relative error of 2^-21 of x at most.
__device__ __forceinline__ void ia_split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}
__device__ __forceinline__ void ia_mma_tf32(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += A B with both operands split (small terms first)
__device__ __forceinline__ void ia_mma_3xtf32(float c[4], const uint32_t ahi[4], const uint32_t alo[4], const float4 bf) {
    const uint32_t b0h = __float_as_uint(bf.x), b0l = __float_as_uint(bf.y), b1h = __float_as_uint(bf.z), b1l = __float_as_uint(bf.w);
    ia_mma_tf32(c, alo, b0h, b1h);
    ia_mma_tf32(c, ahi, b0l, b1l);
    ia_mma_tf32(c, ahi, b0h, b1h);
}

// Pre-split B fragments of one dense layer y[n] = sum_k W(k, n) x[k]:
//   dst[(s * n_tiles + nt) * 32 + lane] = {b0_hi, b0_lo, b1_hi, b1_lo},  b0 = W(8s + t, 8nt + g), b1 = W(8s + t + 4, 8nt + g)
// `W(k, n)` is any callable (k in [0, 8 * k_steps), n in [0, 8 * n_tiles)) returning 0 for padding.  All threads of the CTA.
template <class WFn>
__device__ __forceinline__ void ia_stage_bfrag(float4* dst, int k_steps, int n_tiles, WFn W) {
    for (int i = threadIdx.x; i < k_steps * n_tiles * 32; i += blockDim.x) {
        const int lane = i & 31, nt = (i >> 5) % n_tiles, s = (i >> 5) / n_tiles;
        const int g = lane >> 2, t = lane & 3;
        uint32_t h0, l0, h1, l1;
        ia_split_tf32(W(8 * s + t, 8 * nt + g), h0, l0);
        ia_split_tf32(W(8 * s + t + 4, 8 * nt + g), h1, l1);
        dst[i] = make_float4(__uint_as_float(h0), __uint_as_float(l0), __uint_as_float(h1), __uint_as_float(l1));
    }
}

// One dense layer for a 16-row tile whose inputs sit in shared memory, row-major with leading dimension `ld` floats
// (ld % 32 == 12 or 4 keeps the A-fragment loads conflict-free): c[nt][.] += X[16 x 8 K_STEPS] W.
// N_TILES consecutive n-tiles of a layer staged with N_TILES_TOTAL per k-step (wfrag points at the first of them).
// ROWS = 8: the tile holds rows 0..7 only (rows 8..15 of the m16 operand are zero, c[.][2..3] stay at their initial value).
template <int K_STEPS, int N_TILES, int N_TILES_TOTAL = N_TILES, int ROWS = 16, class BF = float4>
__device__ __forceinline__ void ia_mma_layer_smem(const float* __restrict__ xs, int ld, const BF* __restrict__ wfrag,
                                                  float c[N_TILES][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int s = 0; s < K_STEPS; s++) {
        uint32_t ahi[4], alo[4];
        ia_split_tf32(xs[g * ld + 8 * s + t], ahi[0], alo[0]);
        ia_split_tf32(xs[g * ld + 8 * s + t + 4], ahi[2], alo[2]);
        if (ROWS == 16) {
            ia_split_tf32(xs[(g + 8) * ld + 8 * s + t], ahi[1], alo[1]);
            ia_split_tf32(xs[(g + 8) * ld + 8 * s + t + 4], ahi[3], alo[3]);
        } else {
            ahi[1] = alo[1] = ahi[3] = alo[3] = 0u;
        }
#pragma unroll
        for (int nt = 0; nt < N_TILES; nt++) ia_mma_3xtf32(c[nt], ahi, alo, wfrag[(s * N_TILES_TOTAL + nt) * 32 + lane]);
    }
}

// The next layer straight from the accumulators of the previous one (no shared-memory round trip): activation h = act(c)
// of a 64-wide layer is in C layout -- lane (g, t) holds columns 8nt + 2t, 8nt + 2t + 1 of rows g, g + 8 -- and is fed as
// the A operand of k-step nt with the k order PERMUTED inside the step: slot t <- column 2t, slot t + 4 <- column 2t + 1.
// The B fragments of that layer must be staged with the same permutation: ia_kperm(k) below.
__host__ __device__ __forceinline__ int ia_kperm(int k) {   // k-slot -> input column of a layer fed from C fragments
    const int s = k >> 3, j = k & 7;
    return 8 * s + (j < 4 ? 2 * j : 2 * (j - 4) + 1);
}
template <int N_TILES_OUT, class BF = float4>
__device__ __forceinline__ void ia_mma_layer_regs(const float h[8][4], const BF* __restrict__ wfrag, float c[N_TILES_OUT][4]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        uint32_t ahi[4], alo[4];
        ia_split_tf32(h[s][0], ahi[0], alo[0]);   // (g, slot t)      <- c0 = (g, 2t)
        ia_split_tf32(h[s][2], ahi[1], alo[1]);   // (g + 8, slot t)  <- c2
        ia_split_tf32(h[s][1], ahi[2], alo[2]);   // (g, slot t + 4)  <- c1 = (g, 2t + 1)
        ia_split_tf32(h[s][3], ahi[3], alo[3]);   // (g + 8, t + 4)   <- c3
#pragma unroll
        for (int nt = 0; nt < N_TILES_OUT; nt++) ia_mma_3xtf32(c[nt], ahi, alo, wfrag[(s * N_TILES_OUT + nt) * 32 + lane]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The radiance network (models/rf/radiance.py:111-135: 67 -> 64 -> 64 -> 3, ReLU, sigmoid) and the material network
// (models/pbr/material.py:31-51: 48 -> 64 -> 64 -> 5) for 16 points whose inputs sit in a shared-memory "shading tile",
// row-major with leading dimension IA_SHADE_LD, columns
//   0..31 hash features | 32..34 the scaled position 2 xn - 1 | 35..47 geometry feature | 48..63 SH(reflected dir) |
//   64..66 world normal | 67 unused
// (the reference's radiance input order is [xyz, hash, feature, SH, normal]: ia_rad_in_of maps a tile column to it; the
// first 48 columns are the material network's input, cat[xyz_embd 35, feature 13]).  The radiance layer 1 runs 9 k-steps
// = 72 columns: columns 67..71 carry zero weights and alias column 67 and the first four columns of the NEXT row (the
// four padding floats of IA_SHADE_TILE for the last row), so every float of the tile must be finite.
#define IA_SHADE_LD 68
// floats per warp tile: 16 rows + 4 floats that stay zero -- the last k-step of row 15 reads them, so that no warp ever
// reads another warp's tile (rows 0..14 read the first four columns of the next row of the SAME tile)
#define IA_SHADE_TILE (16 * IA_SHADE_LD + 4)
__host__ __device__ __forceinline__ int ia_rad_in_of(int k) {    // tile column -> input index of the radiance net (-1: padding)
    return k < 32 ? 3 + k : (k < 35 ? k - 32 : (k < 67 ? k : -1));
}
__host__ __device__ __forceinline__ int ia_mat_in_of(int k) {    // tile column -> input index of the material net (48 inputs)
    return k < 32 ? 3 + k : (k < 35 ? k - 32 : (k < 48 ? k : -1));
}

// `wf`: pre-split fragments of the three layers, contiguous (layer 1 | layer 2 | layer 3: IA_FRAG_RAD1.. / IA_FRAG_MAT1..);
// b1, b2, b3: the biases.  Every lane returns the N_OUT pre-activation outputs of row (lane & 15).
template <int K1_STEPS, int N_OUT>
__device__ __forceinline__ void ia_warp_mlp3(const float* __restrict__ xs, int ld, const float4* __restrict__ wf,
                                             const float* __restrict__ b1, const float* __restrict__ b2,
                                             const float* __restrict__ b3, float out[N_OUT]) {
    const int lane = threadIdx.x & 31, t = lane & 3;
    float c[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(b1 + 8 * nt + 2 * t));
        c[nt][0] = b.x; c[nt][1] = b.y; c[nt][2] = b.x; c[nt][3] = b.y;
    }
    ia_mma_layer_smem<K1_STEPS, 8, 8, 16>(xs, ld, wf, c);
    float h[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(b2 + 8 * nt + 2 * t));
#pragma unroll
        for (int k = 0; k < 4; k++) h[nt][k] = fmaxf(c[nt][k], 0.f);
        c[nt][0] = b.x; c[nt][1] = b.y; c[nt][2] = b.x; c[nt][3] = b.y;
    }
    ia_mma_layer_regs<8>(h, wf + K1_STEPS * 8 * 32, c);
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
#pragma unroll
        for (int k = 0; k < 4; k++) h[nt][k] = fmaxf(c[nt][k], 0.f);
    float o[1][4];
    {
        const float bx = 2 * t < N_OUT ? __ldg(b3 + 2 * t) : 0.f, by = 2 * t + 1 < N_OUT ? __ldg(b3 + 2 * t + 1) : 0.f;
        o[0][0] = bx; o[0][1] = by; o[0][2] = bx; o[0][3] = by;
    }
    ia_mma_layer_regs<1>(h, wf + (K1_STEPS * 8 + 64) * 32, o);
    // lane (g, t) holds outputs 2t, 2t + 1 of rows g (o[0][0..1]) and g + 8 (o[0][2..3]); row j wants them from lanes 4 (j & 7) + (k >> 1)
#pragma unroll
    for (int k = 0; k < N_OUT; k++) {
        const int src = 4 * (lane & 7) + (k >> 1);
        const float vlo = __shfl_sync(0xffffffffu, (k & 1) ? o[0][1] : o[0][0], src);
        const float vhi = __shfl_sync(0xffffffffu, (k & 1) ? o[0][3] : o[0][2], src);
        out[k] = (lane & 8) ? vhi : vlo;
    }
}
