// Subject set-up on the device: what the reference does with torch ops / pytorch3d on the GPU once per
// subject and once per frame before its render kernels run.
//
// Replaces (reference file:line):
//   lbs(): blend_shapes, vertices2joints, batch_rodrigues, pose blend shapes, batch_rigid_transform, skinning
//                                                  models/deformers/smplx/lbs.py:152-248, 251-272, 275-296, 345-401
//   SMPL.forward: translation of vertices / joints / A     models/deformers/smplx/body_models.py:342-358
//   query_weights_smpl: K = 30 nearest vertices, inverse-distance weights, 30 smoothing passes
//                                                  models/deformers/fast_snarf/deformer_torch.py:234-253
//   knn_points (K nearest neighbours, brute force)  lib/pytorch3d/cuda/knn.cu:27-312
//   switch_to_explicit: grid, bbox, offset / scale kernels   deformer_torch.py:139-167
#pragma once

#define IA_KNN_K 30
#define IA_KNN_TILE 1024

// ---------------------------------------------------------------------------------------------------------------------
// SMPL linear blend skinning, batch 1.  All arrays fp32 on the device.
__global__ void k_lbs_shape(const float* __restrict__ v_template, const float* __restrict__ shapedirs, const float* __restrict__ betas,
                            int V, int NB, float* __restrict__ v_shaped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // vertex * 3 + component
    if (i >= V * 3) return;
    float s = v_template[i];
    for (int b = 0; b < NB; b++) s = fmaf(shapedirs[(size_t)i * NB + b], betas[b], s);   // blend_shapes: einsum('bl,mkl->bmk')
    v_shaped[i] = s;
}

// J = J_regressor v_shaped (vertices2joints): one block per joint
__global__ void __launch_bounds__(256) k_lbs_joints(const float* __restrict__ J_regressor, const float* __restrict__ v_shaped, int V,
                                                    float* __restrict__ joints) {
    __shared__ float red[3][8];
    const int j = blockIdx.x;
    float a[3] = {0.f, 0.f, 0.f};
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const float w = J_regressor[(size_t)j * V + v];
        a[0] = fmaf(w, v_shaped[v * 3], a[0]); a[1] = fmaf(w, v_shaped[v * 3 + 1], a[1]); a[2] = fmaf(w, v_shaped[v * 3 + 2], a[2]);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int w = 0; w < 8; w++) s += red[threadIdx.x][w];
        joints[j * 3 + threadIdx.x] = s;
    }
}

// batch_rodrigues (lbs.py:275-296) + batch_rigid_transform (:345-401) for the 24 joints: one warp, lane = joint; the
// chain itself is walked by lane 0 (24 dependent 4x4 products).  pose [72] axis-angle (global_orient, body_pose).
// Outputs: pose_feature [207] = (R[1:] - I), posed joints [24][3], A [24][16] (relative transforms), both with transl added
// the way SMPL.forward does (body_models.py:342-358).
__global__ void __launch_bounds__(32) k_lbs_chain(const float* __restrict__ pose, const float* __restrict__ joints_rest,
                                                  const int* __restrict__ parents, const float* __restrict__ transl,
                                                  float* __restrict__ pose_feature, float* __restrict__ joints_out,
                                                  float* __restrict__ A_out, float* __restrict__ A_rel) {
    __shared__ float R[24][9];
    __shared__ float G[24][16];
    const int j = threadIdx.x;
    if (j < 24) {
        // batch_rodrigues: angle = |r + 1e-8|, axis = r / angle, R = I + sin K + (1 - cos) K K
        const float rx = pose[j * 3], ry = pose[j * 3 + 1], rz = pose[j * 3 + 2];
        const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
        const float angle = sqrtf(ax * ax + ay * ay + az * az);
        const float x = rx / angle, y = ry / angle, z = rz / angle;
        const float s = sinf(angle), c = cosf(angle);
        const float K[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
        float KK[9];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) KK[a * 3 + b] = K[a * 3] * K[b] + K[a * 3 + 1] * K[3 + b] + K[a * 3 + 2] * K[6 + b];
#pragma unroll
        for (int k = 0; k < 9; k++) R[j][k] = ((k % 4 == 0) ? 1.f : 0.f) + s * K[k] + (1.f - c) * KK[k];
        if (j > 0)
#pragma unroll
            for (int k = 0; k < 9; k++) pose_feature[(j - 1) * 9 + k] = R[j][k] - ((k % 4 == 0) ? 1.f : 0.f);
    }
    __syncwarp();
    if (j == 0) {
        for (int i = 0; i < 24; i++) {
            const int pa = parents[i];
            float rel[3];
            for (int k = 0; k < 3; k++) rel[k] = joints_rest[i * 3 + k] - (i > 0 ? joints_rest[pa * 3 + k] : 0.f);
            float T[16] = {R[i][0], R[i][1], R[i][2], rel[0], R[i][3], R[i][4], R[i][5], rel[1], R[i][6], R[i][7], R[i][8], rel[2],
                           0.f, 0.f, 0.f, 1.f};
            if (i == 0) {
                for (int k = 0; k < 16; k++) G[0][k] = T[k];
            } else {
                for (int a = 0; a < 4; a++)
                    for (int b = 0; b < 4; b++) {
                        float s = 0.f;
                        for (int k = 0; k < 4; k++) s += G[pa][a * 4 + k] * T[k * 4 + b];
                        G[i][a * 4 + b] = s;
                    }
            }
        }
    }
    __syncwarp();
    if (j < 24) {
        // posed joint = last column; rel_transform = G - pad(G [joint; 0])
        float jh[3] = {joints_rest[j * 3], joints_rest[j * 3 + 1], joints_rest[j * 3 + 2]};
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const float corr = G[j][a * 4] * jh[0] + G[j][a * 4 + 1] * jh[1] + G[j][a * 4 + 2] * jh[2];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                float v = G[j][a * 4 + b] - (b == 3 ? corr : 0.f);
                A_rel[j * 16 + a * 4 + b] = v;                       // lbs(): skinning uses the untranslated transforms
                if (b == 3 && a < 3) v += transl[a];
                A_out[j * 16 + a * 4 + b] = v;
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) joints_out[j * 3 + k] = G[j][k * 4 + 3] + transl[k];
    }
}

// pose blend shapes + skinning (lbs.py:222-245) with the untranslated transforms, then + transl (body_models.py:354): one
// thread per vertex
__global__ void k_lbs_skin(const float* __restrict__ v_shaped, const float* __restrict__ posedirs, const float* __restrict__ pose_feature,
                           const float* __restrict__ lbs_weights, const float* __restrict__ A_rel, const float* __restrict__ transl, int V,
                           float* __restrict__ vertices) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float vp[3] = {v_shaped[v * 3], v_shaped[v * 3 + 1], v_shaped[v * 3 + 2]};
    for (int k = 0; k < 207; k++) {                       // posedirs [207][V * 3]
        const float f = pose_feature[k];
        const float* pd = posedirs + (size_t)k * V * 3 + v * 3;
        vp[0] = fmaf(f, pd[0], vp[0]); vp[1] = fmaf(f, pd[1], vp[1]); vp[2] = fmaf(f, pd[2], vp[2]);
    }
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; k++) T[k] = 0.f;
    for (int j = 0; j < 24; j++) {
        const float w = lbs_weights[(size_t)v * 24 + j];
#pragma unroll
        for (int k = 0; k < 12; k++) T[k] = fmaf(w, A_rel[j * 16 + k], T[k]);
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
        vertices[v * 3 + a] = (T[a * 4] * vp[0] + T[a * 4 + 1] * vp[1] + T[a * 4 + 2] * vp[2] + T[a * 4 + 3]) + transl[a];
}

// ---------------------------------------------------------------------------------------------------------------------
// Skinning-weight voxelisation.
// bbox of the vertices -> offset (centre), scale (max half extent * 1.2)   (deformer_torch.py:150-152)
__global__ void __launch_bounds__(256) k_vox_bbox(const float* __restrict__ verts, int V, float* __restrict__ mnmx) {
    __shared__ float smn[3][8], smx[3][8];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int v = threadIdx.x; v < V; v += blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], verts[v * 3 + k]); mx[k] = fmaxf(mx[k], verts[v * 3 + k]); }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if ((threadIdx.x & 31) == 0) { smn[k][threadIdx.x >> 5] = mn[k]; smx[k][threadIdx.x >> 5] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = INFINITY, b = -INFINITY;
        for (int w = 0; w < 8; w++) { a = fminf(a, smn[threadIdx.x][w]); b = fmaxf(b, smx[threadIdx.x][w]); }
        mnmx[threadIdx.x] = a; mnmx[3 + threadIdx.x] = b;
    }
}

// K nearest vertices of every grid point (brute force over vertex tiles staged in shared memory; the K best are kept
// sorted by insertion), inverse-distance blend of their skinning weights -> vox [24][P] (query_weights_smpl :236-243).
__global__ void __launch_bounds__(128) k_vox_knn(const float* __restrict__ verts, const float* __restrict__ weights, int V, int D, int H,
                                                 int W, float off0, float off1, float off2, float scale, float ratio,
                                                 float* __restrict__ vox) {
    __shared__ float sv[IA_KNN_TILE * 3];
    const int P = D * H * W;
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    float g[3] = {0.f, 0.f, 0.f};
    if (pidx < P) {
        const int xi = pidx % W, yi = (pidx / W) % H, zi = pidx / (W * H);
        // torch.linspace(-1, 1, n)[i], then denormalize: z / ratio, * scale, + offset  (deformer_torch.py:145-148, 176-181)
        const float lx = W > 1 ? -1.f + 2.f * (float)xi / (float)(W - 1) : -1.f;
        const float ly = H > 1 ? -1.f + 2.f * (float)yi / (float)(H - 1) : -1.f;
        const float lz = D > 1 ? -1.f + 2.f * (float)zi / (float)(D - 1) : -1.f;
        g[0] = lx * scale + off0; g[1] = ly * scale + off1; g[2] = (lz / ratio) * scale + off2;
    }
    float bd[IA_KNN_K];
    int bi[IA_KNN_K];
#pragma unroll 1
    for (int k = 0; k < IA_KNN_K; k++) { bd[k] = INFINITY; bi[k] = 0; }
    for (int t0 = 0; t0 < V; t0 += IA_KNN_TILE) {
        const int nt = min(IA_KNN_TILE, V - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt * 3; i += blockDim.x) sv[i] = verts[(size_t)t0 * 3 + i];
        __syncthreads();
        if (pidx < P) {
            for (int i = 0; i < nt; i++) {
                const float e0 = g[0] - sv[i * 3], e1 = g[1] - sv[i * 3 + 1], e2 = g[2] - sv[i * 3 + 2];
                const float d2 = e0 * e0 + e1 * e1 + e2 * e2;
                if (d2 < bd[IA_KNN_K - 1]) {
                    int k = IA_KNN_K - 1;
#pragma unroll 1
                    while (k > 0 && bd[k - 1] > d2) { bd[k] = bd[k - 1]; bi[k] = bi[k - 1]; k--; }
                    bd[k] = d2; bi[k] = t0 + i;
                }
            }
        }
    }
    if (pidx >= P) return;
    const int K = min(IA_KNN_K, V);
    float ws[IA_KNN_K], wsum = 0.f;
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        const float d = fminf(fmaxf(sqrtf(bd[k]), 0.0001f), 1.0f);   // dist.sqrt().clamp_(0.0001, 1.)
        ws[k] = 1.0f / d;
        wsum += ws[k];
    }
    for (int c = 0; c < 24; c++) {
        float s = 0.f;
#pragma unroll 1
        for (int k = 0; k < K; k++) s += (ws[k] / wsum) * weights[(size_t)bi[k] * 24 + c];
        vox[(size_t)c * P + pidx] = s;
    }
}

// one smoothing pass (deformer_torch.py:246-252): interior voxels move 30 % towards the mean of their six neighbours
// (computed from the OLD values), then every voxel is renormalised over the 24 channels
__global__ void k_vox_smooth(const float* __restrict__ src, float* __restrict__ dst, int D, int H, int W) {
    const int P = D * H * W;
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= P) return;
    const int xi = pidx % W, yi = (pidx / W) % H, zi = pidx / (W * H);
    const bool interior = xi > 0 && xi < W - 1 && yi > 0 && yi < H - 1 && zi > 0 && zi < D - 1;
    float v[24], sum = 0.f;
    for (int c = 0; c < 24; c++) {
        const float* s = src + (size_t)c * P;
        float x = s[pidx];
        if (interior) {
            const float mean = (s[pidx + W * H] + s[pidx - W * H] + s[pidx + W] + s[pidx - W] + s[pidx + 1] + s[pidx - 1]) / 6.0f;
            x = (x - mean) * 0.7f + mean;
        }
        v[c] = x;
        sum += x;
    }
    for (int c = 0; c < 24; c++) dst[(size_t)c * P + pidx] = v[c] / sum;
}
