// Wavefront secondary-ray integrator (second kernel generation of the shading stage).
//
// Why: the first generation ran one 16-lane team per secondary ray; ncu showed 10.5 of 32 lanes active
// per instruction (the 13 Broyden chains of a query run in lock-step until the slowest ends) and 45 %
// instruction-fetch stalls (profiles/r1_k_shade_team_summary.md).  Here one persistent CTA per SM keeps
// WF_R rays in flight with their state in shared memory and advances all of them one SDF query per
// round; every round is a sequence of CTA-wide phases that are each dense and convergent:
//
//   feed     : a tile of shading samples -> light pick, cosine test -> ring of live secondary rays
//   advance  : one thread per ray slot: consume the SDF of the previous query, step the lazy state
//              machine of compute_indirect_radiance (grid march -> first +/- crossing -> CDF walk of
//              ray_resampling_sdf_fine -> 4 fine intervals), emit the next query point; a finished ray
//              is shaded (BRDF x light / pdf, accumulated into its pixel) and the slot re-filled
//   broyden  : the 13 x n_q Broyden chains as independent tasks, grabbed dynamically, ONE voxel fetch
//              per loop trip for every lane (a lane whose chain ended starts the next task in the same trip)
//   filter   : duplicate-root removal per query -> compact list of (query, root) geometry tasks
//   geometry : 16-lane teams evaluate hash grid + MLP for the listed roots only
//
// Same arithmetic, in the same order, as the reference (models/intrinsic_avatar.py:396-545; cdf.cu:536-638;
// fuse_cuda_kernel_fast.cu:250-413; filter.cu:10-54); only the accumulation order into a pixel differs.
#pragma once

#ifndef WF_THREADS
#define WF_THREADS 512
#endif
#ifndef WF_CTAS_PER_SM
#define WF_CTAS_PER_SM 1   // resident CTAs per SM (WF_THREADS * WF_CTAS_PER_SM * regs <= 64K)
#endif
#ifndef WF_R
#define WF_R 1024      // ray slots per CTA (two per thread): longer phases amortise the barriers and phase tails.
#endif                 // Measured at 512^2 x 1024 spp: R = 512 775 ms, 1024 743 ms, 1536 752 ms, 2048 751 ms, 4096 804 ms;
                       // 2 CTAs x 256 threads per SM: R = 256 773 ms, 512 756 ms, 1024 751 ms (no gain from co-resident CTAs)
#ifndef WF_FEED
#define WF_FEED 1024   // items examined per feed step
#endif
// ring capacity: a power of two >= WF_R + WF_FEED (the ring is fed while it holds fewer than WF_R entries)
#if WF_R + WF_FEED <= 2048
#define WF_QCAP 2048
#elif WF_R + WF_FEED <= 4096
#define WF_QCAP 4096
#else
#define WF_QCAP 8192
#endif
#define WF_NST 36      // state words per ray
// prune phase: 32-pair trips per warp (its share of the 13 x WF_R pairs, rounded up to 32)
#define WF_PRUNE_ITERS (((WF_R * 13 + (WF_THREADS / 32) - 1) / (WF_THREADS / 32) + 31) / 32)
#ifndef WF_LIGHT_MAJOR
#define WF_LIGHT_MAJOR 0   // light-table modes: shading samples are stored and fed light-major -- for each light direction all
#endif                     // hit pixels in ray order (k_resample) -- so that the rays in flight are bundles of parallel rays from
                           // neighbouring surface points and, with WF_QSORT, neighbouring lanes gather the same voxel / hash cells.
                           // 0: pixel-major (all samples of a pixel, then the next pixel).  Only the processing order differs.
                           // Measured at 512^2 x 1024 spp, GI on, frames 2-4: 1036 ms pixel-major, 1066 ms (+ 6 ms of scattered writes in
                           // k_resample) light-major: a voxel cell is shared by 2-4 live rays either way, and the gathers only get
                           // cheaper from ~8 lanes per cell on (scripts/gather_microbench.cu).
#ifndef WF_QSORT
#define WF_QSORT 1     // 1: the pending queries of a round are sorted along a Morton curve over their posed position before the
#endif                 // task lists are built: neighbouring lanes of the Broyden and geometry phases then touch the same voxel
                       // cells / hash cells (scripts/gather_microbench.cu: a warp whose lanes stay within 2 cells gathers 1.5x, within
                       // one cell 4.7x faster than a warp of scattered lanes).  Only the processing order changes.
#ifndef WF_STATE_GLOBAL
#define WF_STATE_GLOBAL 1
#endif
#ifndef WF_DEFER_FINISH
#define WF_DEFER_FINISH 1   // the estimator of a finished ray (BRDF, 6 atomics) is not evaluated in the refill pass, which the whole
#endif                      // CTA waits for, but listed and evaluated in front of the dynamically scheduled Broyden phase
#ifndef WF_BTASK_SMEM
#define WF_BTASK_SMEM 0   // 1: Broyden task list in shared memory, so that a chain start reads its task id with shared-memory
#endif                    // latency (from the global scratch that read is an exposed L2 round trip: 6 % of the Broyden phase's
                          // stall samples).  Measured: 731 -> 746 ms -- the 26 KB taken from L1 cost more than the latency saved.
// per-CTA scratch: roots [R][13][3] f32, their SDFs [R][13] f32, 2 task lists [R*13] u16, GI task list [R] uint2, ray state
#define WF_OFF_CSDF (WF_R * IA_N_INIT * 3 * 4)
#define WF_OFF_GTASK (WF_OFF_CSDF + WF_R * IA_N_INIT * 4)
#define WF_OFF_BTASK (WF_OFF_GTASK + WF_R * IA_N_INIT * 2)
#define WF_OFF_GITASK (WF_OFF_BTASK + WF_R * IA_N_INIT * 2)
#define WF_OFF_GIXC (WF_OFF_GITASK + WF_R * 8)
#define WF_FCAP (2 * WF_R)     // finished rays whose estimator is deferred to the front of the Broyden phase (per round)
#define WF_OFF_FLIST (WF_OFF_GIXC + WF_R * 3 * 4)
#define WF_OFF_STATE (WF_OFF_FLIST + WF_FCAP * 9 * 4)
#define WF_SCRATCH_BYTES (WF_OFF_STATE + WF_NST * WF_R * 4)

enum { WF_C_Q = 0, WF_C_FETCH, WF_C_GEO, WF_C_RAYS, WF_C_SKIP, WF_C_QG, WF_C_RAD };
enum { WF_IDLE = 0, WF_FIRST = 1, WF_SEARCH = 2, WF_CDF = 3, WF_FINE = 4, WF_GIWAIT = 5,
       WF_DONE = 6, WF_DONE_IND = 7 };   // ray ended (6: without, 7: with indirect radiance in WS_IND); estimator pending
enum { WF_ACT_NEXT = 0, WF_ACT_CDF = 1, WF_ACT_FINE_START = 2, WF_ACT_FINISH = 3 };

// state word indices
#define WS_O 0        // 0..2  origin
#define WS_D 3        // 3..5  direction
#define WS_ID 6       // item id (uint32)
#define WS_PACK 7     // stage(3) | j(3) << 3 | i(3) << 6 | aux(16) << 16
#define WS_TLAST 8
#define WS_TMAX 9
#define WS_TDIST 10   // 10..12
#define WS_DELTA 13   // 13..15
#define WS_CELL 16    // cur(3x6) | (step+1)(3x2) << 18 | continuous << 24 | done << 25
#define WS_OVER 17    // (over+1)(3x7)
#define WS_SDFPREV 18
#define WS_CS 19
#define WS_CE 20
#define WS_TS 21
#define WS_TE 22
#define WS_TRANS 23   // trans (CDF) / Tacc (fine)
#define WS_CDFPREV 24 // cdf_prev (CDF) / acc (fine)
#define WS_CDFNEXT 25
#define WS_CDFU 26
#define WS_TPL 27     // 27..31
#define WS_IND 32      // 32..34 indirect radiance accumulated over the fine samples (global illumination)

// What the tensor-core paths read of the fp32 geometry weights, staged compactly in shared memory (the 35 -> 64 layer and the
// rows 1..12 of the 64 -> 13 layer are read as B fragments): 576 bytes instead of the 12.6 KB blob -- the rest is L1.
#define IA_GEOC_B1 0      // [64] bias of the 35 -> 64 layer
#define IA_GEOC_W2 64     // [64] row 0 of the 64 -> 13 layer (the sdf)
#define IA_GEOC_B2 128    // [16] bias of the 64 -> 13 layer (13 used)
#define IA_GEOC_END 144
__device__ __forceinline__ void ia_stage_geoc(float* __restrict__ dst, const float* __restrict__ mlp) {
    for (int i = threadIdx.x; i < IA_GEOC_END; i += blockDim.x)
        dst[i] = i < 64 ? mlp[IA_GEO_B1 + i] : (i < 128 ? mlp[IA_GEO_W2 + i - 64] : mlp[IA_GEO_B2 + i - 128]);
}
#define IA_GEO_KSTEPS 5   // geometry tile: [16][IA_GEO_LD], 40 columns read (layout: ia_warp_geometry)
#define IA_GEO_LD 44
// per-warp input tiles of the tensor-core phases: [16][IA_GEO_LD] for the geometry phase; with global illumination the same
// memory is the [16][IA_SHADE_LD] shading tile of the GI phase (+ 4 floats: its last k-step reads 16 bytes past the end)
// SW = number of warps that own a shading tile: WF_GI_WARPS in the shading stage's GI phase (it runs next to the Broyden phase
// and a round has <= 16 batches: half the warps are enough, and every KB not spent here is L1 for the gathers), all of them
// in k_prim_shade_wf (0 = geometry tiles only).
#ifndef WF_GI_WARPS
#define WF_GI_WARPS 8
#endif
#define WF_TILE_FLOATS_SW(SW) ((SW) * IA_SHADE_TILE > (WF_THREADS / 32) * 16 * IA_GEO_LD ? (SW) * IA_SHADE_TILE : (WF_THREADS / 32) * 16 * IA_GEO_LD)
#define WF_TILE_FLOATS(GI) WF_TILE_FLOATS_SW((GI) ? WF_GI_WARPS : 0)
struct WfShared {
    float* w;               // MLP weights staged behind this struct (geometry only, or geometry + radiance for GI)
    float tfs13[IA_N_INIT * 12];
    IaLevel lvl[IA_N_LEVELS];
#if WF_STATE_GLOBAL
    float (*st)[WF_R];      // [WF_NST][WF_R] ray state, in the CTA's global scratch (read/written once per round)
#else
    float st[WF_NST][WF_R];
#endif
    float qx[3][WF_R];
    unsigned int qmask[WF_R];
    unsigned short qlist[WF_R];
    unsigned pballot[(WF_THREADS / 32) * WF_PRUNE_ITERS];   // prune phase: live ballots of every warp's pair range
    unsigned short rlist[WF_R];   // advance phase: slots to finish and / or refill
    int n_rlist;
#if WF_QSORT
    unsigned int sortk[WF_R];     // (Morton code of the query position << 10) | slot
#endif
    uint2 ring[WF_QCAP];
    // tensor-core geometry phase (ia_mma.cuh): pre-split B fragments of the 35 -> 64 layer, one 16-point input tile per warp
    float4* w1f;            // [IA_GEO_KSTEPS][8][32]
    float* xs;              // [warps][16][IA_GEO_LD]
    // CTA-private scratch in GLOBAL memory (low traffic; keeping it out of shared memory leaves the
    // L1 carve-out to the voxel_J gathers, which is what the kernel is bound by -- DESIGN.md):
    float* cand;            // [WF_R][13][3] Broyden roots
    float* csdf;            // [WF_R][13] SDF of the kept roots
    uint2* gitask;          // [WF_R] GI: (slot | root << 16, weight bits) of the fine samples consumed this round
    float* flist;           // [WF_FCAP][9] finished rays: entry.x, entry.y, T, ind[3], d[3] (WF_DEFER_FINISH)
    int n_flist;
    float* gixc;            // [WF_R][3] their canonical roots (copied: the GI phase runs next to the Broyden phase, which
                            // overwrites the slot's roots)
    int n_gitask;
    unsigned short* gtask;  // [WF_R * 13] geometry task list
#if WF_BTASK_SMEM
    unsigned short btask[WF_R * IA_N_INIT];  // Broyden task list (pruned)
#else
    unsigned short* btask;  // [WF_R * 13] Broyden task list (pruned)
#endif
    int n_btask;
    int n_q, task_next, n_gtask, ring_head, ring_tail, tile, more_tiles, pad;
    unsigned cnt[8];        // work counters of this CTA (WF_C_*), flushed to the global counters when the kernel ends
};

// The arrays staged behind the struct, addressed from &S (an address the compiler can prove to be shared memory: loads and
// stores through the pointer FIELDS S.w / S.w1f / S.xs compile to generic LD / ST)
__device__ __forceinline__ float* wf_w(WfShared& S) {
    return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&S) + ((sizeof(WfShared) + 15) & ~(size_t)15));
}
#ifndef WF_W1F_SMEM
#define WF_W1F_SMEM 0   // B fragments of the geometry 35 -> 64 layer: 1 = staged in shared memory (20 KB), 0 = read from global memory
#endif                  // through L1 like the other layers' (IA_FRAG_GEO1).
// Shared memory is taken from L1 in steps (.. 64, 100, 132, 164 .. KB per SM): measured at 512^2 x 1024 spp, GI on, frames
// 2 / 13 / 1, shade stage: 146 KB (16 shading tiles) 795 ms -> 122 KB (8 tiles) 775 ms -> 110 KB (compact fp32 weights: same
// step, 776 ms) -> 90 KB (these fragments through L1) 764 ms.
#define WF_W1F_FLOATS (WF_W1F_SMEM ? IA_GEO_KSTEPS * 8 * 32 * 4 : 0)
__device__ __forceinline__ const float4* wf_w1f(const IaFrame& p, WfShared& S) {
#if WF_W1F_SMEM
    return reinterpret_cast<const float4*>(wf_w(S) + IA_GEOC_END);
#else
    return reinterpret_cast<const float4*>(p.mlp + IA_MLP_END) + IA_FRAG_GEO1 * 32;
#endif
}
__device__ __forceinline__ float* wf_xs(WfShared& S) { return wf_w(S) + IA_GEOC_END + WF_W1F_FLOATS; }

// ------------------------------------------------------------------------------------------------
// marcher <-> shared state
__device__ __forceinline__ void wf_store_marcher(WfShared& S, int t, const IaMarcher& m) {
    S.st[WS_TLAST][t] = m.t_last;
    S.st[WS_TMAX][t] = m.this_tmax;
#pragma unroll
    for (int k = 0; k < 3; k++) { S.st[WS_TDIST + k][t] = m.tdist[k]; S.st[WS_DELTA + k][t] = m.delta[k]; }
    unsigned a = 0, b = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a |= ((unsigned)m.cur[k] & 63u) << (6 * k);
        a |= ((unsigned)(m.step[k] + 1) & 3u) << (18 + 2 * k);
        b |= ((unsigned)(m.over[k] + 1) & 127u) << (7 * k);
    }
    a |= (m.continuous ? 1u : 0u) << 24;
    a |= (m.done ? 1u : 0u) << 25;
    S.st[WS_CELL][t] = __uint_as_float(a);
    S.st[WS_OVER][t] = __uint_as_float(b);
}
__device__ __forceinline__ void wf_load_marcher(const WfShared& S, int t, IaMarcher& m, float dt) {
    m.dt = dt;
    m.t_last = S.st[WS_TLAST][t];
    m.this_tmax = S.st[WS_TMAX][t];
#pragma unroll
    for (int k = 0; k < 3; k++) { m.tdist[k] = S.st[WS_TDIST + k][t]; m.delta[k] = S.st[WS_DELTA + k][t]; }
    unsigned a = __float_as_uint(S.st[WS_CELL][t]), b = __float_as_uint(S.st[WS_OVER][t]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        m.cur[k] = (int)((a >> (6 * k)) & 63u);
        m.step[k] = (int)((a >> (18 + 2 * k)) & 3u) - 1;
        m.over[k] = (int)((b >> (7 * k)) & 127u) - 1;
    }
    m.continuous = (a >> 24) & 1u;
    m.done = (a >> 25) & 1u;
    m.cell_loaded = false;  // t_trav / occ are recomputed from (tdist, cur): same values
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int wf_grab(WfShared& S, int n_tasks) {
    cg::coalesced_group g = cg::coalesced_threads();
    int base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(&S.task_next, (int)g.size());
    base = g.shfl(base, 0);
    int t = base + (int)g.thread_rank();
    return t < n_tasks ? t : -1;
}

// Broyden chains are taken by all lanes of the CTA from ONE list (bone-major, Morton-ordered queries).  Per-warp stretches of
// the list (a warp's chains stay spatial neighbours, other warps' stretches are taken when its own is exhausted) were measured
// at 512^2 x 1024 spp, GI on, frames 2 / 13 / 1: shade 755.0 -> 789.5 ms -- sixteen warps in sixteen places of the voxel grid
// share less of the L1 than sixteen warps sweeping one window of the list together.
__device__ __forceinline__ int wf_grab_chain(WfShared& S) { return wf_grab(S, S.n_btask); }

// A phase counts its work in a register and adds it to the CTA's counters when it ends (the counters do not live in
// registers across phases: seven of them cost the Broyden loop spills at the 128-register cap).
__device__ __forceinline__ void wf_count(WfShared& S, int which, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&S.cnt[which], v);
}

#if WF_QSORT
// qlist <- the pending queries ordered along a Morton curve (7 bits per axis over the occupancy grid's box, cells of
// ~1.6 cm: the size of a skinning-weight voxel).  Bitonic sort of WF_R keys in shared memory, < 1 % of a round.
static_assert((WF_R & (WF_R - 1)) == 0 && WF_R <= 1024, "WF_QSORT: WF_R must be a power of two <= 1024 (10-bit slot field)");
__device__ __forceinline__ unsigned wf_morton7(unsigned v) {   // 7 bits -> every third bit
    v &= 0x7fu;
    v = (v | (v << 8)) & 0x0000700fu;
    v = (v | (v << 4)) & 0x000430c3u;
    v = (v | (v << 2)) & 0x00049249u;
    return v;
}
__device__ __forceinline__ unsigned wf_morton_key(const IaFrame& p, const WfShared& S, int q) {
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float u = (S.qx[k][q] - p.aabb[k]) / (p.aabb[3 + k] - p.aabb[k]);
        const unsigned c = (unsigned)fminf(fmaxf(u * 128.0f, 0.f), 127.f);
        m |= wf_morton7(c) << k;
    }
    return m;
}
// Bitonic network with the sub-steps of distance <= 32 in registers: a warp owns 64 consecutive keys (lane: keys lane and
// lane + 32 of its block), distance 32 is a compare of the thread's own two keys, distances 16 .. 1 are shuffles; only the
// distances >= 64 go through shared memory (10 of the 55 sub-steps, 14 barriers).  Measured at 512^2 x 1024 spp, GI on,
// frames 2 / 13 / 1 together with WF_FILTER_PRELOAD: shade 765.0 -> 755.2 ms against all 55 sub-steps through shared memory.
// A counting sort on the top 10 bits of the key (4 barriers, arbitrary order inside a bucket) LOST 2.5 % (784.2 ms): the
// gathers of the following phases live on the full order.
__device__ __forceinline__ void wf_cx(unsigned& a, unsigned& b, bool up) {
    if ((a > b) == up) { const unsigned t = a; a = b; b = t; }
}
__device__ __forceinline__ void wf_sort_warp_steps(unsigned& a, unsigned& b, int ia, int k, int j_from) {
    const int lane = threadIdx.x & 31;
    if (j_from >= 32) wf_cx(a, b, (ia & k) == 0);
    for (int j = j_from >= 32 ? 16 : j_from; j > 0; j >>= 1) {
        const unsigned oa = __shfl_xor_sync(0xffffffffu, a, j), ob = __shfl_xor_sync(0xffffffffu, b, j);
        const bool lower = (lane & j) == 0;
        const bool upa = (ia & k) == 0, upb = ((ia + 32) & k) == 0;
        a = (lower == upa) ? min(a, oa) : max(a, oa);
        b = (lower == upb) ? min(b, ob) : max(b, ob);
    }
}
__device__ __forceinline__ void wf_sort_phase(const IaFrame& p, WfShared& S, int n_q) {
    static_assert(WF_R == 2 * WF_THREADS, "WF_QSORT: two keys per thread");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ia = warp * 64 + lane, ib = ia + 32;
    unsigned a = 0xffffffffu, b = 0xffffffffu;
    if (ia < n_q) { const int q = S.qlist[ia]; a = (wf_morton_key(p, S, q) << 10) | (unsigned)q; }
    if (ib < n_q) { const int q = S.qlist[ib]; b = (wf_morton_key(p, S, q) << 10) | (unsigned)q; }
    for (int k = 2; k <= 64; k <<= 1) wf_sort_warp_steps(a, b, ia, k, k >> 1);
    for (int k = 128; k <= WF_R; k <<= 1) {
        S.sortk[ia] = a; S.sortk[ib] = b;
        for (int j = k >> 1; j >= 64; j >>= 1) {
            __syncthreads();
            const int lo = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));
            const int hi = lo | j;
            const unsigned x = S.sortk[lo], y = S.sortk[hi];
            if ((x > y) == ((lo & k) == 0)) { S.sortk[lo] = y; S.sortk[hi] = x; }
        }
        __syncthreads();
        a = S.sortk[ia]; b = S.sortk[ib];
        wf_sort_warp_steps(a, b, ia, k, 32);
    }
    if (ia < n_q) S.qlist[ia] = (unsigned short)(a & 1023u);
    if (ib < n_q) S.qlist[ib] = (unsigned short)(b & 1023u);
}
#endif

// Broyden chains of all pending queries (fuse_cuda_kernel_fast.cu:250-413), one voxel fetch per trip.
// Dense pre-pass over the 13 x n_q (query, init bone) pairs: chains whose initial point has all 8 corners
// outside the voxel grid are exactly invalid (ia_all_corners_oob) and are dropped; the others are
// compacted, bone-major, into the Broyden task list.
__device__ __forceinline__ void wf_prune_phase(const IaFrame& p, WfShared& S, int n_q) {
    // Every warp owns a contiguous range of the bone-major pair list: pass 1 tests its pairs and keeps one ballot per 32 of
    // them, ONE atomic reserves the warp's stretch of the task list, pass 2 writes it.  (Round 1 took one shared atomic per
    // 32 pairs: 416 serialised atomics on one address per round.)
    unsigned c_skip = 0;
    const int n_pairs = n_q * IA_N_INIT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int per = ((n_pairs + n_warps - 1) / n_warps + 31) & ~31;      // pairs per warp, a multiple of 32
    const int k_begin = warp * per, k_end = min(n_pairs, k_begin + per);
    unsigned* ballots = S.pballot + warp * WF_PRUNE_ITERS;
    int n_live = 0;
    for (int k0 = k_begin, it = 0; k0 < k_end; k0 += 32, it++) {
        const int k = k0 + lane;
        bool live = false;
        if (k < k_end) {
            const int c = k / n_q;
            const int q = S.qlist[k - c * n_q];
            const float xd0 = S.qx[0][q], xd1 = S.qx[1][q], xd2 = S.qx[2][q];
            const float* T = S.tfs13 + c * 12;
            float d0 = xd0 - T[3], d1 = xd1 - T[7], d2 = xd2 - T[11];
            float x0 = d0 * T[0] + d1 * T[4] + d2 * T[8];
            float x1 = d0 * T[1] + d1 * T[5] + d2 * T[9];
            float x2 = d0 * T[2] + d1 * T[6] + d2 * T[10];
            live = !ia_all_corners_oob(p, p.scl[0] * (x0 + p.off[0]), p.scl[1] * (x1 + p.off[1]), p.scl[2] * (x2 + p.off[2]));
            if (!live) c_skip++;
        }
        const unsigned b = __ballot_sync(0xffffffffu, live);
        if (lane == 0) ballots[it] = b;
        n_live += __popc(b);
    }
    int base = 0;
    if (lane == 0 && n_live) base = atomicAdd(&S.n_btask, n_live);
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncwarp();
    for (int k0 = k_begin, it = 0; k0 < k_end; k0 += 32, it++) {
        const unsigned b = ballots[it];
        if ((b >> lane) & 1u) {
            const int k = k0 + lane;
            const int c = k / n_q;
            const int q = S.qlist[k - c * n_q];
            S.btask[base + __popc(b & ((1u << lane) - 1u))] = (unsigned short)(q * 16 + c);
        }
        base += __popc(b);
    }
    wf_count(S, WF_C_SKIP, c_skip);
}

__device__ __forceinline__ void wf_broyden_phase(const IaFrame& p, WfShared& S) {
    // chain state in ONE register (the loop sits at the 128-register cap): task id (q * 16 + c) in bits 0..15, rank-1 update
    // count in bits 16..19, bit 20 = the next trip is the chain's first; -1 = no task left
    const int FRESH = 1 << 20;
    int st;
    {
        const int t = wf_grab_chain(S);
        st = t >= 0 ? ((int)S.btask[t] | FRESH) : -1;
    }
    float x0 = 0, x1 = 0, x2 = 0, xd0 = 0, xd1 = 0, xd2 = 0, g0 = 0, g1 = 0, g2 = 0;
    float Ji[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    while (st >= 0) {
        float u0 = 0, u1 = 0, u2 = 0;
        const bool fresh = st & FRESH;
        if (fresh) {
            const int q = (st & 0xffff) >> 4, c = st & 15;
            xd0 = S.qx[0][q]; xd1 = S.qx[1][q]; xd2 = S.qx[2][q];
            const float* T = S.tfs13 + c * 12;
            float d0 = xd0 - T[3], d1 = xd1 - T[7], d2 = xd2 - T[11];
            x0 = d0 * T[0] + d1 * T[4] + d2 * T[8];
            x1 = d0 * T[1] + d1 * T[5] + d2 * T[9];
            x2 = d0 * T[2] + d1 * T[6] + d2 * T[10];
        } else {
            u0 = -Ji[0] * g0 + -Ji[1] * g1 + -Ji[2] * g2;
            u1 = -Ji[3] * g0 + -Ji[4] * g1 + -Ji[5] * g2;
            u2 = -Ji[6] * g0 + -Ji[7] * g1 + -Ji[8] * g2;
            x0 += u0; x1 += u1; x2 += u2;
        }
        const float ix = p.scl[0] * (x0 + p.off[0]);
        const float iy = p.scl[1] * (x1 + p.off[1]);
        const float iz = p.scl[2] * (x2 + p.off[2]);
        float J[12];
        ia_fetch_J(p, ix, iy, iz, J);
        const float n0 = J[0] * x0 + J[1] * x1 + J[2] * x2 + J[3] - xd0;
        const float n1 = J[4] * x0 + J[5] * x1 + J[6] * x2 + J[7] - xd1;
        const float n2 = J[8] * x0 + J[9] * x1 + J[10] * x2 + J[11] - xd2;
        if (fresh) {
            Ji[0] = J[0]; Ji[3] = J[1]; Ji[6] = J[2];
            Ji[1] = J[4]; Ji[4] = J[5]; Ji[7] = J[6];
            Ji[2] = J[8]; Ji[5] = J[9]; Ji[8] = J[10];
            g0 = n0; g1 = n1; g2 = n2;
            st &= 0xffff;   // update count 0, not fresh
        } else {
            const float nrm = n0 * n0 + n1 * n1 + n2 * n2;
            bool fin = false, ok = false;
            if (nrm < 1e-5f * 1e-5f) {
                ok = ix >= -1 && ix <= 1 && iy >= -1 && iy <= 1 && iz >= -1 && iz <= 1;
                fin = true;
            } else if (nrm > 1e-1f * 1e-1f) {
                fin = true;
            } else {
                // rank-1 update of the inverse Jacobian (fuse_J_inv_update, :22-55)
                float dg0 = n0 - g0, dg1 = n1 - g1, dg2 = n2 - g2;
                float c0 = Ji[0] * u0 + Ji[3] * u1 + Ji[6] * u2;
                float c1 = Ji[1] * u0 + Ji[4] * u1 + Ji[7] * u2;
                float c2 = Ji[2] * u0 + Ji[5] * u1 + Ji[8] * u2;
                float s = c0 * dg0 + c1 * dg1 + c2 * dg2;
                float r0 = -Ji[0] * dg0 - Ji[1] * dg1 - Ji[2] * dg2;
                float r1 = -Ji[3] * dg0 - Ji[4] * dg1 - Ji[5] * dg2;
                float r2 = -Ji[6] * dg0 - Ji[7] * dg1 - Ji[8] * dg2;
                Ji[0] += c0 * (r0 + u0) / s; Ji[1] += c1 * (r0 + u0) / s; Ji[2] += c2 * (r0 + u0) / s;
                Ji[3] += c0 * (r1 + u1) / s; Ji[4] += c1 * (r1 + u1) / s; Ji[5] += c2 * (r1 + u1) / s;
                Ji[6] += c0 * (r2 + u2) / s; Ji[7] += c1 * (r2 + u2) / s; Ji[8] += c2 * (r2 + u2) / s;
                g0 = n0; g1 = n1; g2 = n2;
                st += 1 << 16;
                if ((st >> 16) >= 10) fin = true;
            }
            if (fin) {
                const int q = (st & 0xffff) >> 4, c = st & 15, it = st >> 16;
                if (ok) {
                    float* cd = S.cand + (q * IA_N_INIT + c) * 3;
                    cd[0] = x0; cd[1] = x1; cd[2] = x2;
                    atomicOr(&S.qmask[q], 1u << c);
                }
                // voxel fetches of this chain: the initial one + one per later trip (`it` counts the rank-1 updates).
                // Counted here, at the chain's end, so that no counter lives in a register across the loop.
                atomicAdd(&S.cnt[WF_C_FETCH], 1u + (it >= 10 ? 10u : (unsigned)it + 1u));
                const int t = wf_grab_chain(S);
                st = t >= 0 ? ((int)S.btask[t] | FRESH) : -1;
            }
        }
    }
}

// filter.cu:10-54 per pending query, then the list of geometry tasks
#ifndef WF_FILTER_PRELOAD
#define WF_FILTER_PRELOAD 1   // 1: a query with two or more roots loads all of them at once (independent loads from the CTA's
#endif                        // global scratch, one L2 round trip) and compares in registers; 0: the reference's nested loops,
                              // one dependent round trip per compared pair while the rest of the warp waits
__device__ __forceinline__ void wf_filter_phase(WfShared& S, int n_q) {
    for (int k0 = 0; k0 < n_q; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        const bool act = k < n_q;
        const int q = act ? S.qlist[k] : 0;
        const unsigned mask = act ? S.qmask[q] : 0u;
        unsigned keep = mask;
        const float* cd = S.cand + q * IA_N_INIT * 3;
#if WF_FILTER_PRELOAD
        if (mask & (mask - 1u)) {
            float c0[IA_N_INIT], c1[IA_N_INIT], c2[IA_N_INIT];
#pragma unroll
            for (int i = 0; i < IA_N_INIT; i++) {
                const bool v = (mask >> i) & 1u;
                c0[i] = v ? cd[i * 3] : 0.f; c1[i] = v ? cd[i * 3 + 1] : 0.f; c2[i] = v ? cd[i * 3 + 2] : 0.f;
            }
            // root i goes when a LATER valid root lies within 1e-4 of it
#pragma unroll
            for (int i = 0; i < IA_N_INIT - 1; i++) {
                bool dup = false;
#pragma unroll
                for (int j = i + 1; j < IA_N_INIT; j++) {
                    const float e0 = c0[i] - c0[j], e1 = c1[i] - c1[j], e2 = c2[i] - c2[j];
                    dup = dup || (((mask >> j) & 1u) && e0 * e0 + e1 * e1 + e2 * e2 < 0.0001f * 0.0001f);
                }
                if (dup) keep &= ~(1u << i);
            }
            keep &= mask;
        }
#else
        unsigned mi = mask;
        while (mi) {
            int i = __ffs(mi) - 1;
            mi &= mi - 1;
            unsigned mj = mi;  // later valid candidates
            while (mj) {
                int j = __ffs(mj) - 1;
                mj &= mj - 1;
                float e0 = cd[i * 3] - cd[j * 3], e1 = cd[i * 3 + 1] - cd[j * 3 + 1], e2 = cd[i * 3 + 2] - cd[j * 3 + 2];
                if (e0 * e0 + e1 * e1 + e2 * e2 < 0.0001f * 0.0001f) { keep &= ~(1u << i); break; }
            }
        }
#endif
        if (act) S.qmask[q] = keep;
        // the warp's kept roots go to ONE stretch of the task list, in query (Morton) order: one atomic per warp
        const int lane = threadIdx.x & 31;
        const int n = act ? __popc(keep) : 0;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 0 && total) base = atomicAdd(&S.n_gtask, total);
        base = __shfl_sync(0xffffffffu, base, 0) + incl - n;
        unsigned m = keep;
        while (m) {
            int c = __ffs(m) - 1;
            m &= m - 1;
            S.gtask[base++] = (unsigned short)(q * 16 + c);
        }
    }
}

// SDF of up to 16 canonical points by one warp on the tensor cores (ia_mma.cuh).  Lane j < n_pts holds point j in
// (px, py, pz); returns the sdf of point j in lane j < 16.  Tile columns: 0..31 hash features (level l: 2l, 2l + 1),
// 32..34 the scaled position 2 xn - 1, 35..39 zero (set once at kernel start, never written).
#ifndef WF_MMA_ROWS
#define WF_MMA_ROWS 16   // points per tensor-core batch of a warp: 16 (the full m16 tile) or 8 (half the shared memory for the input
#endif                   // tiles -- L1 capacity for the gathers -- at twice the weight-fragment reads per point)
__device__ __forceinline__ float ia_geo_w1(const float* __restrict__ mlp, int k, int n) {   // layer-1 weight of tile column k
    const int in = k < 32 ? 3 + k : (k < 35 ? k - 32 : -1);
    return in < 0 ? 0.f : mlp[IA_GEO_W1T + in * 64 + n];
}
template <int ROWS>
__device__ __forceinline__ float ia_warp_geometry(const IaFrame& p, const IaLevel* __restrict__ lvl, const float* __restrict__ w,
                                                    const float4* __restrict__ w1f, float* __restrict__ xs, float px, float py,
                                                    float pz, int n_pts) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, l = lane & 15, half = lane >> 4;
    // ---- A: hash-grid features, two points per trip (lane = level), into the warp's input tile
    for (int i = 0; i < n_pts; i += 2) {
        const int r = i + half;
        const float x0 = __shfl_sync(FULL, px, r & 15), x1 = __shfl_sync(FULL, py, r & 15), x2 = __shfl_sync(FULL, pz, r & 15);
        if (r < n_pts) {
            float xn[3] = {(x0 - p.center[0]) / p.scale[0] + 0.5f, (x1 - p.center[1]) / p.scale[1] + 0.5f,
                           (x2 - p.center[2]) / p.scale[2] + 0.5f};
            float f0, f1;
            ia_hash_level<false>(p.geo_hash, lvl[l], xn, f0, f1, nullptr);
            *reinterpret_cast<float2*>(xs + r * IA_GEO_LD + 2 * l) = make_float2(f0, f1);
            if (l < 3) xs[r * IA_GEO_LD + 32 + l] = xn[l] * 2.0f - 1.0f;
        }
    }
    __syncwarp();
    // ---- B + C: 35 -> 64 layer (16 x 64 pre-activations in C layout, two halves of 32 hidden units to bound the
    //      accumulator registers), softplus, row 0 of the 64 -> 13 layer (the sdf)
    const int t = lane & 3;
    float lo = 0.f, hi = 0.f;
#pragma unroll
    for (int nh = 0; nh < 2; nh++) {
        float c[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const float2 b = *reinterpret_cast<const float2*>(w + IA_GEOC_B1 + 8 * (4 * nh + nt) + 2 * t);
            c[nt][0] = b.x; c[nt][1] = b.y; c[nt][2] = b.x; c[nt][3] = b.y;
        }
        ia_mma_layer_smem<IA_GEO_KSTEPS, 4, 8, ROWS>(xs, IA_GEO_LD, w1f + 4 * nh * 32, c);
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const float2 w2 = *reinterpret_cast<const float2*>(w + IA_GEOC_W2 + 8 * (4 * nh + nt) + 2 * t);
            lo = fmaf(w2.x, ia_softplus100(c[nt][0]), lo); lo = fmaf(w2.y, ia_softplus100(c[nt][1]), lo);
            if (ROWS == 16) { hi = fmaf(w2.x, ia_softplus100(c[nt][2]), hi); hi = fmaf(w2.y, ia_softplus100(c[nt][3]), hi); }
        }
    }
    __syncwarp();   // the tile may be overwritten by the next batch
    // reduction over the four lanes of a row
    lo += __shfl_xor_sync(FULL, lo, 1); hi += __shfl_xor_sync(FULL, hi, 1);
    lo += __shfl_xor_sync(FULL, lo, 2); hi += __shfl_xor_sync(FULL, hi, 2);
    // row j < 8 sits in lanes 4j .. 4j + 3 (lo), row j >= 8 in lanes 4 (j - 8) .. (hi)
    const float slo = __shfl_sync(FULL, lo, 4 * (lane & 7)), shi = __shfl_sync(FULL, hi, 4 * (lane & 7));
    return ((lane & 8) ? shi : slo) + w[IA_GEOC_B2];
}

__device__ __forceinline__ void wf_geometry_phase(const IaFrame& p, WfShared& S) {
    unsigned c_geo = 0;
    const int n = S.n_gtask;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    // an even split of the task list over the warps (the gathers of step A are the cost, and they are per point).  Batches dealt
    // round-robin instead (the sixteen warps walking the Morton-ordered list side by side) measured the same: 756.6 vs 754.9 ms.
    const int per = (n + n_warps - 1) / n_warps;
    const int end = min(n, (warp + 1) * per);
    const int b_step = WF_MMA_ROWS;
    int b0 = warp * per;
    float* xs = wf_xs(S) + warp * WF_MMA_ROWS * IA_GEO_LD;
    // software pipeline: the task ids and roots of the NEXT batch (two dependent L2 round trips: the lists were just
    // written by other warps) are fetched while the current one is evaluated
    int ci = 0;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (b0 + lane < end && lane < WF_MMA_ROWS) {
        const int tk = S.gtask[b0 + lane];
        ci = (tk >> 4) * IA_N_INIT + (tk & 15);
        const float* cd = S.cand + ci * 3;
        x0 = cd[0]; x1 = cd[1]; x2 = cd[2];
    }
    while (b0 < end) {
        const int nb = min(WF_MMA_ROWS, end - b0);
        const int bn = b0 + b_step;
        int ci_n = 0;
        float y0 = 0.f, y1 = 0.f, y2 = 0.f;
        if (bn + lane < end && lane < WF_MMA_ROWS) {
            const int tk = S.gtask[bn + lane];
            ci_n = (tk >> 4) * IA_N_INIT + (tk & 15);
            const float* cd = S.cand + ci_n * 3;
            y0 = cd[0]; y1 = cd[1]; y2 = cd[2];
        }
        const float s = ia_warp_geometry<WF_MMA_ROWS>(p, S.lvl, wf_w(S), wf_w1f(p, S), xs, x0, x1, x2, nb);
        if (lane < nb) { S.csdf[ci] = s; c_geo++; }
        b0 = bn; ci = ci_n; x0 = y0; x1 = y1; x2 = y2;
    }
    wf_count(S, WF_C_GEO, c_geo);
}

// ------------------------------------------------------------------------------------------------
// Tensor-core shading of up to 16 canonical points by one warp: geometry feature + SDF gradient, blended forward rotation,
// radiance network (rgb_alpha_fn, models/intrinsic_avatar.py:430-456; models/rf/geometry.py:147-172 forward + autograd
// normal; models/rf/radiance.py:111-135).  Lane j < 16 passes point j (canonical root) and its ray direction (SMPL space);
// lanes j and j + 16 return rgb of point j.  `xs` is the warp's shading tile [16][IA_SHADE_LD] (layout: ia_mma.cuh).
//   A1  hash-grid features of the geometry network, two points per trip (lane = level)             -> tile cols 0..34
//   B   35 -> 64 layer as 3xTF32 mma, softplus; the 64 -> 13 layer (feature) and the BACKWARD of the sdf row through the
//       35 -> 64 layer (d sdf / d tile column), both fed from the accumulator registers            -> cols 35..47, 0..31, 64..66
//   A2  per point (16-lane team): hash-grid gradient (corners re-gathered, L1-hot) x d sdf / d feature -> canonical normal,
//       blended rotation -> world normal, reflected direction -> SH, radiance hash features        -> cols 0..31, 48..66
//   C   radiance network 67 -> 64 -> 64 -> 3 as 3xTF32 mma, sigmoid
// The weights of B (second half) and C are read as B fragments straight from global memory (IA_FRAG_*, 102 KB pre-split, L1-resident
// while this phase runs and evicted for the voxel gathers otherwise) instead of living in shared memory.
// PRIMARY (the shading kernel of the primary stage, k_prim_shade_wf): lane j also passes the index of its sample record;
// A2 stores the posed-space unit normal and the world normal there (rgb_normal_mats_alpha_fn, models/intrinsic_avatar.py:
// 1066-1156), and the material network (48 -> 64 -> 64 -> 5 on the first 48 tile columns) runs after the radiance network.
template <bool PRIMARY = false>
__device__ __forceinline__ void ia_warp_radiance16(const IaFrame& p, const IaLevel* __restrict__ lvl, const float* __restrict__ w,
                                                   const float4* __restrict__ w1f, float* __restrict__ xs, float px, float py,
                                                   float pz, float dx, float dy, float dz, int n_pts, float rgb[3],
                                                   float mat[5] = nullptr, long long rec = 0, IaSample* __restrict__ samples = nullptr,
                                                   IaSampleAux* __restrict__ aux = nullptr) {
    const unsigned FULL = 0xffffffffu;
    constexpr int LD = IA_SHADE_LD;
    const int lane = threadIdx.x & 31, l = lane & 15, half = lane >> 4, g = lane >> 2, t = lane & 3;
    const float4* frags = reinterpret_cast<const float4*>(p.mlp + IA_MLP_END);
    // ---- A1
#pragma unroll 1
    for (int i = 0; i < n_pts; i += 2) {
        const int r = i + half;
        const float x0 = __shfl_sync(FULL, px, r & 15), x1 = __shfl_sync(FULL, py, r & 15), x2 = __shfl_sync(FULL, pz, r & 15);
        if (r < n_pts) {
            float xn[3] = {(x0 - p.center[0]) / p.scale[0] + 0.5f, (x1 - p.center[1]) / p.scale[1] + 0.5f,
                           (x2 - p.center[2]) / p.scale[2] + 0.5f};
            float f0, f1;
            ia_hash_level<false>(p.geo_hash, lvl[l], xn, f0, f1, nullptr);
            *reinterpret_cast<float2*>(xs + r * LD + 2 * l) = make_float2(f0, f1);
            if (l < 3) xs[r * LD + 32 + l] = xn[l] * 2.0f - 1.0f;
        }
    }
    __syncwarp();
    // ---- B
    {
        float c[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const float2 b = *reinterpret_cast<const float2*>(w + IA_GEOC_B1 + 8 * nt + 2 * t);
            c[nt][0] = b.x; c[nt][1] = b.y; c[nt][2] = b.x; c[nt][3] = b.y;
        }
        ia_mma_layer_smem<IA_GEO_KSTEPS, 8, 8, 16>(xs, LD, w1f, c);
        __syncwarp();   // columns 0..39 of the tile are free
        // h = softplus_100(pre) in place; delta = W2[0][n] softplus'(pre) = W2[0][n] sigmoid(100 pre)
        float dl[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const float2 w2 = *reinterpret_cast<const float2*>(w + IA_GEOC_W2 + 8 * nt + 2 * t);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float x = c[nt][k];
                dl[nt][k] = ((k & 1) ? w2.y : w2.x) * ia_sigmoid(100.f * x);
                c[nt][k] = ia_softplus100(x);
            }
        }
        float f[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            const float2 b = *reinterpret_cast<const float2*>(w + IA_GEOC_B2 + 8 * nt + 2 * t);   // [16], 13 used, rest 0
            f[nt][0] = b.x; f[nt][1] = b.y; f[nt][2] = b.x; f[nt][3] = b.y;
        }
        ia_mma_layer_regs<2>(c, frags + IA_FRAG_FEAT * 32, f);
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int col = 8 * nt + 2 * t + k;
                if (col < 13) { xs[g * LD + 35 + col] = f[nt][k]; xs[(g + 8) * LD + 35 + col] = f[nt][2 + k]; }
            }
        float gi[5][4];
#pragma unroll
        for (int nt = 0; nt < 5; nt++) gi[nt][0] = gi[nt][1] = gi[nt][2] = gi[nt][3] = 0.f;
        ia_mma_layer_regs<5>(dl, frags + IA_FRAG_BWD * 32, gi);
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            *reinterpret_cast<float2*>(xs + g * LD + 8 * nt + 2 * t) = make_float2(gi[nt][0], gi[nt][1]);
            *reinterpret_cast<float2*>(xs + (g + 8) * LD + 8 * nt + 2 * t) = make_float2(gi[nt][2], gi[nt][3]);
        }
        // d sdf / d (2 xn - 1): tile columns 32..34 keep the position for the radiance network, the gradient goes to 64..66
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int col = 2 * t + k;
            if (col < 3) { xs[g * LD + 64 + col] = gi[4][k]; xs[(g + 8) * LD + 64 + col] = gi[4][2 + k]; }
        }
    }
    __syncwarp();
    // ---- A2
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());   // = half-warp: lane l, point i + half
#pragma unroll 1
    for (int i = 0; i < n_pts; i += 2) {
        const int r = i + half;
        const float x0 = __shfl_sync(FULL, px, r & 15), x1 = __shfl_sync(FULL, py, r & 15), x2 = __shfl_sync(FULL, pz, r & 15);
        const float d0 = __shfl_sync(FULL, dx, r & 15), d1 = __shfl_sync(FULL, dy, r & 15), d2 = __shfl_sync(FULL, dz, r & 15);
        const long long rec_r = PRIMARY ? __shfl_sync(FULL, rec, r & 15) : 0;
        if (r < n_pts) {
            float* row = xs + r * LD;
            const float xc[3] = {x0, x1, x2};
            float xn[3] = {(x0 - p.center[0]) / p.scale[0] + 0.5f, (x1 - p.center[1]) / p.scale[1] + 0.5f,
                           (x2 - p.center[2]) / p.scale[2] + 0.5f};
            float f0, f1, dfdx[6];
            ia_hash_level<true>(p.geo_hash, lvl[l], xn, f0, f1, dfdx);
            const float2 gf = *reinterpret_cast<const float2*>(row + 2 * l);
            float gc[3];
#pragma unroll
            for (int d = 0; d < 3; d++)
                gc[d] = (row[64 + d] * 2.0f + ia_team_sum(team, gf.x * dfdx[d * 2 + 0] + gf.y * dfdx[d * 2 + 1])) / p.scale[d];
            float R[9];
            ia_team_fwd_rotation(team, p, xc, R);
            const float gs[3] = {R[0] * gc[0] + R[1] * gc[1] + R[2] * gc[2], R[3] * gc[0] + R[4] * gc[1] + R[5] * gc[2],
                                 R[6] * gc[0] + R[7] * gc[1] + R[8] * gc[2]};
            const float dir[3] = {d0, d1, d2};
            float nw[3], view_w[3];
            ia_dir_s2w(p, gs, nw);
            ia_dir_s2w(p, dir, view_w);
            if (PRIMARY && l == 0) {
                float nsm[3];
                ia_normalize(gs, nsm, 1e-6f);
                samples[rec_r].n[0] = nsm[0]; samples[rec_r].n[1] = nsm[1]; samples[rec_r].n[2] = nsm[2];
                aux[rec_r].nw[0] = nw[0]; aux[rec_r].nw[1] = nw[1]; aux[rec_r].nw[2] = nw[2];
            }
            ia_hash_level<false>(p.rad_hash, lvl[l], xn, f0, f1, nullptr);
            // reflect(-view, n) (models/utils.py:115), then the (d + 1) / 2 -> 2 x - 1 round trip of the encoding
            const float v[3] = {-view_w[0], -view_w[1], -view_w[2]};
            const float dn = v[0] * nw[0] + v[1] * nw[1] + v[2] * nw[2];
            float rr[3], sh[16];
#pragma unroll
            for (int d = 0; d < 3; d++) rr[d] = ((2.f * dn * nw[d] - v[d] + 1.f) / 2.f) * 2.f - 1.f;
            ia_sh4(rr[0], rr[1], rr[2], sh);
            float my_sh = 0.f, my_n = 0.f;
#pragma unroll
            for (int o = 0; o < 16; o++) my_sh = l == o ? sh[o] : my_sh;
#pragma unroll
            for (int o = 0; o < 3; o++) my_n = l == o ? nw[o] : my_n;
            team.sync();   // every lane of the team has read the row's gradient columns
            *reinterpret_cast<float2*>(row + 2 * l) = make_float2(f0, f1);
            row[48 + l] = my_sh;
            if (l < 3) row[64 + l] = my_n;
        }
    }
    __syncwarp();
    // ---- C
    float o[3];
    ia_warp_mlp3<9, 3>(xs, LD, frags + IA_FRAG_RAD1 * 32, p.mlp + IA_RAD_B1, p.mlp + IA_RAD_B2, p.mlp + IA_RAD_B3, o);
#pragma unroll
    for (int k = 0; k < 3; k++) rgb[k] = ia_sigmoid(o[k]);
    if (PRIMARY) {
        float m[5];
        ia_warp_mlp3<6, 5>(xs, LD, frags + IA_FRAG_MAT1 * 32, p.mlp + IA_MAT_B1, p.mlp + IA_MAT_B2, p.mlp + IA_MAT_B3, m);
#pragma unroll
        for (int k = 0; k < 5; k++) mat[k] = ia_sigmoid(m[k]) * p.mat_scale[k] + p.mat_bias[k];
#pragma unroll
        for (int k = 0; k < 3; k++) mat[k] *= p.albedo_ratio[k];
    }
    __syncwarp();   // the tile may be overwritten by the next batch
}

// The geometry phase reads columns 35..39 of the rows of ITS tiles ([warps][16][IA_GEO_LD] over the same memory as the shading
// tiles) against zero weights; they must stay finite: a warp that used its shading tile zeroes the ones inside it.
__device__ __forceinline__ void wf_restore_geo_pads(WfShared& S, int warp) {
    const int lane = threadIdx.x & 31;
    float* all = wf_xs(S);
    const int f0 = warp * IA_SHADE_TILE, f1 = f0 + IA_SHADE_TILE;
    for (int r = f0 / IA_GEO_LD + (lane / 5); r * IA_GEO_LD + 35 < f1 && r < (WF_THREADS / 32) * 16; r += 6) {
        const int o = r * IA_GEO_LD + 35 + lane % 5;
        if (lane < 30 && o >= f0 && o < f1) all[o] = 0.f;
    }
}

// GI: radiance at the arg-min root of every fine sample consumed this round.
__device__ __forceinline__ void wf_gi_phase(const IaFrame& p, WfShared& S) {
    const int n = S.n_gitask;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int n_warps = WF_GI_WARPS;     // the warps that own a shading tile; the others go straight to the Broyden phase
    if (warp >= n_warps) return;
    // whole 16-row batches: a round has 100-200 tasks, i.e. 6-12 rows per warp if split evenly, and the tensor-core stages
    // cost the same for 6 rows as for 16.  The warps left without a batch go straight to the Broyden phase's task loop.
    const int per = (((n + n_warps - 1) / n_warps) + 15) & ~15;
    const int end = min(n, (warp + 1) * per);
    float* xs = wf_xs(S) + warp * IA_SHADE_TILE;
    for (int b0 = warp * per; b0 < end; b0 += 16) {
        const int nb = min(16, end - b0);
        int t = 0;
        float wgt = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 1.f;
        if ((lane & 15) < nb) {
            const uint2 tk = S.gitask[b0 + (lane & 15)];
            t = tk.x & 0xffffu;
            wgt = __uint_as_float(tk.y);
            const float* cd = S.gixc + (b0 + (lane & 15)) * 3;
            x0 = cd[0]; x1 = cd[1]; x2 = cd[2];
            d0 = S.st[WS_D][t]; d1 = S.st[WS_D + 1][t]; d2 = S.st[WS_D + 2][t];
        }
        float rgb[3];
        ia_warp_radiance16(p, S.lvl, wf_w(S), wf_w1f(p, S), xs, x0, x1, x2, d0, d1, d2, nb, rgb);
        if (lane < nb) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) S.st[WS_IND + ch][t] += wgt * rgb[ch];
        }
    }
    if (end > warp * per) wf_restore_geo_pads(S, warp);
    wf_count(S, WF_C_QG, (lane == 0 && end > warp * per) ? (unsigned)(end - warp * per) : 0u);   // = geometry evaluations with gradient = radiance evaluations = skinning fetches of this phase
}

// ------------------------------------------------------------------------------------------------
// One round of the ray state machines.  Policy P supplies rays and consumes their transmittance:
//   bool P::init(const uint2 entry, float o[3], float d[3])   ray of a ring entry
//   void P::finish(const uint2 entry, float T, const float ind[3], const float d[3])   shade / store (ind = indirect
//                                                                     radiance, GI; d = the ray direction)
//   int P::tile_items()   shading samples examined per feed step (each may push up to WF_QCAP / 2 / tile_items() rays)
// The advance phase runs in two passes so that the expensive, rare events do not serialise the warps (round 1 ran
// everything in one pass: a slot whose ray ended -- ~1 of 12 per round -- dragged its 31 neighbours through BRDF
// evaluation, ray set-up and the march to the first occupied cell at 2-3 lanes of 32):
//   consume : one thread per slot: SDF of the previous query -> state machine -> next query point.  A slot whose ray has
//             ended only records its transmittance and joins the refill list, as does an idle slot.
//   refill  : one thread per LIST ENTRY (dense): estimator of the finished ray (P::finish), next ray from the ring
//             (P::init, marcher set-up, march to its first sample), its first query.
// Measured at 512^2 x 1024 spp, GI on, frames 2-4: 1159 -> 1102 ms (and 1088 -> 1035 ms with WF_QSORT).
__device__ __forceinline__ void wf_push_gi(WfShared& S, int t, int best, float w) {
    const int gi = atomicAdd(&S.n_gitask, 1);
    S.gitask[gi] = make_uint2((unsigned)t | ((unsigned)best << 16), __float_as_uint(w));
    const float* cd = S.cand + (t * IA_N_INIT + best) * 3;
    S.gixc[gi * 3] = cd[0]; S.gixc[gi * 3 + 1] = cd[1]; S.gixc[gi * 3 + 2] = cd[2];
}
template <class P>
__device__ __forceinline__ void wf_finish(P& pol, WfShared& S, const uint2 entry, float T, const float ind[3], const float d[3]) {
#if WF_DEFER_FINISH
    const int k = atomicAdd(&S.n_flist, 1);
    if (k < WF_FCAP) {
        float* f = S.flist + k * 9;
        f[0] = __uint_as_float(entry.x); f[1] = __uint_as_float(entry.y); f[2] = T;
        f[3] = ind[0]; f[4] = ind[1]; f[5] = ind[2]; f[6] = d[0]; f[7] = d[1]; f[8] = d[2];
        return;
    }
#endif
    pol.finish(entry, T, ind, d);
}
template <class P>
__device__ __forceinline__ void wf_finish_phase(P& pol, WfShared& S) {
#if WF_DEFER_FINISH
    const int n = min(S.n_flist, WF_FCAP);
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const float* f = S.flist + k * 9;
        const float ind[3] = {f[3], f[4], f[5]}, d[3] = {f[6], f[7], f[8]};
        pol.finish(make_uint2(__float_as_uint(f[0]), __float_as_uint(f[1])), f[2], ind, d);
    }
#endif
}
template <bool GI, class P>
__device__ __forceinline__ void wf_advance_consume(const IaFrame& p, P& pol, WfShared& S, const int t, unsigned& c_q) {
    unsigned pack = __float_as_uint(S.st[WS_PACK][t]);
    int stage = pack & 7u, j = (pack >> 3) & 7u, i = (pack >> 6) & 7u;
    const unsigned ey = pack >> 16;
    if (stage == WF_IDLE) {
        S.rlist[atomicAdd(&S.n_rlist, 1)] = (unsigned short)t;
        return;
    }
    float o[3], d[3];
    IaMarcher m;
    float sdf_prev = 0, cs = 0, ce = 0, ts = 0, te = 0, trans = 0, cdf_prev = 0, cdf_next = 0, cdf_u = 0;
    int action = WF_ACT_NEXT;
    bool pending = false;
    float sdf_cur = 0, Tfin = 1.0f;
    const float cdf_step_size = (1.0f - 1.0 / 5) / 4;
    // ---- min SDF over the kept roots of the previous query (snarf_deformer.py:242-259)
    unsigned keep = S.qmask[t];
    float sdf = 1e5f;
    int best = -1;
    while (keep) {
        int c = __ffs(keep) - 1;
        keep &= keep - 1;
        float s = S.csdf[t * IA_N_INIT + c];
        if (s < sdf) { sdf = s; best = c; }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { o[k] = S.st[WS_O + k][t]; d[k] = S.st[WS_D + k][t]; }
    wf_load_marcher(S, t, m, p.sec_step);
    ts = S.st[WS_TS][t]; te = S.st[WS_TE][t];
    if (p.sec_mode == IA_SEC_PLAIN) {
        // secondary_importance_sample = false (models/intrinsic_avatar.py:482-520 skipped): the coarse samples themselves are
        // rendered -- query at the MIDPOINT of every sample, w = T alpha, T *= 1 - alpha, radiance at every sample (GI).
        // Stage SEARCH = accumulating; WS_TRANS = T, WS_CDFPREV = sum of weights.
        float Tacc = 1.0f, acc = 0.0f;
        if (stage == WF_GIWAIT) {
            Tfin = S.st[WS_TRANS][t];
            stage = WF_FINE;                       // (marks "indirect radiance present" for the finish)
            action = WF_ACT_FINISH;
        } else {
            if (stage == WF_SEARCH) { Tacc = S.st[WS_TRANS][t]; acc = S.st[WS_CDFPREV][t]; }
            else if (GI) { S.st[WS_IND][t] = 0.f; S.st[WS_IND + 1][t] = 0.f; S.st[WS_IND + 2][t] = 0.f; }   // first sample
            const float al = ia_alpha(sdf, te - ts, p.beta);
            const float w = Tacc * al;
            Tacc *= (1.0f - al);
            acc += w;
            bool gi_pushed = false;
            if (GI && best >= 0) {
                wf_push_gi(S, t, best, w);
                gi_pushed = true;
            }
            bool cont;
            if (m.next(p.occ_bits, p.occ_res, ts, te, cont)) {
#pragma unroll
                for (int k = 0; k < 3; k++) S.qx[k][t] = o[k] + d[k] * ((ts + te) / 2.0f);
                wf_store_marcher(S, t, m);
                S.st[WS_TS][t] = ts; S.st[WS_TE][t] = te;
                S.st[WS_TRANS][t] = Tacc; S.st[WS_CDFPREV][t] = acc;
                int idx = atomicAdd(&S.n_q, 1);
                S.qlist[idx] = (unsigned short)t;
                S.qmask[t] = 0;
                c_q++;
                S.st[WS_PACK][t] = __uint_as_float((unsigned)WF_SEARCH | (ey << 16));
                return;
            }
            Tfin = 1.0f - acc;
            if (GI && gi_pushed) {
                S.st[WS_TRANS][t] = Tfin;
                S.st[WS_PACK][t] = __uint_as_float((unsigned)WF_GIWAIT | (ey << 16));
                return;
            }
            stage = GI ? WF_FINE : WF_SEARCH;
            action = WF_ACT_FINISH;
        }
    } else if (stage == WF_FIRST && p.sec_mode == IA_SEC_IMPORTANCE) {
        // zero_crossing_search = false: ray_resampling_fine (cdf.cu:403-478) over the weights of ALL coarse samples, i.e.
        // the CDF walk starts at the first sample (weights_sum = max(sum w, 1) = 1: sum w = 1 - T <= 1)
        cs = ts; ce = te;
        j = 0;
        float a = ia_alpha(sdf, ce - cs, p.beta);
        trans = 1.0f - a;
        cdf_prev = 0.0f; cdf_next = a;
        cdf_u = 1.0 / (2 * 5);
        action = WF_ACT_CDF;
    } else if (stage == WF_FIRST) {
        sdf_prev = sdf; cs = ts; ce = te;
        stage = WF_SEARCH;
        action = WF_ACT_NEXT;
    } else if (stage == WF_SEARCH) {
        sdf_prev = S.st[WS_SDFPREV][t];
        if (sdf_prev >= 0 && sdf < 0) {
            cs = S.st[WS_CS][t]; ce = S.st[WS_CE][t];
            sdf_cur = sdf;
            pending = true;  // (ts, te, sdf_cur) is the already-queried interval after the crossing one
            j = 0;
            float a = ia_alpha(sdf_prev, ce - cs, p.beta);
            float weight = a;
            trans = 1.0f;
            trans *= (1.0f - a);
            cdf_prev = 0.0f; cdf_next = weight;
            cdf_u = 1.0 / (2 * 5);
            action = WF_ACT_CDF;
        } else {
            sdf_prev = sdf; cs = ts; ce = te;
            action = WF_ACT_NEXT;
        }
    } else if (stage == WF_CDF) {
        trans = S.st[WS_TRANS][t]; cdf_prev = S.st[WS_CDFPREV][t]; cdf_next = S.st[WS_CDFNEXT][t];
        cdf_u = S.st[WS_CDFU][t];
        cs = ts; ce = te;
        float a = ia_alpha(sdf, ce - cs, p.beta);
        float weight = trans * a;
        trans *= (1.0f - a);
        cdf_prev = cdf_next;
        cdf_next += weight;
        action = WF_ACT_CDF;
    } else if (GI && stage == WF_GIWAIT) {
        Tfin = S.st[WS_TRANS][t];
        stage = WF_FINE;
        action = WF_ACT_FINISH;
    } else {  // WF_FINE
        float Tacc = S.st[WS_TRANS][t], acc = S.st[WS_CDFPREV][t];
        float s0 = S.st[WS_TPL + i][t], e0 = S.st[WS_TPL + i + 1][t];
        float al = ia_alpha(sdf, e0 - s0, p.beta);
        float w = Tacc * al;
        Tacc *= (1.0f - al);
        acc += w;
        bool gi_pushed = false;
        if (GI && best >= 0) {
            // radiance at the arg-min root of this fine sample is added by the GI phase of this round: ind += w * rgb
            wf_push_gi(S, t, best, w);
            gi_pushed = true;
        }
        i++;
        if (i + 1 < j) {
            float s1 = S.st[WS_TPL + i][t], e1 = S.st[WS_TPL + i + 1][t];
            float mid = (s1 + e1) / 2.0f;
            S.st[WS_TRANS][t] = Tacc; S.st[WS_CDFPREV][t] = acc;
            S.st[WS_PACK][t] = __uint_as_float((unsigned)WF_FINE | (j << 3) | (i << 6) | (ey << 16));
            int idx = atomicAdd(&S.n_q, 1);
            S.qlist[idx] = (unsigned short)t;
            S.qmask[t] = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) S.qx[k][t] = o[k] + d[k] * mid;
            c_q++;
            return;
        }
        Tfin = 1.0f - acc;
        if (GI && gi_pushed) {
            // the last fine sample's radiance arrives in this round's GI phase: finish next round
            S.st[WS_TRANS][t] = Tfin;
            S.st[WS_PACK][t] = __uint_as_float((unsigned)WF_GIWAIT | (j << 3) | (i << 6) | (ey << 16));
            return;
        }
        action = WF_ACT_FINISH;
    }
    // ---- run the state machine until it needs an SDF or the ray is finished
    bool have_q = false;
    float tq = 0.f;
    for (;;) {
        if (action == WF_ACT_NEXT) {
            bool cont;
            if (m.next(p.occ_bits, p.occ_res, ts, te, cont)) { tq = ts; have_q = true; break; }
            if (stage == WF_CDF) { action = WF_ACT_FINE_START; continue; }
            Tfin = 1.0f;  // no crossing: fully visible
            action = WF_ACT_FINISH;
            continue;
        }
        if (action == WF_ACT_CDF) {
            bool need_next = false;
            while (j < 5) {
                if (cdf_u < cdf_next) {
                    float scaling = (ce - cs) / (cdf_next - cdf_prev);
                    float tt = (cdf_u - cdf_prev) * scaling + cs;
                    S.st[WS_TPL + j][t] = tt;
                    cdf_u += cdf_step_size;
                    j += 1;
                } else if (pending) {
                    cs = ts; ce = te;
                    pending = false;
                    float a = ia_alpha(sdf_cur, ce - cs, p.beta);
                    float weight = trans * a;
                    trans *= (1.0f - a);
                    cdf_prev = cdf_next;
                    cdf_next += weight;
                } else {
                    need_next = true;
                    break;
                }
            }
            stage = WF_CDF;
            action = need_next ? WF_ACT_NEXT : WF_ACT_FINE_START;
            continue;
        }
        if (action == WF_ACT_FINE_START) {
            i = 0;
            if (j >= 2) {
                float s1 = S.st[WS_TPL][t], e1 = S.st[WS_TPL + 1][t];
                tq = (s1 + e1) / 2.0f;
                trans = 1.0f;      // Tacc
                cdf_prev = 0.0f;   // acc
                if (GI) { S.st[WS_IND][t] = 0.f; S.st[WS_IND + 1][t] = 0.f; S.st[WS_IND + 2][t] = 0.f; }
                stage = WF_FINE;
                have_q = true;
                break;
            }
            Tfin = 1.0f;
            action = WF_ACT_FINISH;
            continue;
        }
        // WF_ACT_FINISH: the estimator runs in the refill pass; remember the transmittance and whether the slot carries
        // indirect radiance (stage FINE)
        S.st[WS_TRANS][t] = Tfin;
        S.st[WS_PACK][t] = __uint_as_float((unsigned)(GI && stage == WF_FINE ? WF_DONE_IND : WF_DONE) | (ey << 16));
        S.rlist[atomicAdd(&S.n_rlist, 1)] = (unsigned short)t;
        return;
    }
    // have_q
#pragma unroll
    for (int k = 0; k < 3; k++) S.qx[k][t] = o[k] + d[k] * tq;
    wf_store_marcher(S, t, m);
    S.st[WS_SDFPREV][t] = sdf_prev; S.st[WS_CS][t] = cs; S.st[WS_CE][t] = ce;
    S.st[WS_TS][t] = ts; S.st[WS_TE][t] = te;
    S.st[WS_TRANS][t] = trans; S.st[WS_CDFPREV][t] = cdf_prev; S.st[WS_CDFNEXT][t] = cdf_next; S.st[WS_CDFU][t] = cdf_u;
    int idx = atomicAdd(&S.n_q, 1);
    S.qlist[idx] = (unsigned short)t;
    S.qmask[t] = 0;
    c_q++;
    S.st[WS_PACK][t] = __uint_as_float((unsigned)stage | (j << 3) | (i << 6) | (ey << 16));
}

template <bool GI, class P>
__device__ __forceinline__ void wf_advance_refill(const IaFrame& p, P& pol, WfShared& S, const int t, int ring_tail,
                                                  unsigned& c_q, unsigned& c_rays) {
    const unsigned pack = __float_as_uint(S.st[WS_PACK][t]);
    const int stage0 = pack & 7u;
    if (stage0 == WF_DONE || stage0 == WF_DONE_IND) {
        const uint2 entry = make_uint2(__float_as_uint(S.st[WS_ID][t]), pack >> 16);
        const float d[3] = {S.st[WS_D][t], S.st[WS_D + 1][t], S.st[WS_D + 2][t]};
        float ind[3] = {0.f, 0.f, 0.f};
        if (GI && stage0 == WF_DONE_IND) { ind[0] = S.st[WS_IND][t]; ind[1] = S.st[WS_IND + 1][t]; ind[2] = S.st[WS_IND + 2][t]; }
        wf_finish(pol, S, entry, S.st[WS_TRANS][t], ind, d);
    }
    for (int tries = 0; tries < 4; tries++) {
        int h = atomicAdd(&S.ring_head, 1);
        if (h >= ring_tail) { atomicSub(&S.ring_head, 1); break; }
        const uint2 entry = S.ring[h & (WF_QCAP - 1)];
        float o[3], d[3];
        pol.init(entry, o, d);
        IaMarcher m;
        m.init(p, o, d, p.sec_near, p.sec_far, p.sec_step);
        c_rays++;
        float ts, te;
        bool cont;
        if (m.next(p.occ_bits, p.occ_res, ts, te, cont)) {
            const float tq = p.sec_mode == IA_SEC_PLAIN ? (ts + te) / 2.0f : ts;
#pragma unroll
            for (int k = 0; k < 3; k++) { S.st[WS_O + k][t] = o[k]; S.st[WS_D + k][t] = d[k]; S.qx[k][t] = o[k] + d[k] * tq; }
            S.st[WS_ID][t] = __uint_as_float(entry.x);
            wf_store_marcher(S, t, m);
            S.st[WS_TS][t] = ts; S.st[WS_TE][t] = te;
            int idx = atomicAdd(&S.n_q, 1);
            S.qlist[idx] = (unsigned short)t;
            S.qmask[t] = 0;
            c_q++;
            S.st[WS_PACK][t] = __uint_as_float((unsigned)WF_FIRST | (entry.y << 16));
            return;
        }
        // the ray meets no occupied cell: fully visible, no indirect radiance
        const float zero[3] = {0.f, 0.f, 0.f};
        wf_finish(pol, S, entry, 1.0f, zero, d);
    }
    S.st[WS_PACK][t] = __uint_as_float((unsigned)WF_IDLE);
}

// shared memory + scratch set-up of a wavefront CTA (weights, fragments, tiles, per-CTA scratch pointers, empty ray slots)
template <int SHADE_WARPS>
__device__ __forceinline__ void wf_setup(const IaFrame& p, WfShared& S, unsigned char* __restrict__ scratch) {
    const int tid = threadIdx.x;
    const int n_w = IA_GEOC_END;
    if (tid == 0) {
        unsigned char* mine = scratch + (size_t)blockIdx.x * WF_SCRATCH_BYTES;
        S.w = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&S) + ((sizeof(WfShared) + 15) & ~15));
        S.w1f = reinterpret_cast<float4*>(S.w + n_w);
        S.xs = S.w + n_w + WF_W1F_FLOATS;
        S.cand = reinterpret_cast<float*>(mine);
        S.csdf = reinterpret_cast<float*>(mine + WF_OFF_CSDF);
        S.gtask = reinterpret_cast<unsigned short*>(mine + WF_OFF_GTASK);
#if !WF_BTASK_SMEM
        S.btask = reinterpret_cast<unsigned short*>(mine + WF_OFF_BTASK);
#endif
        S.gitask = reinterpret_cast<uint2*>(mine + WF_OFF_GITASK);
        S.gixc = reinterpret_cast<float*>(mine + WF_OFF_GIXC);
        S.flist = reinterpret_cast<float*>(mine + WF_OFF_FLIST);
#if WF_STATE_GLOBAL
        S.st = reinterpret_cast<float (*)[WF_R]>(mine + WF_OFF_STATE);
#endif
    }
    __syncthreads();
    ia_stage_geoc(S.w, p.mlp);
#if WF_W1F_SMEM
    ia_stage_bfrag(S.w1f, IA_GEO_KSTEPS, 8, [&](int k, int n) { return ia_geo_w1(p.mlp, k, n); });
#endif
    for (int i = tid; i < (int)(WF_TILE_FLOATS_SW(SHADE_WARPS)); i += blockDim.x) S.xs[i] = 0.f;
    if (tid < IA_N_INIT * 12) S.tfs13[tid] = p.tfs[p.init_bones[tid / 12]][tid % 12];
    if (tid < IA_N_LEVELS) S.lvl[tid] = ia_level(p, tid);
    for (int t = tid; t < WF_R; t += blockDim.x) S.st[WS_PACK][t] = __uint_as_float(0u);
    if (tid == 0) { S.ring_head = 0; S.ring_tail = 0; S.more_tiles = 1; S.n_q = 0; S.n_gtask = 0; S.task_next = 0; }
    if (tid < 8) S.cnt[tid] = 0;
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
template <bool GI, class P>
__device__ __forceinline__ void wf_run(const IaFrame& p, P& pol, WfShared& S, unsigned char* __restrict__ scratch,
                                       unsigned long long* __restrict__ counters) {
    const int tid = threadIdx.x;
    wf_setup<(GI ? WF_GI_WARPS : 0)>(p, S, scratch);
    const int tile_items = pol.tile_items();
    const long long n_tiles = (pol.n_items() + tile_items - 1) / tile_items;
    while (true) {
        // ---- feed the ring while it cannot fill every slot
        while (true) {
            const int cnt = S.ring_tail - S.ring_head;
            const int more = S.more_tiles;
            if (cnt >= WF_R || !more) break;
            __syncthreads();
            if (tid == 0) {
                int tl = atomicAdd(pol.tile_counter(), 1);
                S.tile = tl;
                if (tl >= n_tiles) S.more_tiles = 0;
            }
            __syncthreads();
            const int tl = S.tile;
            if (tl < n_tiles) pol.feed((long long)tl * tile_items, S);
            __syncthreads();
        }
        __syncthreads();
        if (tid == 0) { S.n_q = 0; S.n_gtask = 0; S.task_next = 0; S.n_btask = 0; S.n_gitask = 0; S.n_rlist = 0; S.n_flist = 0; }
        const int ring_tail = S.ring_tail;
        __syncthreads();
        {
            unsigned c_q = 0, c_rays = 0;
            // (asking for the state lines of a warp's slots all at once with prefetch.global.L1 in front of this loop changes
            //  nothing: 752.9 vs 752.6 ms -- the pass is not bound by its dependent L2 round trips)
            for (int t = tid; t < WF_R; t += blockDim.x) wf_advance_consume<GI>(p, pol, S, t, c_q);
            __syncthreads();
            const int n_r = S.n_rlist;
            for (int k = tid; k < n_r; k += blockDim.x) wf_advance_refill<GI>(p, pol, S, S.rlist[k], ring_tail, c_q, c_rays);
            wf_count(S, WF_C_Q, c_q);
            wf_count(S, WF_C_RAYS, c_rays);
        }
        __syncthreads();
        const int n_gi = GI ? S.n_gitask : 0;
        const int n_q = S.n_q;
        if (n_q == 0) {
            wf_finish_phase(pol, S);
            if (n_gi) wf_gi_phase(p, S);
            // (GI: rays waiting for their last radiance, WF_GIWAIT, need one more advance round)
            if (n_gi == 0 && S.ring_tail - S.ring_head <= 0 && !S.more_tiles) break;
            continue;
        }
#if WF_QSORT
        wf_sort_phase(p, S, n_q);
        __syncthreads();
#endif
        wf_prune_phase(p, S, n_q);
        __syncthreads();
        // The GI phase (static share per warp, uneven in time) runs WITHOUT a barrier in front of the Broyden phase, whose
        // tasks are grabbed dynamically: a warp that finishes its radiance batches early takes more chains instead of
        // waiting (the barrier behind the GI phase was 2.5 % of the kernel's stall samples).  The two phases share no
        // data: the GI tasks carry their own copy of the root.
        wf_finish_phase(pol, S);   // (deferred estimators of the rays that ended this round: same reasoning)
        if (n_gi) wf_gi_phase(p, S);
        wf_broyden_phase(p, S);
        __syncthreads();
        wf_filter_phase(S, n_q);
        __syncthreads();
        wf_geometry_phase(p, S);
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned long long q = S.cnt[WF_C_Q], f = S.cnt[WF_C_FETCH], g = S.cnt[WF_C_GEO], r = S.cnt[WF_C_RAYS],
                                 k = S.cnt[WF_C_SKIP], qg = S.cnt[WF_C_QG];
        if (q) atomicAdd(&counters[IA_CNT_QUERIES], q);
        if (f) atomicAdd(&counters[IA_CNT_BROYDEN_FETCH], f);
        if (g + qg) atomicAdd(&counters[IA_CNT_GEO_EVAL], g + qg);
        if (r) atomicAdd(&counters[IA_CNT_SECONDARY_RAYS], r);
        if (k) atomicAdd(&counters[IA_CNT_CHAINS_SKIPPED], k);
        if (qg) {
            atomicAdd(&counters[IA_CNT_QUERIES_GRAD], qg);
            atomicAdd(&counters[IA_CNT_SKIN_FETCH], qg);
            atomicAdd(&counters[IA_CNT_RAD_EVAL], qg);
        }
    }
}

__device__ __forceinline__ void wf_ring_push(WfShared& S, bool live, uint2 e) {
    unsigned b = __ballot_sync(0xffffffffu, live);
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(&S.ring_tail, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) S.ring[(base + __popc(b & ((1u << lane) - 1))) & (WF_QCAP - 1)] = e;
}

// ------------------------------------------------------------------------------------------------
// Policy 1: the shading stage.  One policy, four integrators (config.model.render_mode):
//   IA_MODE_LIGHT          pbr_light_forward          models/intrinsic_avatar.py:755-861  (table of spp light directions)
//   IA_MODE_UNIFORM_LIGHT  pbr_uniform_light_forward  :654-753  (stratified sphere table, inv_pdf = 4 pi, visibility map)
//   IA_MODE_MATS           pbr_mats_forward           :863-948  (BSDF sampling, one ray per shading sample)
//   IA_MODE_MIS            pbr_mis_forward            :547-652  (BSDF + light sampling, two rays per shading sample)
enum { IA_MODE_LIGHT = 0, IA_MODE_UNIFORM_LIGHT = 1, IA_MODE_MATS = 2, IA_MODE_MIS = 3 };

template <int MODE>
struct WfShadePolicy {
    static constexpr int mode = MODE;   // compile-time: the default light path carries none of the other integrators' code
    const IaFrame* p;
    const int* hit_rays; const float* hit_od; const IaSample* samples;
    const float* rs_t; const int* rs_src; const float* rs_w;
    int* work; int spp; long long ray_index_base; uint32_t seed;
    const float* light_dir_s; const float* light_em; const float* light_pdf;
    float* acc6;
    long long n_total;
    int gi;
    IaEnv env;     // mats / mis: per-direction envmap look-ups
    float* vis;    // uniform_light: [n_rays] visibility accumulator
    const float* bg_rgb;  // [n_rays][3] radiance of a background-assigned sample (background colour or envmap)

    static constexpr bool light_major = WF_LIGHT_MAJOR && MODE <= IA_MODE_UNIFORM_LIGHT;
    int n_hit;     // hit rays of the frame (light-major: position e = kk * n_hit + slot)
    // hit-ray slot of a ring entry (e.x = position in the resampled streams, e.y = light index in the light-table modes)
    __device__ __forceinline__ int slot_of(const uint2 e) const {
        return light_major ? (int)(e.x - e.y * (unsigned)n_hit) : (int)(e.x / (unsigned)spp);
    }
    __device__ __forceinline__ long long n_items() const { return n_total; }
    __device__ __forceinline__ int* tile_counter() const { return &work[IA_W_TILE_NEXT]; }
    __device__ __forceinline__ int tile_items() const { return mode == IA_MODE_MIS ? WF_FEED / 2 : WF_FEED; }

    // One shading sample of the tile: background-assigned samples are accumulated on the spot; returns whether the
    // sample traces a ray, and its ring entry.
    __device__ __forceinline__ bool classify(long long s, uint2& e) const {
        e = make_uint2(0, 0);
        if (s >= n_total) return false;
        const unsigned su = (unsigned)s;
        int slot, j;
        uint32_t kk = 0;
        if (light_major) { kk = su / (unsigned)n_hit; slot = (int)(su - kk * (unsigned)n_hit); j = 0; }
        else { slot = (int)(su / (unsigned)spp); j = (int)(su % (unsigned)spp); }
        const int src = rs_src[s];
        if (src < 0) {
            // background-assigned shading sample (models/intrinsic_avatar.py:1319-1341)
            const float w = rs_w[s];
            float* pa = acc6 + (size_t)hit_rays[slot] * 6;
            const float* bg = bg_rgb + (size_t)hit_rays[slot] * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                atomicAdd(&pa[k], w * bg[k]);
                atomicAdd(&pa[3 + k], w * bg[k]);
            }
            return false;
        }
        if (mode >= IA_MODE_MATS) {
            e = make_uint2(su, 0);
            return true;  // no cosine mask: every foreground sample traces its sampled direction(s)
        }
        const float* n = samples[src].n;
        if (!light_major) {
            uint32_t key = ia_pixel_key(seed, (uint64_t)(ray_index_base + hit_rays[slot]));
            kk = ia_permute((uint32_t)j, (uint32_t)spp, key);
        }
        const float* wo = light_dir_s + kk * 3;
        float cosv = n[0] * wo[0] + n[1] * wo[1] + n[2] * wo[2];
        e = make_uint2(su, kk);
        return cosv > 1e-6f;
    }

    __device__ __forceinline__ void feed(long long s0, WfShared& S) {
        const int n_it = tile_items();
        for (int i = threadIdx.x; i < n_it; i += blockDim.x) {
            uint2 e;
            const bool live = classify(s0 + i, e);
            wf_ring_push(S, live, e);
            if (mode == IA_MODE_MIS) wf_ring_push(S, live, make_uint2(e.x, 1));  // the light-sampled ray of the pair
        }
    }

    __device__ __forceinline__ void init(const uint2 e, float o[3], float d[3]) const {
        const int slot = slot_of(e);
        const float t = rs_t[e.x];
        const float* od = hit_od + (size_t)slot * 8;
#pragma unroll
        for (int k = 0; k < 3; k++) o[k] = od[k] + od[3 + k] * t;
        if (mode >= IA_MODE_MATS) {
            const uint32_t j = e.x % (unsigned)spp;
            const uint32_t key = ia_pixel_key(seed, (uint64_t)(ray_index_base + hit_rays[slot]));
            if (e.y == 0) {
                // scatterer.sample (MultiLobe.sample, lib/torch_pbr/bxdf.py:332-388) in the SMPL frame
                const IaSample sm = samples[rs_src[e.x]];
                const float wi[3] = {-od[3], -od[4], -od[5]};
                ia_multilobe_sample(sm.n, wi, sm.rough, sm.albedo, sm.metal, ia_rng_uniform(key, j, 0),
                                    ia_rng_uniform(key, j, 1), d);
            } else {
                // emitter.sample -> transform_dirs_w2s (models/intrinsic_avatar.py:578-580)
                float dw[3];
                ia_env_sample(env, ia_rng_uniform(key, j, 2), ia_rng_uniform(key, j, 3), dw);
                ia_dir_w2s(*p, dw, d);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; k++) d[k] = light_dir_s[e.y * 3 + k];
        }
    }

    __device__ __forceinline__ void finish(const uint2 e, float T, const float ind[3], const float d[3]) const {
        const int slot = slot_of(e);
        const IaSample sm = samples[rs_src[e.x]];
        const float w = rs_w[e.x];
        const float* od = hit_od + (size_t)slot * 8;
        const float wi[3] = {-od[3], -od[4], -od[5]};
        float* pa = acc6 + (size_t)hit_rays[slot] * 6;
        float diff, spec[3];
        if (mode >= IA_MODE_MATS) {
            const float wo[3] = {d[0], d[1], d[2]};
            ia_brdf_multilobe(wi, sm.n, wo, sm.rough, sm.albedo, sm.metal, diff, spec);
            const float pdf_s = ia_multilobe_pdf(wi, sm.n, wo, sm.rough, sm.albedo, sm.metal);
            float dw[3], em[3];
            ia_dir_s2w(*p, wo, dw);
            ia_env_eval(env, dw, em);
            float pdf = 1.0f, misw = 0.f;
            if (mode == IA_MODE_MIS) {
                const float sum = pdf_s + ia_env_pdf(env, dw);
                misw = sum > 1e-6f ? 1.0f / sum : 0.f;
            } else {
                pdf = pdf_s > 0.f ? pdf_s : 1.0f;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float Li = em[k] * T;  // (transmittance is not clamped on these two paths)
                if (gi) Li += ind[k];
                float Ld, Ls;
                if (mode == IA_MODE_MIS) { Ld = (Li * diff) * misw; Ls = (Li * spec[k]) * misw; }
                else { Ld = Li * diff / pdf; Ls = Li * spec[k] / pdf; }
                float kd = (1.0f - sm.metal) * sm.albedo[k];
                atomicAdd(&pa[k], w * (kd * Ld + Ls));
                atomicAdd(&pa[3 + k], w * (Ld + Ls));
            }
            return;
        }
        const unsigned kk = e.y;
        const float wo[3] = {light_dir_s[kk * 3], light_dir_s[kk * 3 + 1], light_dir_s[kk * 3 + 2]};
        float tr = fminf(fmaxf(T, 0.f), 1.f);
        ia_brdf_multilobe(wi, sm.n, wo, sm.rough, sm.albedo, sm.metal, diff, spec);
        bool lit = tr > 0.0f;
        if (mode == IA_MODE_UNIFORM_LIGHT) {
            const float inv_pdf = 4.0f * 3.14159265358979323846f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float em = lit ? light_em[kk * 3 + k] : 0.f;
                float Li = em * tr;
                if (gi) Li += ind[k];
                float Ld = Li * diff * inv_pdf, Ls = Li * spec[k] * inv_pdf;
                float kd = (1.0f - sm.metal) * sm.albedo[k];
                atomicAdd(&pa[k], w * (kd * Ld + Ls));
                atomicAdd(&pa[3 + k], w * (Ld + Ls));
            }
            if (vis) atomicAdd(&vis[hit_rays[slot]], w * (2.0f * tr));
            return;
        }
        float pdf = lit ? light_pdf[kk] : 1.0f;
        if (!(pdf > 0)) pdf = 1.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float em = lit ? light_em[kk * 3 + k] : 0.f;
            float Li = em * tr;
            if (gi) Li += ind[k];
            float Ld = Li * diff / pdf, Ls = Li * spec[k] / pdf;
            float kd = (1.0f - sm.metal) * sm.albedo[k];
            atomicAdd(&pa[k], w * (kd * Ld + Ls));
            atomicAdd(&pa[3 + k], w * (Ld + Ls));
        }
    }
};

#define WF_SMEM_BYTES_SW(SW) (((sizeof(WfShared) + 15) & ~(size_t)15) + IA_GEOC_END * sizeof(float) + \
                              WF_W1F_FLOATS * sizeof(float) + WF_TILE_FLOATS_SW(SW) * sizeof(float))
#define WF_SMEM_BYTES(GI) WF_SMEM_BYTES_SW((GI) ? WF_GI_WARPS : 0)

template <bool GI, int MODE>
__global__ void __launch_bounds__(WF_THREADS, WF_CTAS_PER_SM) k_shade_wf(const __grid_constant__ IaFrame p, WfShadePolicy<MODE> pol,
                                                            unsigned char* __restrict__ scratch,
                                                            unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char wf_smem[];
    WfShared& S = *reinterpret_cast<WfShared*>(wf_smem);
    pol.p = &p;
    pol.n_hit = pol.work[IA_W_NHIT];
    pol.n_total = (long long)pol.n_hit * pol.spp;
    pol.gi = GI ? 1 : 0;
    wf_run<GI>(p, pol, S, scratch, counters);
}

// Policy 2: op-level secondary rays (ia_op_secondary, gi = 0)
struct WfRaysPolicy {
    const float* ro; const float* rd; long long n; float* T_out; float* rgb_out; int* work;
    __device__ __forceinline__ long long n_items() const { return n; }
    __device__ __forceinline__ int* tile_counter() const { return &work[IA_W_TILE_NEXT]; }
    __device__ __forceinline__ int tile_items() const { return WF_FEED; }
    __device__ __forceinline__ void feed(long long s0, WfShared& S) {
        for (int i = threadIdx.x; i < WF_FEED; i += blockDim.x) {
            const long long s = s0 + i;
            wf_ring_push(S, s < n, make_uint2((unsigned)s, 0));
        }
    }
    __device__ __forceinline__ void init(const uint2 e, float o[3], float d[3]) const {
#pragma unroll
        for (int k = 0; k < 3; k++) { o[k] = ro[(size_t)e.x * 3 + k]; d[k] = rd[(size_t)e.x * 3 + k]; }
    }
    __device__ __forceinline__ void finish(const uint2 e, float T, const float ind[3], const float*) const {
        T_out[e.x] = T;
        if (rgb_out) { rgb_out[(size_t)e.x * 3] = ind[0]; rgb_out[(size_t)e.x * 3 + 1] = ind[1]; rgb_out[(size_t)e.x * 3 + 2] = ind[2]; }
    }
};

template <bool GI>
__global__ void __launch_bounds__(WF_THREADS, WF_CTAS_PER_SM) k_rays_wf(const __grid_constant__ IaFrame p, WfRaysPolicy pol,
                                                           unsigned char* __restrict__ scratch,
                                                           unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char wf_smem[];
    WfShared& S = *reinterpret_cast<WfShared*>(wf_smem);
    wf_run<GI>(p, pol, S, scratch, counters);
}
