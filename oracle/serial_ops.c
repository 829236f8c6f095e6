/*
 * ORACLE (test infrastructure, not product) -- plain C restatement of the reference's
 * one-thread-per-ray kernels.  Each CUDA thread body becomes one loop iteration here; float /
 * double mixing in the arithmetic is kept exactly as written in the reference so the results are
 * bit-comparable with the reference kernels compiled without fast-math.
 *
 *   cdf_resampling          <- lib/nerfacc/cuda/csrc/cdf.cu:9-149   (ray_resampling)
 *   cdf_resampling_merge    <- lib/nerfacc/cuda/csrc/cdf.cu:217-334 (ray_resampling_merge)
 *   cdf_resampling_fine     <- lib/nerfacc/cuda/csrc/cdf.cu:403-478 (ray_resampling_fine)
 *   cdf_resampling_sdf_fine <- lib/nerfacc/cuda/csrc/cdf.cu:536-638 (ray_resampling_sdf_fine)
 *   unpack_info / unpack_data <- lib/nerfacc/cuda/csrc/pack.cu:7-28, 54-82
 *   traverse_grid           <- nerfacc==0.5.3 `traverse_grids` (NOT under /root/reference;
 *                              requirements.txt:8). Restated from the published algorithm as
 *                              recalled in SURVEY.md Appendix B: parity unpinned.
 *                              Call sites: models/intrinsic_avatar.py:84-93.
 *   render_weight_from_alpha <- nerfacc==0.5.3 (same status); call sites
 *                              models/intrinsic_avatar.py:1199, models/volrend.py:162,952.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC serial_ops.c -o _build/liboracle_serial.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef unsigned char bool8;

/* ------------------------------------------------------------------------------------------ */
void cdf_resampling(int n_rays, const int *packed_info, const float *starts, const float *ends,
                    const float *weights_all, const float *sdfs_all, const int *resample_packed_info,
                    float *resample_ts_all, float *resample_offsets_all, int64_t *surface_idx,
                    int64_t *resample_indices_all, int32_t *resample_fg_counts_all,
                    int32_t *resample_bg_counts) {
    for (int i = 0; i < n_rays; i++) {
        const int base = packed_info[i * 2 + 0];
        const int steps = packed_info[i * 2 + 1];
        const int resample_base = resample_packed_info[i * 2 + 0];
        const int resample_steps = resample_packed_info[i * 2 + 1];
        if (steps == 0) continue;
        const float *st = starts + base, *en = ends + base, *weights = weights_all + base,
                    *sdfs = sdfs_all + base;
        int32_t *fg = resample_fg_counts_all + base;
        float *resample_ts = resample_ts_all + resample_base;
        float *resample_offsets = resample_offsets_all + resample_base;
        int64_t *resample_indices = resample_indices_all + resample_base;

        float weights_sum = 0.0f;
        for (int j = 0; j < steps; j++) weights_sum += weights[j];
        weights_sum += fmaxf(1.0f - weights_sum, 0.0f);

        int num_bins = resample_steps;
        float cdf_step_size = (float)((1.0f - 1.0 / num_bins) / (resample_steps - 1));
        int idx = 0, j = 0;
        float cdf_prev = 0.0f, cdf_next = weights[idx] / weights_sum;
        float cdf_u = (float)(1.0 / (2 * num_bins));
        float sdf_prev = sdfs[0];
        float sdf_next = 0.0f;
        if (steps > 1) sdf_next = sdfs[1];
        int found_surface = 0;
        while (j < num_bins && idx < steps) {
            if (cdf_u < cdf_next) {
                float scaling = (en[idx] - st[idx]) / (cdf_next - cdf_prev);
                float offset = (cdf_u - cdf_prev) * scaling;
                float t = offset + st[idx];
                if (sdf_prev >= 0 && sdf_next < 0 && !found_surface) {
                    float sdf_approx = sdf_prev + (sdf_next - sdf_prev) * (offset / (en[idx] - st[idx]));
                    resample_ts[j] = sdf_approx >= 0 ? t : (j > 0 ? resample_ts[j - 1] : st[idx]);
                } else if (found_surface) {
                    resample_ts[j] = j > 0 ? resample_ts[j - 1] : st[idx];
                } else {
                    resample_ts[j] = t;
                }
                resample_offsets[j] = offset;
                resample_indices[j] = idx + base;
                fg[idx] += 1;
                cdf_u += cdf_step_size;
                j += 1;
            } else if (idx < steps - 1) {
                idx += 1;
                if (sdf_prev >= 0 && sdf_next < 0 && !found_surface) {
                    surface_idx[i] = idx - 1 + base;
                    found_surface = 1;
                }
                sdf_prev = sdfs[idx];
                sdf_next = idx < steps - 1 ? sdfs[idx + 1] : 0.0f;
                cdf_prev = cdf_next;
                cdf_next += weights[idx] / weights_sum;
            } else {
                break;
            }
        }
        while (j < num_bins) {
            float offset = 10000.f;
            float t = offset + en[steps - 1];
            resample_ts[j] = t;
            resample_offsets[j] = offset;
            resample_indices[j] = steps - 1 + base;
            cdf_u += cdf_step_size;
            j += 1;
            resample_bg_counts[i] += 1;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
void cdf_resampling_merge(int n_rays, const int *packed_info, const float *vals_all,
                          const bool8 *is_left_all, const bool8 *is_right_all, const float *weights_all,
                          const int *resample_packed_info, float *resample_vals_all,
                          float *resample_dists_all, bool8 *resample_is_left_all,
                          bool8 *resample_is_right_all, bool8 *is_resample_all, bool8 *is_fg_sample_all) {
    for (int i = 0; i < n_rays; i++) {
        const int base = packed_info[i * 2 + 0];
        const int steps = packed_info[i * 2 + 1];
        const int resample_base = resample_packed_info[i * 2 + 0];
        const int resample_steps = resample_packed_info[i * 2 + 1] - steps;
        if (steps == 0) continue;
        const float *vals = vals_all + base, *weights = weights_all + base;
        const bool8 *is_left = is_left_all + base, *is_right = is_right_all + base;
        bool8 *is_fg_sample = is_fg_sample_all + resample_base;
        float *resample_vals = resample_vals_all + resample_base;
        float *resample_dists = resample_dists_all + resample_base;
        bool8 *resample_is_left = resample_is_left_all + resample_base;
        bool8 *resample_is_right = resample_is_right_all + resample_base;
        bool8 *is_resample = is_resample_all + resample_base;

        float weights_sum = 0.0f;
        for (int j = 0; j < steps - 1; j++)
            weights_sum += (is_left[j] && is_right[j + 1]) ? weights[j] : 0.0f;
        weights_sum += fmaxf(1.0f - weights_sum, 0.0f);

        int num_bins = resample_steps;
        float cdf_step_size = (float)((1.0f - 1.0 / num_bins) / (resample_steps - 1));
        int idx = 0, j = 0;
        float start = 0.0f, end = 0.0f;
        float cdf_prev = 0.0f, cdf_next = weights[idx] / weights_sum;
        float cdf_u = (float)(1.0 / (2 * num_bins));
        start = vals[0];
        end = vals[1];
        resample_vals[0] = start;
        is_fg_sample[0] = 1;
        resample_is_left[0] = 1;
        while (j < num_bins && idx < steps - 1) {
            if (cdf_u < cdf_next) {
                float scaling = (end - start) / (cdf_next - cdf_prev);
                float offset = (cdf_u - cdf_prev) * scaling;
                float t = offset + start;
                cdf_u += cdf_step_size;
                resample_dists[j + idx] = t - resample_vals[j + idx];
                j += 1;
                resample_vals[j + idx] = t;
                is_fg_sample[j + idx] = 1;
                is_resample[j + idx] = 1;
                resample_is_left[j + idx] = 1;
                resample_is_right[j + idx] = 1;
            } else {
                resample_dists[j + idx] = end - resample_vals[j + idx];
                idx += 1;
                resample_vals[j + idx] = end;
                is_fg_sample[j + idx] = 1;
                resample_is_right[j + idx] = is_right[idx];
                if (idx >= steps - 1) break;
                start = vals[idx];
                end = vals[idx + 1];
                if (is_left[idx] && is_right[idx + 1]) {
                    cdf_prev = cdf_next;
                    cdf_next += weights[idx] / weights_sum;
                    resample_is_left[j + idx] = 1;
                }
            }
        }
        while (idx < steps - 1) {
            resample_dists[j + idx] = end - resample_vals[j + idx];
            idx += 1;
            resample_vals[j + idx] = end;
            is_fg_sample[j + idx] = 1;
            resample_is_right[j + idx] = is_right[idx];
            if (idx >= steps - 1) break;
            start = vals[idx];
            end = vals[idx + 1];
            if (is_left[idx] && is_right[idx + 1]) resample_is_left[j + idx] = 1;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
void cdf_resampling_fine(int n_rays, const int *packed_info, const float *starts, const float *ends,
                         const float *weights_all, const int *resample_packed_info,
                         float *resample_starts_all, float *resample_ends_all, bool8 *is_fg_sample_all) {
    for (int i = 0; i < n_rays; i++) {
        const int base = packed_info[i * 2 + 0];
        const int steps = packed_info[i * 2 + 1];
        const int resample_base = resample_packed_info[i * 2 + 0];
        const int resample_steps = resample_packed_info[i * 2 + 1];
        if (steps == 0) continue;
        const float *st = starts + base, *en = ends + base, *weights = weights_all + base;
        float *resample_starts = resample_starts_all + resample_base;
        float *resample_ends = resample_ends_all + resample_base;
        bool8 *is_fg_sample = is_fg_sample_all + resample_base;

        float weights_sum = 0.0f;
        for (int j = 0; j < steps; j++) weights_sum += weights[j];
        weights_sum += fmaxf(1.0f - weights_sum, 0.0f);
        int num_bins = resample_steps + 1;
        float cdf_step_size = (float)((1.0f - 1.0 / num_bins) / resample_steps);
        int idx = 0, j = 0;
        float cdf_prev = 0.0f, cdf_next = weights[idx] / weights_sum;
        float cdf_u = (float)(1.0 / (2 * num_bins));
        while (j < num_bins && idx < steps) {
            if (cdf_u < cdf_next) {
                float scaling = (en[idx] - st[idx]) / (cdf_next - cdf_prev);
                float t = (cdf_u - cdf_prev) * scaling + st[idx];
                if (j < num_bins - 1) resample_starts[j] = t;
                if (j > 0) {
                    resample_ends[j - 1] = t;
                    is_fg_sample[j - 1] = 1;
                }
                cdf_u += cdf_step_size;
                j += 1;
            } else {
                idx += 1;
                if (idx >= steps) break;
                cdf_prev = cdf_next;
                cdf_next += weights[idx] / weights_sum;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
void cdf_resampling_sdf_fine(int n_rays, const int *packed_info, const float *starts, const float *ends,
                             const float *alphas_all, const float *sdfs_all,
                             const int *resample_packed_info, float *resample_starts_all,
                             float *resample_ends_all, bool8 *is_fg_sample_all) {
    for (int i = 0; i < n_rays; i++) {
        const int base = packed_info[i * 2 + 0];
        const int steps = packed_info[i * 2 + 1];
        const int resample_base = resample_packed_info[i * 2 + 0];
        const int resample_steps = resample_packed_info[i * 2 + 1];
        if (steps == 0) continue;
        const float *st = starts + base, *en = ends + base, *alphas = alphas_all + base,
                    *sdfs = sdfs_all + base;
        float *resample_starts = resample_starts_all + resample_base;
        float *resample_ends = resample_ends_all + resample_base;
        bool8 *is_fg_sample = is_fg_sample_all + resample_base;

        int idx = 0;
        float sdf_prev = sdfs[0];
        int found_surface = 0;
        while (idx < steps) {
            idx += 1;
            if (idx >= steps) break;
            if (sdf_prev >= 0 && sdfs[idx] < 0 && !found_surface) {
                idx -= 1;
                found_surface = 1;
                break;
            }
            sdf_prev = sdfs[idx];
        }
        if (!found_surface) continue;

        int num_bins = resample_steps + 1;
        float cdf_step_size = (float)((1.0f - 1.0 / num_bins) / resample_steps);
        int j = 0;
        float trans = 1.0f;
        float weight = alphas[idx];
        trans *= (1.0f - alphas[idx]);
        float cdf_prev = 0.0f, cdf_next = weight;
        float cdf_u = (float)(1.0 / (2 * num_bins));
        while (j < num_bins && idx < steps) {
            if (cdf_u < cdf_next) {
                float scaling = (en[idx] - st[idx]) / (cdf_next - cdf_prev);
                float t = (cdf_u - cdf_prev) * scaling + st[idx];
                if (j < num_bins - 1) resample_starts[j] = t;
                if (j > 0) {
                    resample_ends[j - 1] = t;
                    is_fg_sample[j - 1] = 1;
                }
                cdf_u += cdf_step_size;
                j += 1;
            } else {
                idx += 1;
                if (idx >= steps) break;
                weight = trans * alphas[idx];
                trans *= (1.0f - alphas[idx]);
                cdf_prev = cdf_next;
                cdf_next += weight;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
void unpack_info(int n_rays, const int *packed_info, int64_t *ray_indices) {
    for (int i = 0; i < n_rays; i++) {
        const int base = packed_info[i * 2 + 0], steps = packed_info[i * 2 + 1];
        for (int j = 0; j < steps; ++j) ray_indices[base + j] = i;
    }
}

void unpack_data_f32(int n_rays, const int *packed_info, int data_dim, const float *data,
                     int n_per_ray, float *unpacked) {
    for (int i = 0; i < n_rays; i++) {
        const int base = packed_info[i * 2 + 0], steps = packed_info[i * 2 + 1];
        for (int j = 0; j < steps; j++)
            for (int k = 0; k < data_dim; k++)
                unpacked[((int64_t)i * n_per_ray + j) * data_dim + k] = data[((int64_t)base + j) * data_dim + k];
    }
}

/* per-ray exclusive product of (1-alpha): weights = T*alpha, trans = T */
void render_weight_from_alpha(int n_rays, const int *packed_info, const float *alphas, float *weights,
                              float *trans) {
    for (int i = 0; i < n_rays; i++) {
        const int base = packed_info[i * 2 + 0], steps = packed_info[i * 2 + 1];
        float T = 1.0f;
        for (int j = 0; j < steps; j++) {
            trans[base + j] = T;
            weights[base + j] = T * alphas[base + j];
            T *= (1.0f - alphas[base + j]);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* nerfacc 0.5.3 traverse_grids, single level, cone_angle = 0.
 * Pass 1 (vals == NULL): counts only -> n_edges[i], n_samples[i].
 * Pass 2: writes edges at edge_base[i].., samples at sample_base[i]...                      */
static int ray_aabb(const float *o, const float *d, const float *aabb, float *tmin_o, float *tmax_o) {
    float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float tmin, tmax, tmin_t, tmax_t;
    if (inv[0] >= 0) { tmin = (aabb[0] - o[0]) * inv[0]; tmax = (aabb[3] - o[0]) * inv[0]; }
    else             { tmin = (aabb[3] - o[0]) * inv[0]; tmax = (aabb[0] - o[0]) * inv[0]; }
    if (inv[1] >= 0) { tmin_t = (aabb[1] - o[1]) * inv[1]; tmax_t = (aabb[4] - o[1]) * inv[1]; }
    else             { tmin_t = (aabb[4] - o[1]) * inv[1]; tmax_t = (aabb[1] - o[1]) * inv[1]; }
    if (tmin > tmax_t || tmin_t > tmax) return 0;
    if (tmin_t > tmin) tmin = tmin_t;
    if (tmax_t < tmax) tmax = tmax_t;
    if (inv[2] >= 0) { tmin_t = (aabb[2] - o[2]) * inv[2]; tmax_t = (aabb[5] - o[2]) * inv[2]; }
    else             { tmin_t = (aabb[5] - o[2]) * inv[2]; tmax_t = (aabb[2] - o[2]) * inv[2]; }
    if (tmin > tmax_t || tmin_t > tmax) return 0;
    if (tmin_t > tmin) tmin = tmin_t;
    if (tmax_t < tmax) tmax = tmax_t;
    if (tmax <= 0) return 0;
    *tmin_o = tmin;
    *tmax_o = tmax;
    return 1;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void traverse_grid(int n_rays, const float *rays_o, const float *rays_d, const bool8 *binaries,
                   int res, const float *aabb, float near_plane, float far_plane, float step_size,
                   int *n_edges, int *n_samples, const int *edge_base, const int *sample_base,
                   float *vals, bool8 *is_left, bool8 *is_right, float *t_starts, float *t_ends) {
    const float eps = 1e-6f;
    const int write = vals != 0;
    for (int i = 0; i < n_rays; i++) {
        const float *o = rays_o + 3 * i, *d = rays_d + 3 * i;
        int ne = 0, ns = 0;
        float tmin, tmax;
        int eb = write ? edge_base[i] : 0, sb = write ? sample_base[i] : 0;
        if (ray_aabb(o, d, aabb, &tmin, &tmax)) {
            float this_tmin = fmaxf(tmin, near_plane), this_tmax = fminf(tmax, far_plane);
            if (this_tmin < this_tmax) {
                float t_last = near_plane;
                int continuous = 0;
                /* setup_traversal */
                float voxel[3], inv[3], tdist[3], delta[3];
                int cur[3], step[3], fin[3];
                for (int k = 0; k < 3; k++) {
                    voxel[k] = (aabb[3 + k] - aabb[k]) / (float)res;
                    inv[k] = 1.0f / d[k];
                    float rs = o[k] + d[k] * (this_tmin + eps);
                    float re = o[k] + d[k] * (this_tmax - eps);
                    cur[k] = clampi((int)(((rs - aabb[k]) / (aabb[3 + k] - aabb[k])) * (float)res), 0, res - 1);
                    fin[k] = clampi((int)(((re - aabb[k]) / (aabb[3 + k] - aabb[k])) * (float)res), 0, res - 1);
                    int idelta = d[k] > 0 ? 1 : 0;
                    float tm = ((aabb[k] + ((float)(cur[k] + idelta) * voxel[k]) - rs) * inv[k]) + this_tmin;
                    tdist[k] = (d[k] == 0.0f) ? this_tmax : tm;
                    float sf = (d[k] == 0.0f) ? 0.0f : (d[k] > 0.0f ? 1.0f : -1.0f);
                    step[k] = (int)sf;
                    delta[k] = (d[k] == 0.0f) ? this_tmax : voxel[k] * inv[k] * sf;
                }
                int over[3] = {fin[0] + step[0], fin[1] + step[1], fin[2] + step[2]};
                while (1) {
                    float t_trav = fminf(tdist[0], fminf(tdist[1], tdist[2]));
                    t_trav = fminf(t_trav, this_tmax);
                    int64_t cell = ((int64_t)cur[0] * res + cur[1]) * res + cur[2];
                    if (!binaries[cell]) {
                        while (1) {
                            float dt = step_size;
                            if (t_last + dt * 0.5f >= t_trav) break;
                            t_last += dt;
                        }
                        continuous = 0;
                    } else {
                        while (1) {
                            float dt = step_size;
                            if (t_last + dt * 0.5f >= t_trav) break;
                            float t_next = t_last + dt;
                            if (!continuous) {
                                if (write) { vals[eb + ne] = t_last; is_left[eb + ne] = 1; }
                                ne++;
                                if (write) { vals[eb + ne] = t_next; is_right[eb + ne] = 1; }
                                ne++;
                            } else {
                                if (write) { vals[eb + ne] = t_next; is_left[eb + ne - 1] = 1; is_right[eb + ne] = 1; }
                                ne++;
                            }
                            if (write) { t_starts[sb + ns] = t_last; t_ends[sb + ns] = t_next; }
                            ns++;
                            continuous = 1;
                            t_last = t_next;
                            if (t_next >= t_trav) break;
                        }
                    }
                    /* single_traversal */
                    int ok = 1;
                    if (tdist[0] < tdist[1] && tdist[0] < tdist[2]) {
                        cur[0] += step[0]; tdist[0] += delta[0];
                        if (cur[0] == over[0]) ok = 0;
                    } else if (tdist[1] < tdist[2]) {
                        cur[1] += step[1]; tdist[1] += delta[1];
                        if (cur[1] == over[1]) ok = 0;
                    } else {
                        cur[2] += step[2]; tdist[2] += delta[2];
                        if (cur[2] == over[2]) ok = 0;
                    }
                    if (!ok) break;
                    if (t_trav >= this_tmax) break;
                }
            }
        }
        n_edges[i] = ne;
        n_samples[i] = ns;
    }
}
