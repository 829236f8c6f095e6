"""ORACLE (test infrastructure): the per-frame render path end to end on CPU (PyTorch fp32).

Follows the reference's eval-mode control flow with its own intermediate tensors (no lazy
evaluation, no fusion) so it is an independent check of the fused kernels:

  prepare            models/intrinsic_avatar.py:281-381 (+ models/utils.py:152-163)
  forward_           models/intrinsic_avatar.py:950-1651  (eval, enable_phys, importance_sample,
                     render_mode = light | uniform_light | mats | mis; albedo_only, add_emitter supported)
  compute_indirect_radiance   models/intrinsic_avatar.py:396-545
  pbr_light_forward / pbr_uniform_light_forward / pbr_mats_forward / pbr_mis_forward
                     models/intrinsic_avatar.py:755-861, 654-753, 863-948, 547-652
  rendering_with_normals_mats_sdf / rendering   models/volrend.py:810-1020, 19-194
  sample_volume_interaction   models/pbr/utils.py:70-229
  transform_rays_w2s / dirs   models/deformers/snarf_deformer.py:128-158

Randomness is an explicit input: occupancy jitter ``[res^3,3,3]``, light uniforms ``u1,u2 [spp]``
and the ``seed`` of the per-ray permutation / of the counter-based uniforms that replace ``torch.rand``
in MultiLobe.sample and emitter.sample (oracle/pbr.py) -- SURVEY.md section 7 "RNG".
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .deformer import deform, precompute
from .pbr import (EnvLight, kensler_permute, multilobe_eval, multilobe_pdf, multilobe_sample, pixel_key, rgb_to_srgb,
                  rng_uniform, uniform_sphere_stratified)

SCENE_AABB = (-1.25, -1.55, -1.25, 1.25, 0.95, 1.25)  # configs/dataset/animation/male-3-casual.yaml:10


def _normalize(v, eps=1e-6):
    return F.normalize(v, p=2, dim=-1, eps=eps)


def occupancy_ema_step(occs, occ, ema_decay, occ_thre, R):
    """The grid logic of OccGridEstimator._update (models/occ_grid/temporal_occ_grid.py:392-411): EMA with max, 3^3 max-pool,
    threshold min(mean, occ_thre), largest connected component (models/utils.py:152-163).  Pinned to the reference's own
    class by tests/golden/reference_vectors_occ_ema.npz.  Returns (new occs [R^3], binaries [R,R,R])."""
    occs = torch.maximum(torch.as_tensor(occs, dtype=torch.float32).reshape(-1) * ema_decay,
                         torch.as_tensor(occ, dtype=torch.float32).reshape(-1))
    occs_ = F.max_pool3d(occs.reshape(1, 1, R, R, R), kernel_size=3, stride=1, padding=1)[0, 0].reshape(-1)
    thre = torch.clamp(occs_[occs_ >= 0].mean(), max=occ_thre)
    binaries = (occs_ > thre).reshape(1, R, R, R)
    comp = torch.arange(1, R ** 3 + 1).reshape(1, 1, R, R, R).float()
    grid = binaries[None]
    comp[~grid] = 0
    for _ in range(R * 3):
        comp = F.max_pool3d(comp, kernel_size=3, stride=1, padding=1)
        comp *= grid
    mcc = comp[0, 0]
    label = torch.mode(mcc[binaries[0]], 0).values
    return occs, (mcc == label).reshape(R, R, R)


class OracleRenderer:
    def __init__(self, fields, lbs_voxel, offset_kernel, scale_kernel, *, samples_per_pixel=4,
                 global_illumination=False, num_samples_per_ray=128, num_samples_per_secondary_ray=64,
                 secondary_near=0.0, secondary_far=1.5, occ_thre=0.001, grid_res=64,
                 query_chunk=65536, render_mode="light", add_emitter=False, secondary_importance_sample=True,
                 zero_crossing_search=True):
        assert render_mode in ("light", "uniform_light", "mats", "mis")
        self.render_mode = render_mode
        self.add_emitter = add_emitter
        # models/intrinsic_avatar.py:482-520
        self.secondary_importance_sample, self.zero_crossing_search = secondary_importance_sample, zero_crossing_search
        self.fields = fields
        self.lbs_voxel = torch.as_tensor(lbs_voxel, dtype=torch.float32)
        self.offset = torch.as_tensor(offset_kernel, dtype=torch.float32)
        self.scale = torch.as_tensor(scale_kernel, dtype=torch.float32)
        self.spp = samples_per_pixel
        self.gi = global_illumination
        aabb = torch.tensor(SCENE_AABB)
        self.render_step_size = float(torch.norm(aabb[3:] - aabb[:3])) / num_samples_per_ray
        self.sec_near, self.sec_far = secondary_near, secondary_far
        self.sec_step = (secondary_far - secondary_near) / (num_samples_per_secondary_ray - 1)
        self.occ_thre = occ_thre
        self.grid_res = grid_res
        self.chunk = query_chunk
        self.background = torch.ones(3)
        self.albedo_align_ratio = None     # set externally at test time (systems/intrinsic_avatar.py:601-617)
        self.counters = {"n_query": 0, "n_query_grad": 0, "n_radiance": 0, "n_secondary_rays": 0}

    # ------------------------------------------------------------------ per-frame state ----
    def set_pose(self, tfs, w2s):
        self.tfs = torch.as_tensor(tfs, dtype=torch.float32)
        self.w2s = torch.as_tensor(w2s, dtype=torch.float32)
        self.voxel_J = precompute(self.lbs_voxel, self.tfs)

    def _deform(self, x, with_grad=False):
        outs = []
        for i in range(0, x.shape[0], self.chunk):
            outs.append(deform(x[i:i + self.chunk], self.fields, self.voxel_J, self.lbs_voxel, self.tfs,
                               self.offset, self.scale, with_grad))
        self.counters["n_query_grad" if with_grad else "n_query"] += x.shape[0]
        if not outs:
            outs = [deform(x, self.fields, self.voxel_J, self.lbs_voxel, self.tfs, self.offset, self.scale,
                           with_grad)]
        return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}

    def build_occupancy(self, aabb, jitter):
        """_compute_occupancy_grid + prepare_test_occupancy_grid (intrinsic_avatar.py:307-381)."""
        R = self.grid_res
        aabb = torch.as_tensor(aabb, dtype=torch.float32).reshape(6)
        ar = torch.arange(R)
        coords = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), -1).reshape(-1, 1, 3).float()
        x = (coords + torch.as_tensor(jitter, dtype=torch.float32).reshape(-1, 3, 3)) / R
        x = x.reshape(-1, 3) * (aabb[3:] - aabb[:3]) + aabb[:3]
        sdf = self._deform(x)["sdf"]
        alpha = self.fields.alpha_from_sdf(sdf, self.render_step_size)
        occs = alpha.reshape(-1, 3).max(1)[0]
        occs_ = F.max_pool3d(occs.reshape(1, 1, R, R, R), kernel_size=3, stride=1, padding=1)[0, 0].reshape(-1)
        thre = torch.clamp(occs_[occs_ >= 0].mean(), max=self.occ_thre)
        binaries = (occs_ > thre).reshape(1, R, R, R)
        comp = torch.arange(1, R ** 3 + 1).reshape(1, 1, R, R, R).float()
        grid = binaries[None]
        comp[~grid] = 0
        for _ in range(R * 3):
            comp = F.max_pool3d(comp, kernel_size=3, stride=1, padding=1)
            comp *= grid
        mcc = comp[0, 0]
        label = torch.mode(mcc[binaries[0]], 0).values
        self.binaries = (mcc == label).reshape(R, R, R)
        self.grid_aabb = aabb
        self.occs = occs
        return self.binaries

    def update_occupancy_ema(self, aabb, jitter, occs, ema_decay=0.8, occ_thre=0.001):
        """OccGridEstimator._update for one level (models/occ_grid/temporal_occ_grid.py:369-411) with the occ_eval_fn of
        IntrinsicAvatarModel.update_step (models/intrinsic_avatar.py:243-254): one jittered point per cell, EMA with max,
        max-pool, threshold, largest component.  Returns (new occs [R^3], binaries [R,R,R])."""
        R = self.grid_res
        aabb = torch.as_tensor(aabb, dtype=torch.float32).reshape(6)
        ar = torch.arange(R)
        coords = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), -1).reshape(-1, 3).float()
        x = (coords + torch.as_tensor(jitter, dtype=torch.float32).reshape(-1, 3)) / R
        x = x * (aabb[3:] - aabb[:3]) + aabb[:3]
        occ = self.fields.alpha_from_sdf(self._deform(x)["sdf"], self.render_step_size).reshape(-1)
        return occupancy_ema_step(occs, occ, ema_decay, occ_thre, R)

    def set_light(self, envmap, u1, u2):
        self.env = EnvLight(torch.as_tensor(envmap, dtype=torch.float32))
        self.light_dirs_world = self.env.sample(torch.as_tensor(u1, dtype=torch.float32),
                                                torch.as_tensor(u2, dtype=torch.float32))
        assert self.light_dirs_world.shape[0] == self.spp

    def set_light_uniform(self, envmap, n_rows=16, n_cols=32):
        """render_mode = uniform_light: stratified sphere directions, used in the SMPL frame as they are."""
        self.env = EnvLight(torch.as_tensor(envmap, dtype=torch.float32))
        self.uniform_dirs = uniform_sphere_stratified(n_rows, n_cols)
        assert self.uniform_dirs.shape[0] == self.spp

    # ------------------------------------------------------------------------ transforms ----
    def dirs_w2s(self, d):
        return _normalize(d @ self.w2s[:3, :3].t())

    def dirs_s2w(self, d):
        return _normalize(d @ self.w2s[:3, :3])

    # ------------------------------------------------------------------- secondary rays ----
    def compute_indirect_radiance(self, rays_o, rays_d):
        n_rays = rays_o.shape[0]
        self.counters["n_secondary_rays"] += n_rays
        tg = ops.traverse_grid(rays_o, rays_d, self.binaries, self.grid_aabb, self.sec_near, self.sec_far,
                               self.sec_step)
        t_starts, t_ends, ray_indices = tg["t_starts"], tg["t_ends"], tg["sample_ray_indices"]
        if t_starts.numel() > 0 and self.secondary_importance_sample:
            pos = rays_o[ray_indices] + rays_d[ray_indices] * t_starts[:, None]
            sdfs = self._deform(pos)["sdf"]
            alphas = self.fields.alpha_from_sdf(sdfs, t_ends - t_starts)
            packed = ops.pack_info(ray_indices, n_rays)
            if self.zero_crossing_search:
                rpi, rs, re, is_fg = ops.ray_resampling_sdf_fine(packed, t_starts[:, None], t_ends[:, None],
                                                                 alphas, sdfs, 4)
            else:
                weights, _ = ops.render_weight_from_alpha(alphas, packed)
                rpi, rs, re, is_fg = ops.ray_resampling_fine(packed, t_starts[:, None], t_ends[:, None], weights, 4)
            rri = ops.unpack_info(rpi, len(rs))
            ray_indices = rri[is_fg]
            t_starts = rs[is_fg, 0]
            t_ends = re[is_fg, 0]
        rgb = torch.zeros(n_rays, 3)
        acc = torch.zeros(n_rays, 1)
        if t_starts.numel() > 0:
            o, d = rays_o[ray_indices], rays_d[ray_indices]
            pos = o + d * (t_starts + t_ends)[:, None] / 2.0
            q = self._deform(pos, with_grad=True)
            alphas = self.fields.alpha_from_sdf(q["sdf"], t_ends - t_starts)
            view_w = self.dirs_s2w(d)
            normal_w = self.dirs_s2w(q["grad"])
            rgbs, _ = self.fields.radiance(q["x_c"], q["feature"], view_w, normal_w)
            self.counters["n_radiance"] += pos.shape[0]
            packed = ops.pack_info(ray_indices, n_rays)
            weights, _ = ops.render_weight_from_alpha(alphas, packed)
            rgb = ops.accumulate_along_rays(weights, rgbs, ray_indices, n_rays)
            acc = ops.accumulate_along_rays(weights, None, ray_indices, n_rays)
        return 1.0 - acc, rgb

    # ------------------------------------------------------------------------- shading ----
    def pbr_light_forward(self, normal, albedo, roughness, metallic, positions, dirs, light_idx):
        wi = -dirs
        sec_d = self.dirs_w2s(self.light_dirs_world)[light_idx]
        sec_o = positions
        cos_mask = (normal * sec_d).sum(-1) > 1e-6
        tr = torch.zeros(len(cos_mask), 1)
        sec_rgb = torch.zeros(len(cos_mask), 3)
        if cos_mask.sum() > 0:
            t_, r_ = self.compute_indirect_radiance(sec_o[cos_mask], sec_d[cos_mask])
            tr[cos_mask] = t_
            sec_rgb[cos_mask] = r_
            tr.clamp_(0.0, 1.0)
        tr_mask = tr[:, 0] > 0.0
        diff = torch.zeros_like(albedo[:, :1])
        spec = torch.zeros_like(albedo)
        if cos_mask.sum() > 0:
            diff[cos_mask], spec[cos_mask] = multilobe_eval(
                wi[cos_mask], normal[cos_mask], sec_d[cos_mask], roughness[cos_mask, 0], albedo[cos_mask],
                metallic[cos_mask])
        em = torch.zeros_like(sec_rgb)
        m = cos_mask & tr_mask
        if m.sum() > 0:
            em[m] = self.env.eval(self.dirs_s2w(sec_d[m]))
        Li = em * tr + sec_rgb if self.gi else em * tr
        pdf = torch.ones_like(em[:, :1])
        if m.sum() > 0:
            pdf[m] = self.env.pdf(self.dirs_s2w(sec_d[m]))
        pdf = torch.where(pdf > 0, pdf, torch.ones_like(pdf))
        Lo_diff = Li * diff / pdf
        Lo_spec = Li * spec / pdf
        kd = (1.0 - metallic) * albedo
        Lo = kd * Lo_diff + Lo_spec
        return Lo, Lo_diff, Lo_spec

    def _compose(self, Li_diff, Li_spec, albedo, metallic):
        kd = (1.0 - metallic) * albedo
        return kd * Li_diff + Li_spec

    def pbr_uniform_light_forward(self, normal, albedo, roughness, metallic, positions, dirs, light_idx):
        """models/intrinsic_avatar.py:654-753 (eval: fixed 16x32 directions, inv_pdf = 4 pi)."""
        wi = -dirs
        sec_d = self.uniform_dirs[light_idx]
        inv_pdf = 4 * np.pi
        cos_mask = (normal * sec_d).sum(-1) > 1e-6
        tr = torch.zeros(len(cos_mask), 1)
        sec_rgb = torch.zeros(len(cos_mask), 3)
        if cos_mask.sum() > 0:
            t_, r_ = self.compute_indirect_radiance(positions[cos_mask], sec_d[cos_mask])
            tr[cos_mask] = t_
            sec_rgb[cos_mask] = r_
            tr.clamp_(0.0, 1.0)
        tr_mask = tr[:, 0] > 0.0
        diff = torch.zeros_like(albedo[:, :1])
        spec = torch.zeros_like(albedo)
        if cos_mask.sum() > 0:
            diff[cos_mask], spec[cos_mask] = multilobe_eval(
                wi[cos_mask], normal[cos_mask], sec_d[cos_mask], roughness[cos_mask, 0], albedo[cos_mask],
                metallic[cos_mask])
        em = torch.zeros_like(sec_rgb)
        m = cos_mask & tr_mask
        if m.sum() > 0:
            em[m] = self.env.eval(self.dirs_s2w(sec_d[m]))
        Li = em * tr + sec_rgb if self.gi else em * tr
        Lo_diff = Li * diff * inv_pdf
        Lo_spec = Li * spec * inv_pdf
        vis = 2 * torch.ones_like(em) * tr
        return self._compose(Lo_diff, Lo_spec, albedo, metallic), Lo_diff, Lo_spec, vis

    def _scatter_dirs(self, normal, albedo, roughness, metallic, wi, key, j):
        u = torch.stack([rng_uniform(key, j, 0), rng_uniform(key, j, 1)], -1)
        return multilobe_sample(normal, wi, roughness[:, 0], albedo, metallic, u)

    def pbr_mats_forward(self, normal, albedo, roughness, metallic, positions, dirs, key, j):
        """models/intrinsic_avatar.py:863-948: BSDF sampling, no cosine mask, transmittance not clamped."""
        wi = -dirs
        sec_d = self._scatter_dirs(normal, albedo, roughness, metallic, wi, key, j)
        tr, sec_rgb = self.compute_indirect_radiance(positions, sec_d)
        pdf = multilobe_pdf(wi, normal, sec_d, roughness[:, 0], albedo, metallic)
        pdf = torch.where(pdf > 0, pdf, torch.ones_like(pdf))
        diff, spec = multilobe_eval(wi, normal, sec_d, roughness[:, 0], albedo, metallic)
        em = self.env.eval(self.dirs_s2w(sec_d))
        Li = em * tr + sec_rgb if self.gi else em * tr
        Lo_diff = Li * diff / pdf
        Lo_spec = Li * spec / pdf
        return self._compose(Lo_diff, Lo_spec, albedo, metallic), Lo_diff, Lo_spec

    def pbr_mis_forward(self, normal, albedo, roughness, metallic, positions, dirs, key, j):
        """models/intrinsic_avatar.py:547-652: one BSDF-sampled and one light-sampled ray per shading sample,
        each weighted 1 / (pdf_scatter + pdf_light)."""
        wi = -dirs
        scatter_d = self._scatter_dirs(normal, albedo, roughness, metallic, wi, key, j)
        light_d = self.dirs_w2s(self.env.sample(rng_uniform(key, j, 2), rng_uniform(key, j, 3)))
        sec_d = torch.cat([scatter_d, light_d], 0)
        tr, sec_rgb = self.compute_indirect_radiance(positions.repeat(2, 1), sec_d)
        n2, wi2, r2, a2, m2 = normal.repeat(2, 1), wi.repeat(2, 1), roughness[:, 0].repeat(2), albedo.repeat(2, 1), \
            metallic.repeat(2, 1)
        pdf_s = multilobe_pdf(wi2, n2, sec_d, r2, a2, m2)
        pdf_l = self.env.pdf(self.dirs_s2w(sec_d))
        diff, spec = multilobe_eval(wi2, n2, sec_d, r2, a2, m2)
        em = self.env.eval(self.dirs_s2w(sec_d))
        Li = em * tr + sec_rgb if self.gi else em * tr
        w = torch.where(pdf_s + pdf_l > 1e-6, torch.reciprocal(pdf_s + pdf_l), torch.zeros_like(pdf_s))
        Lo_diff = (Li * diff) * w
        Lo_spec = (Li * spec) * w
        Lo = self._compose(Lo_diff, Lo_spec, a2, m2)
        return (Lo.reshape(2, -1, 3).sum(0), Lo_diff.reshape(2, -1, 3).sum(0), Lo_spec.reshape(2, -1, 3).sum(0))

    # ------------------------------------------------------------------- primary rays -----
    def forward(self, rays, seed=0, ray_chunk=4096, albedo_only=False):
        rays = torch.as_tensor(rays, dtype=torch.float32)
        outs = []
        for i in range(0, rays.shape[0], ray_chunk):
            outs.append(self.forward_(rays[i:i + ray_chunk], i, seed, albedo_only))
        return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}

    def forward_(self, rays, ray_offset, seed, albedo_only=False):
        R, t = self.w2s[:3, :3], self.w2s[:3, 3]
        rays_o = rays[:, :3] @ R.t() + t
        rays_d = rays[:, 3:6] @ R.t()
        far = torch.linalg.norm(rays_o, dim=-1) + 1
        n_rays = rays.shape[0]
        step = self.render_step_size
        F_ = self.fields

        tg = ops.traverse_grid(rays_o, rays_d, self.binaries, self.grid_aabb, 0.0, 1e10, step)
        vals, is_left, is_right = tg["vals"], tg["is_left"], tg["is_right"]
        e_ray, packed = tg["ray_indices"], tg["packed_info"].int()

        if tg["t_starts"].numel() > 0:
            for it in range(2):
                if it == 0:
                    pos = rays_o[e_ray] + rays_d[e_ray] * vals[:, None]
                    sdf = self._deform(pos)["sdf"]
                    sdf_merge = torch.full_like(vals, 1e10)
                    sdf_merge[is_left] = torch.minimum(sdf[is_left], sdf[is_right])
                    alphas = F_.alpha_from_sdf(sdf_merge, torch.full_like(vals, step))
                else:
                    ri = e_ray[is_left]
                    ts, te = vals[is_left], vals[is_right]
                    pos = rays_o[ri] + rays_d[ri] * (ts + te)[:, None] / 2.0
                    sdf_c = self._deform(pos)["sdf"]
                    sdf = torch.full_like(vals, 1e10)
                    sdf[is_left] = sdf_c
                    dists = torch.zeros_like(vals)
                    dists[is_left] = te - ts
                    alphas = F_.alpha_from_sdf(sdf, dists)
                weights, _ = ops.render_weight_from_alpha(alphas, packed)
                rpi, rv, rd, ril, rir, isr, isfg = ops.ray_resampling_merge(packed, vals, is_left, is_right,
                                                                           weights, 16)
                e_ray = ops.unpack_info(rpi, len(rv))[isfg]
                packed = ops.pack_info(e_ray, n_rays)
                vals, is_left, is_right = rv[isfg], ril[isfg], rir[isfg]

        t_starts, t_ends = vals[is_left], vals[is_right]
        ray_indices = e_ray[is_left]
        S = t_starts.shape[0]

        # rendering_with_normals_mats_sdf(rgb_normal_mats_alpha_fn)
        o, d = rays_o[ray_indices], rays_d[ray_indices]
        pos = o + d * (t_starts + t_ends)[:, None] / 2.0
        q = self._deform(pos, with_grad=True) if S > 0 else None
        if S > 0:
            normal_smpl = _normalize(q["grad"])
            normal_world = self.dirs_s2w(q["grad"])
            view_w = self.dirs_s2w(d)
            alphas = F_.alpha_from_sdf(q["sdf"], t_ends - t_starts)
            rgbs, xyz_embd = F_.radiance(q["x_c"], q["feature"], view_w, normal_world)
            self.counters["n_radiance"] += S
            mats = F_.material(xyz_embd, q["feature"])
            if self.albedo_align_ratio is not None:                      # models/intrinsic_avatar.py:1114-1115
                mats = torch.cat([mats[:, :3] * torch.as_tensor(self.albedo_align_ratio, dtype=torch.float32), mats[:, 3:]], -1)
            sdf_s = q["sdf"]
        else:
            normal_smpl = normal_world = rgbs = torch.zeros(0, 3)
            alphas = sdf_s = torch.zeros(0)
            mats = torch.zeros(0, 5)
        spacked = ops.pack_info(ray_indices, n_rays)
        weights, _ = ops.render_weight_from_alpha(alphas, spacked)
        albedo, rough, metal = mats[:, :3], mats[:, 3:4], mats[:, 4:]
        acc = ops.accumulate_along_rays
        rgb_map = acc(weights, rgbs, ray_indices, n_rays)
        normal_map = acc(weights, normal_world, ray_indices, n_rays)
        albedo_map = acc(weights, albedo, ray_indices, n_rays)
        rough_map = acc(weights, rough, ray_indices, n_rays)
        metal_map = acc(weights, metal, ray_indices, n_rays)
        acc_map = acc(weights, None, ray_indices, n_rays)
        depth_map = acc(weights, (t_starts + t_ends)[:, None] / 2.0, ray_indices, n_rays)
        depth_map = depth_map + (1.0 - acc_map) * far[:, None]

        bgc = self.background
        # add_emitter: the envmap seen along the primary ray replaces the background colour
        # (models/intrinsic_avatar.py:1319-1341, 1454-1490)
        bg_ray = self.env.eval(self.dirs_s2w(rays_d)) if (self.add_emitter and not albedo_only) else \
            bgc[None].expand(n_rays, 3)
        rgb_phys = bg_ray.clone()
        demod_phys = rgb_phys.clone()
        visibility = torch.zeros(n_rays, 1)
        if S > 0 and not albedo_only:
            spp = self.spp
            (rpi, rts, roffs, ridx, fg_cnt, bg_cnt, _surf) = ops.ray_resampling(
                spacked, t_starts[:, None], t_ends[:, None], weights, sdf_s, spp)
            fg_i = torch.nonzero(roffs[:, 0] < 1e4)[:, 0]
            bg_i = torch.nonzero(roffs[:, 0] >= 1e4)[:, 0]
            rri = ops.unpack_info(rpi, len(rts))
            if fg_i.numel() > 0:
                fg_ray, bg_ray_i = rri[fg_i], rri[bg_i]
                src = ridx[fg_i]
                rw = torch.zeros(len(rts))
                rw[fg_i] = weights[src] / fg_cnt[src].float()
                rw[bg_i] = (1.0 - acc_map)[bg_ray_i, 0] / bg_cnt[bg_ray_i].float()
                t = rts[fg_i]
                positions = rays_o[fg_ray] + rays_d[fg_ray] * t
                # per-(ray, sample) light index: sample j of a hit ray is its j-th resample
                j = (fg_i - rpi[fg_ray, 0].long()).numpy()
                key = pixel_key(seed, fg_ray.numpy() + ray_offset)
                light_idx = torch.from_numpy(kensler_permute(j.astype(np.uint64), spp, key))
                Lo = torch.zeros(len(rts), 3)
                Lo[bg_i] = bg_ray[bg_ray_i]
                Lo_demod = Lo.clone()
                args = (normal_smpl[src], albedo[src], rough[src], metal[src], positions, rays_d[fg_ray])
                if self.render_mode == "light":
                    fg_Lo, fg_d, fg_s = self.pbr_light_forward(*args, light_idx)
                elif self.render_mode == "uniform_light":
                    fg_Lo, fg_d, fg_s, fg_vis = self.pbr_uniform_light_forward(*args, light_idx)
                    vis = torch.zeros(len(rts), 3)
                    vis[fg_i] = fg_vis
                    visibility = acc(rw, vis, rri, n_rays).mean(-1, keepdim=True)
                else:
                    fwd = self.pbr_mats_forward if self.render_mode == "mats" else self.pbr_mis_forward
                    fg_Lo, fg_d, fg_s = fwd(*args, key, j)
                Lo[fg_i] = fg_Lo
                Lo_demod[fg_i] = fg_d + fg_s
                rgb_phys = acc(rw, Lo, rri, n_rays)
                demod_phys = acc(rw, Lo_demod, rri, n_rays)
                empty = torch.nonzero(rpi[:, 1] <= 0)[:, 0]
                rgb_phys[empty] = bg_ray[empty]
                demod_phys[empty] = bg_ray[empty]

        out = {
            "comp_rgb": rgb_map, "comp_normal": normal_map, "opacity": acc_map, "depth": depth_map,
            "rays_valid": acc_map > 0, "rays_valid_phys": acc_map > 0,
            "comp_rgb_phys": rgb_phys, "comp_demod_phys": demod_phys, "comp_albedo": albedo_map,
            "comp_metallic": metal_map, "comp_roughness": rough_map,
        }
        bgm = bgc.mean()
        full = {
            "comp_rgb_full": rgb_to_srgb(rgb_map + bgc[None] * (1.0 - acc_map)).clamp(0, 1),
            "comp_rgb_phys_full": rgb_to_srgb(rgb_phys).clamp(0, 1),
            "comp_demod_phys_full": rgb_to_srgb(demod_phys).clamp(0, 1),
            "comp_albedo_full": albedo_map + 0.0 * (1.0 - acc_map),
            "comp_metallic_full": metal_map + bgm * (1.0 - acc_map),
            "comp_roughness_full": rough_map + bgm * (1.0 - acc_map),
        }
        out.update(full)
        if self.render_mode == "uniform_light":
            out["visibility"] = visibility
        out["num_samples_per_ray"] = spacked[:, 1:2].clone()
        return out
