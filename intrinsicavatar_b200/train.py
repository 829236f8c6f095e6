"""Training-mode seam of the render path: the reference's differentiable graph for the radiance-field
branch of ``forward_`` -- posed sample -> canonical root -> SDF / feature -> radiance + material networks -> Laplace-density
volume rendering -- as three autograd nodes over the CUDA ops, and ``render_radiance`` which chains them.

* ``fused_query``: ``SNARFDeformer.deform(pts, model, eval_mode=False)`` (models/deformers/snarf_deformer.py:170-261) as ONE
  differentiable call.  The reference builds the graph out of ``ForwardDeformer.forward`` (search under ``no_grad`` + the
  implicit-differentiation correction, models/deformers/fast_snarf/deformer_torch.py:34-76), ``VolumeSDF`` (tiny-cuda-nn hash
  grid + MLP, models/rf/geometry.py:147-172) and ``torch.min`` over the candidate roots, and lets autograd walk it.  Here the
  forward is ``ia_op_query_train`` and the backward ``ia_op_query_backward`` (+ ``ia_op_deform_backward`` for a gradient that
  arrives on the canonical point itself), csrc/ia_train.cuh.
* ``shade_fields``: ``VolumeRadiance`` + the material network at the canonical point (models/rf/radiance.py:82-135,
  models/pbr/material.py:31-51): ``ia_op_shade_fields`` / ``ia_op_shade_fields_backward``.
* ``volrend``: ``get_alpha`` + ``render_weight_from_alpha`` + ``accumulate_along_rays`` (models/rf/density.py:17-34,
  models/volrend.py:336-364): ``ia_op_volrend`` / ``ia_op_volrend_backward``.

This module only routes the ops' buffers into autograd.  There is no fallback: without the CUDA library the calls raise.

    sdf, feature, x_c, valid = fused_query(engine, xd, geo_hash, geo_w1, geo_b1, geo_w2, geo_b2, tfs)

The parameter tensors are the ones the engine's fields were set from (``RenderEngine.set_fields``: the hash tables and the
effective -- weight-norm folded -- MLP weights) and ``tfs`` [24,4,4] the bone transforms of ``set_pose``: they are graph
leaves here, the VALUES are the engine's, so call ``set_fields`` / ``set_pose`` after every optimiser step.

Not differentiated: the SDF normal is a constant of the graph (the reference differentiates through it --
second order -- for the eikonal loss and the radiance network's normal input), and so are the sample positions along the ray.
"""
from __future__ import annotations

import torch


class _FusedQuery(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, xd, geo_hash, geo_w1, geo_b1, geo_w2, geo_b2, tfs):
        fwd = engine.op_query_train(xd.detach())
        ctx.engine = engine
        ctx.fwd = {k: fwd[k] for k in ("x_c", "valid", "J_inv")}
        ctx.meta = (geo_hash.shape, geo_hash.dtype, tfs.shape)
        ctx.mark_non_differentiable(fwd["valid"], fwd["grad"])
        return fwd["sdf"], fwd["feature"], fwd["x_c"], fwd["valid"], fwd["grad"]

    @staticmethod
    def backward(ctx, g_sdf, g_feature, g_xc, _g_valid, _g_normal):
        x_c, valid, J_inv = ctx.fwd["x_c"], ctx.fwd["valid"], ctx.fwd["J_inv"]
        n, dev = x_c.shape[0], x_c.device
        # the 13 network outputs at the arg-min root: `feature` IS the output vector (channel 0 = sdf, geometry.py:160-166)
        d_out = torch.zeros(n, 13, device=dev) if g_feature is None else g_feature.to(dev, torch.float32).clone()
        if g_sdf is not None:
            d_out[:, 0] += g_sdf.to(dev, torch.float32)
        g = ctx.engine.op_query_backward(ctx.fwd, d_out)
        g_tfs3 = g["tfs"]
        if g_xc is not None and n:
            # a gradient on the canonical point itself reaches the bone transforms through the same implicit-differentiation
            # correction; ia_op_deform_backward takes roots in groups of 13, so pad the list with invalid ones
            pad = (-n) % 13
            def grp(t, *shape):
                return torch.cat([t, t.new_zeros((pad,) + t.shape[1:])]).reshape(-1, 13, *shape)
            g_tfs3 = g_tfs3 + ctx.engine.op_deform_backward(grp(x_c, 3), grp(valid.to(torch.uint8)), grp(J_inv, 3, 3),
                                                            grp(g_xc.to(dev, torch.float32), 3))
        hash_shape, hash_dtype, tfs_shape = ctx.meta
        g_tfs = torch.zeros(tfs_shape, device=dev)
        g_tfs[..., :3, :] = g_tfs3.reshape(g_tfs[..., :3, :].shape)
        return (None, None, g["hash"].reshape(hash_shape).to(hash_dtype), g["w1"], g["b1"], g["w2"], g["b2"], g_tfs)


def fused_query(engine, xd, geo_hash, geo_w1, geo_b1, geo_w2, geo_b2, tfs, with_normal=False):
    """Differentiable posed point -> (sdf [n], feature [n,13], x_c [n,3], valid [n]) and, ``with_normal``, the posed-space
    SDF gradient [n,3] (a constant of the graph); see the module docstring.  Invalid points: sdf 1e5, no gradient."""
    out = _FusedQuery.apply(engine, xd, geo_hash, geo_w1, geo_b1, geo_w2, geo_b2, tfs)
    return out if with_normal else out[:4]


GEO_PARAMS = ("geo_hash", "geo_w1", "geo_b1", "geo_w2", "geo_b2")


def folded_leaves(named_parameters: dict, device, material_feature: str = "hybrid") -> dict:
    """The effective weights the ops differentiate (``geo_*``, ``rad_*``, ``mat_*``, ``beta``) as tensors on ``device`` that
    stay attached to the reference-keyed parameters they are folded from (``dict(model.named_parameters())`` or a
    checkpoint's ``state_dict`` with ``requires_grad``): weight normalisation of the geometry network
    (models/network_utils.py:201-244), the Lipschitz bound of the material network (:360-428) and
    ``beta = |b| + 1e-4`` (models/rf/density.py:32-34) are ordinary tensor code in ``weights.fold``, so a backward through
    ``render_radiance`` / ``pbr_light`` arrives at the reference's own parameter tree -- what an optimiser over
    ``model.parameters()`` steps.  After the step, upload the new values (``engine.set_fields(weights.fold(...))``)."""
    from .weights import fold
    sd = {k[len("model."):] if k.startswith("model.") else k: v for k, v in named_parameters.items()}
    out = fold(sd, material_feature, keep_graph=True)
    return {k: v.to(device) for k, v in out.items()}


SHADE_PARAMS = ("rad_hash",) + tuple(f"{net}_{t}{i}" for net in ("rad", "mat") for i in (1, 2, 3) for t in ("w", "b"))


class _ShadeFields(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, xc, feature, view_world, normal_world, *params):
        assert len(params) == len(SHADE_PARAMS)
        a = [t.detach() for t in (xc, feature, view_world, normal_world)]
        rgb, mat = engine.op_shade_fields(*a)
        ctx.engine, ctx.a = engine, a
        ctx.meta = (params[0].shape, params[0].dtype)
        return rgb, mat

    @staticmethod
    def backward(ctx, g_rgb, g_mat):
        n, dev = ctx.a[0].shape[0], ctx.engine.dev
        g_rgb = torch.zeros(n, 3, device=dev) if g_rgb is None else g_rgb
        g_mat = torch.zeros(n, 5, device=dev) if g_mat is None else g_mat
        g = ctx.engine.op_shade_fields_backward(*ctx.a, g_rgb, g_mat)
        hash_shape, hash_dtype = ctx.meta
        weights = [g[net][f"{t}{i}"] for net in ("rad", "mat") for i in (1, 2, 3) for t in ("w", "b")]
        return (None, g["x"], g["feature"], None, g["normal"], g["hash"].reshape(hash_shape).to(hash_dtype), *weights)


def shade_fields(engine, xc, feature, view_world, normal_world, params: dict):
    """Differentiable (rgb [n,3], material [n,5]) of the radiance and material networks at canonical points ``xc`` with the
    geometry ``feature``, the ray direction and the world normal.  ``params``: the folded weights by name (``SHADE_PARAMS``).
    Gradients reach ``xc``, ``feature``, ``normal_world`` and every entry of ``params``."""
    return _ShadeFields.apply(engine, xc, feature, view_world, normal_world, *[params[k] for k in SHADE_PARAMS])


class _VolRend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, packed_info, sdf, dists, values, beta):
        b = float(beta.detach()) if torch.is_tensor(beta) else float(beta)
        w, comp, op = engine.op_volrend(packed_info, sdf.detach(), dists, values.detach(), b)
        ctx.engine, ctx.a = engine, (packed_info, sdf.detach(), dists, values.detach(), b)
        ctx.beta_like = beta if torch.is_tensor(beta) else None
        return w, comp, op

    @staticmethod
    def backward(ctx, g_w, g_comp, g_op):
        pi, sdf, dists, values, b = ctx.a
        if g_comp is None:
            g_comp = torch.zeros(pi.shape[0], values.shape[1], device=ctx.engine.dev)
        g_sdf, g_val, g_beta = ctx.engine.op_volrend_backward(pi, sdf, dists, values, b, g_comp, g_op, g_w)
        bl = ctx.beta_like
        return None, None, g_sdf, None, g_val, (g_beta.reshape(bl.shape).to(bl.dtype).to(bl.device) if bl is not None else None)


def volrend(engine, packed_info, sdf, dists, values, beta):
    """Differentiable (weights [m], comp [n_rays,C], opacity [n_rays]) along packed rays (``packed_info``
    [n_rays,2] = first sample, count).  Gradients reach ``sdf``, ``values`` and ``beta`` (a tensor, e.g. the Laplace density's
    parameter, or a float)."""
    return _VolRend.apply(engine, packed_info, sdf, dists, values, beta)


def render_radiance(engine, params: dict, tfs, w2s, rays_o, rays_d, packed_info, t_starts, t_ends, beta):
    """The radiance-field branch of the training forward (models/intrinsic_avatar.py:1068-1182: ``rgb_normal_mats_alpha_fn`` at
    the interval midpoints, then ``rendering_with_normals_mats_sdf``) for samples already placed on the rays: dict with
    ``comp_rgb`` [n_rays,3], ``comp_mats`` [n_rays,5], ``comp_normal`` [n_rays,3] (unnormalised accumulation), ``depth``,
    ``opacity``, ``weights`` and the per-sample ``sdf`` / ``valid`` / ``materials`` [m,5] / ``normal_smpl`` [m,3] (unit SDF
    gradient in SMPL space, no graph) / ``ray_indices`` -- what ``rendering_with_normals_mats_sdf`` hands on in ``extras``.  ``params``: folded weights by name (geometry + shading);
    ``tfs`` the bone transforms and ``w2s`` [4,4] the world-to-SMPL transform of ``set_pose`` (rays are SMPL-space; view direction
    and normal go to the networks in world space, ``transform_dirs_s2w``); ``beta`` the Laplace density's scale.  A loss on the result back-propagates to all of them."""
    dev = engine.dev
    pi = packed_info.to(dev, torch.int32)
    counts = pi[:, 1].long()
    ridx = torch.repeat_interleave(torch.arange(pi.shape[0], device=dev), counts)
    t0, t1 = t_starts.to(dev, torch.float32).reshape(-1), t_ends.to(dev, torch.float32).reshape(-1)
    o, d = rays_o.to(dev, torch.float32)[ridx], rays_d.to(dev, torch.float32)[ridx]
    t_mid = 0.5 * (t0 + t1)
    xd = o + d * t_mid[:, None]
    sdf, feature, x_c, valid, normal = fused_query(engine, xd, *[params[k] for k in GEO_PARAMS], tfs, with_normal=True)
    rot = torch.as_tensor(w2s, dtype=torch.float32, device=dev)[:3, :3]
    view_w = torch.nn.functional.normalize(d @ rot, dim=-1, eps=1e-6)
    normal_w = torch.nn.functional.normalize(normal @ rot, dim=-1, eps=1e-6)
    rgb, mat = shade_fields(engine, x_c, feature, view_w, normal_w, params)
    values = torch.cat([rgb, mat, normal_w, t_mid[:, None]], dim=-1)
    weights, comp, opacity = volrend(engine, pi, sdf, t1 - t0, values, beta)
    return {"comp_rgb": comp[:, 0:3], "comp_mats": comp[:, 3:8], "comp_normal": comp[:, 8:11], "depth": comp[:, 11],
            "opacity": opacity, "weights": weights, "sdf": sdf, "valid": valid, "materials": mat,
            "normal_smpl": torch.nn.functional.normalize(normal, dim=-1, eps=1e-6), "ray_indices": ridx}


class _PbrShade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, wi, n, wo, rough, albedo, metal, Li, inv_pdf):
        a = [t.detach() for t in (wi, n, wo, rough, albedo, metal, Li, inv_pdf)]
        ctx.engine, ctx.a = engine, a
        ctx.shapes = (rough.shape, metal.shape)
        return engine.op_pbr_shade(*a)

    @staticmethod
    def backward(ctx, g_Lo, g_Ld, g_Ls):
        m, dev = ctx.a[0].shape[0], ctx.engine.dev
        g = ctx.engine.op_pbr_shade_backward(*ctx.a, torch.zeros(m, 3, device=dev) if g_Lo is None else g_Lo, g_Ld, g_Ls)
        return (None, None, g["normal"], None, g["rough"].reshape(ctx.shapes[0]), g["albedo"], g["metal"].reshape(ctx.shapes[1]),
                g["Li"], None)


def pbr_shade(engine, wi, n, wo, rough, albedo, metal, Li, inv_pdf):
    """Differentiable (Lo, Lo_diff, Lo_spec) [n,3] of one light direction ``wo`` per shading sample: ``MultiLobe.eval`` under the
    cosine mask times the incoming radiance ``Li`` and the inverse pdf (models/intrinsic_avatar.py:708-751).  Gradients reach
    the normal (pass the output of a normalisation: the component along ``n`` is left to its backward), ``rough``, ``albedo``,
    ``metal`` and ``Li``."""
    return _PbrShade.apply(engine, wi, n, wo, rough, albedo, metal, Li, inv_pdf)


class _EnvEval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, dirs_world, env):
        ctx.engine, ctx.dirs, ctx.meta = engine, dirs_world.detach(), (env.shape, env.dtype, env.device)
        return engine.op_env(dirs_world=ctx.dirs)[2]

    @staticmethod
    def backward(ctx, g_em):
        shape, dtype, device = ctx.meta
        g = ctx.engine.op_env_backward(ctx.dirs, g_em, shape[:2])
        return None, None, g.reshape(shape).to(dtype).to(device)


def env_eval(engine, dirs_world, env):
    """Differentiable ``EnvironmentLightTensor.eval`` (lib/torch_pbr/light.py:298-339): emission [n,3] of world directions on the
    map the engine's light was set from; ``env`` [H,W,3] is that map as a graph leaf (or the softplus of the trainable one)."""
    return _EnvEval.apply(engine, dirs_world, env)


def pbr_light(engine, env, w2s, normal, albedo, rough, metal, positions, view_dirs, light_dirs, inv_pdf, gi=True):
    """One light direction per shading sample, the way the training-time integrators combine it (``pbr_uniform_light_forward``,
    models/intrinsic_avatar.py:654-760, the training default; its siblings differ in where ``light_dirs`` / ``inv_pdf`` come
    from): secondary rays traced without a graph (transmittance and indirect radiance are constants, :673-706), emission
    looked up in world space, BRDF and combine differentiable.  ``normal`` unit, SMPL space, like ``view_dirs`` (camera ray
    direction), ``light_dirs`` and ``positions``.  Returns (Lo, Lo_diff, Lo_spec, vis) per sample."""
    dev = engine.dev
    n_, wo = normal.to(dev, torch.float32), light_dirs.to(dev, torch.float32)
    with torch.no_grad():
        cos_mask = (n_ * wo).sum(-1) > 1e-6
        tr = torch.zeros(wo.shape[0], device=dev)
        rgb = torch.zeros(wo.shape[0], 3, device=dev)
        if bool(cos_mask.any()):
            t_, r_ = engine.op_secondary(positions.to(dev, torch.float32)[cos_mask], wo[cos_mask], gi=gi)
            tr[cos_mask], rgb[cos_mask] = t_.clamp(0.0, 1.0), r_
        rot = torch.as_tensor(w2s, dtype=torch.float32, device=dev)[:3, :3]
        dirs_world = torch.nn.functional.normalize(wo @ rot, dim=-1, eps=1e-6)
    em = env_eval(engine, dirs_world, env)
    Li = em * tr[:, None] + rgb if gi else em * tr[:, None]
    Lo, Lo_diff, Lo_spec = pbr_shade(engine, -view_dirs.to(dev, torch.float32), n_, wo, rough, albedo, metal, Li,
                                     inv_pdf.to(dev, torch.float32))
    return Lo, Lo_diff, Lo_spec, 2.0 * tr[:, None].expand(-1, 3)


def render_phys(engine, params: dict, tfs, w2s, env, rays_o, rays_d, packed_info, t_starts, t_ends, beta, light_dirs, inv_pdf,
                light_index, spp=512, background=1.0, gi=False):
    """Both branches of the training forward for samples already placed on the rays (models/intrinsic_avatar.py:1241-1470 with
    ``render_mode = uniform_light``, the training default): ``render_radiance``, then -- without a graph -- ``spp`` shading
    samples per ray drawn from the compositing weights (``sample_volume_interaction``, models/pbr/utils.py:70-229:
    ``ia_op_ray_resampling``, zero-crossing snap included), each with its source interval's normal and materials and the
    weight ``w_source / count`` (foreground) or ``(1 - opacity) / count`` (background), both of which keep their graph;
    one light direction per foreground sample -- ``light_dirs`` / ``inv_pdf`` [L,3] / [L] indexed by ``light_index``
    [n_rays, spp] (the reference's per-ray shuffle of the stratified sphere, :1391-1411) -- through ``pbr_light``;
    background samples carry ``background``; ``comp_rgb_phys = sum w Lo`` per ray, rays without samples = background.
    Returns ``render_radiance``'s dict plus ``comp_rgb_phys`` [n_rays,3], ``visibility`` [n_rays,1] and ``n_shading_samples``."""
    dev = engine.dev
    out = render_radiance(engine, params, tfs, w2s, rays_o, rays_d, packed_info, t_starts, t_ends, beta)
    pi = packed_info.to(dev, torch.int32)
    n_rays = pi.shape[0]
    weights, mats, opacity = out["weights"], out["materials"], out["opacity"]
    bgc = torch.as_tensor(background, dtype=torch.float32, device=dev).expand(3)
    t0, t1 = t_starts.to(dev, torch.float32).reshape(-1), t_ends.to(dev, torch.float32).reshape(-1)
    with torch.no_grad():
        rpi, t_res, offs, src, fg_cnt, bg_cnt, _ = engine.op_ray_resampling(pi, t0, t1, weights.detach(), out["sdf"].detach(), spp)
        rcount = rpi[:, 1].long()
        r_ridx = torch.repeat_interleave(torch.arange(n_rays, device=dev), rcount)
        rank = torch.arange(r_ridx.shape[0], device=dev) - rpi[:, 0].long()[r_ridx]          # position within its ray
        is_fg = offs.reshape(-1) < 1e4
        fg, bg = torch.nonzero(is_fg).reshape(-1), torch.nonzero(~is_fg).reshape(-1)
        fg_src, fg_ray, bg_ray = src[fg], r_ridx[fg], r_ridx[bg]
        o, d = rays_o.to(dev, torch.float32), rays_d.to(dev, torch.float32)
        positions = o[fg_ray] + d[fg_ray] * t_res.reshape(-1)[fg][:, None]
        li = light_index.to(dev).long()[fg_ray, rank[fg]]
        wo, ip = light_dirs.to(dev, torch.float32)[li], inv_pdf.to(dev, torch.float32)[li]
    w_fg = weights[fg_src] / fg_cnt[fg_src].float()
    w_bg = (1.0 - opacity)[bg_ray] / bg_cnt[bg_ray].float()
    m = mats[fg_src]
    Lo, _, _, vis = pbr_light(engine, env, w2s, out["normal_smpl"][fg_src], m[:, 0:3], m[:, 3], m[:, 4], positions, d[fg_ray], wo,
                              ip, gi=gi)
    phys = torch.zeros(n_rays, 3, device=dev).index_add(0, fg_ray, w_fg[:, None] * Lo)
    phys = phys.index_add(0, bg_ray, w_bg[:, None] * bgc[None, :])
    empty = rcount == 0
    phys = torch.where(empty[:, None], bgc[None, :].expand(n_rays, 3), phys)
    visibility = torch.zeros(n_rays, 3, device=dev).index_add(0, fg_ray, w_fg.detach()[:, None] * vis).mean(-1, keepdim=True)
    out.update(comp_rgb_phys=phys, visibility=visibility, n_shading_samples=int(fg.shape[0]))
    return out
