#!/usr/bin/env python
"""Run the REFERENCE's own ``IntrinsicAvatarModel.forward_`` on the CPU (build container only: needs /root/reference).

The reference cannot run as shipped here -- no GPU, no nerfacc / tiny-cuda-nn / Lightning / SMPL data.
This harness imports its first-party Python (models/intrinsic_avatar.py, models/volrend.py, models/pbr/utils.py,
models/deformers/*, models/rf/*, models/pbr/material.py, lib/torch_pbr/*, lib/nerfacc/{cdf,pack}.py) from where it lies
and replaces ONLY what is not first-party Python:

  replaced                                               by
  -----------------------------------------------------  ---------------------------------------------------------------
  nerfacc 0.5.3: traverse_grids, render_weight_from_alpha, oracle/ops.py (C restatement, "parity unpinned": third party,
    accumulate_along_rays, OccGridEstimator, RayIntervals   sources absent)
  tinycudann.Encoding (HashGrid, SphericalHarmonics)      oracle/fields.py hashgrid / sh4 ("parity unpinned")
  first-party CUDA extensions: fuse_cuda.fuse_broyden,    oracle/deformer.py, oracle/ops.py -- each pinned to the
    filter.filter, precompute.precompute, lib.nerfacc.cuda   reference's own compiled kernel on the GPU box
    ray_resampling*, unpack_info, unpack_data                (tests/test_gpu_ref_ab.py)
  SMPL .pkl + smplx.SMPL                                  per-frame bone transforms of SnarfSetup (pinned to the
                                                            reference's prepare_deformer: reference_vectors_snarf.npz)
  torch.rand / torch.rand_like (global RNG)               the explicit tables product and oracle share (light uniforms,
                                                            per-ray light permutation)
  device="cuda" / .cuda()                                 CPU

So the CONTROL FLOW and all glue of forward_ / compute_indirect_radiance / pbr_*_forward / rendering_with_normals_mats_sdf
/ sample_volume_interaction / SNARFDeformer.deform / ForwardDeformer.forward are the reference's own code, executed.
``scripts/make_golden.py e2e`` stores its outputs as tests/golden/reference_vectors_e2e.npz; tests/test_oracle_golden.py
holds oracle/render.py to them.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)


class Cfg(dict):
    """Attribute-access dict standing in for OmegaConf's DictConfig."""
    __getattr__ = dict.__getitem__

    def copy(self):
        return Cfg(self)


def cfg(d):
    return Cfg({k: cfg(v) if isinstance(v, dict) else v for k, v in d.items()})


def load_yaml(path, **subst):
    import yaml
    txt = open(os.path.join(REF, "configs", path)).read()
    for k, v in subst.items():
        txt = txt.replace(k, str(v))
    return cfg(yaml.safe_load(txt))


_installed = False


def install():
    """Stub modules + patches.  Idempotent."""
    global _installed
    if _installed:
        return
    _installed = True
    import make_golden as MG
    MG.install_stubs()
    _stub, _pkg = MG._stub, MG._pkg
    from oracle import deformer as odef, fields as OF, ops as oops
    from intrinsicavatar_b200.weights import hashgrid_layout
    layout = hashgrid_layout()

    # ---- tiny-cuda-nn
    class Encoding(torch.nn.Module):
        def __init__(self, n_input_dims, config, dtype=torch.float32):
            super().__init__()
            self.n_input_dims, self.otype = n_input_dims, config["otype"]
            if self.otype == "HashGrid":
                self.params = torch.nn.Parameter(torch.zeros(layout["total"] * 2))
                self.n_output_dims = 32
            else:
                assert self.otype == "SphericalHarmonics" and config["degree"] == 4
                self.n_output_dims = 16

        def forward(self, x):
            if self.otype == "HashGrid":
                return OF.hashgrid(x, self.params, layout)
            return OF.sh4(x * 2.0 - 1.0)
    sys.modules["tinycudann"].Encoding = Encoding
    sys.modules["tinycudann"].free_temporary_memory = lambda: None
    sys.modules["omegaconf"].OmegaConf.to_container = staticmethod(lambda c, resolve=True: {k: v for k, v in c.items()})
    torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    torch.cuda.empty_cache = lambda: None
    _stub("cv2")
    _stub("torchgeometry")
    _stub("torchgeometry.core")
    _stub("torchgeometry.core.conversions",
          angle_axis_to_rotation_matrix=lambda aa: torch.eye(4)[None].repeat(aa.shape[0], 1, 1))
    sys.modules["torchgeometry.core"].conversions = sys.modules["torchgeometry.core.conversions"]

    # ---- first-party CUDA extensions of fast-SNARF -> oracle restatements (pinned on the GPU box)
    import torch.utils.cpp_extension as cpp
    cpp.load = lambda *a, **k: types.SimpleNamespace()
    lib = _pkg("lib", os.path.join(REF, "lib"))
    _pkg("lib.pytorch3d", os.path.join(REF, "lib", "pytorch3d"))

    def knn_points(x, y, K=1):
        outs_v, outs_i = [], []
        for i in range(0, x.shape[1], 4096):
            d2 = torch.cdist(x[:, i:i + 4096].double(), y.double()) ** 2
            val, idx = torch.topk(d2, K, dim=-1, largest=False)
            outs_v.append(val.to(x.dtype)); outs_i.append(idx)
        return torch.cat(outs_v, 1), torch.cat(outs_i, 1), None
    _stub("lib.pytorch3d.ops", knn_points=knn_points)
    sys.modules["lib.pytorch3d"].ops = sys.modules["lib.pytorch3d.ops"]
    _pkg("models.deformers.fast_snarf", os.path.join(REF, "models", "deformers", "fast_snarf"))
    sys.modules["models.deformers.smplx"].SMPL = object
    from models.deformers.fast_snarf import deformer_torch as ref_def

    def fuse_broyden(x, xd_tgt, voxel, voxel_J, tfs, bones, align, J_inv, is_valid, offset, scale, cvg, dvg):
        assert x.shape[0] == 1 and abs(cvg - odef.CVG) < 1e-12 and abs(dvg - odef.DVG) < 1e-12
        assert bones.tolist() == odef.INIT_BONES
        ox, oJ, ov = odef.broyden(xd_tgt[0].float(), voxel_J[0], tfs[0], offset.reshape(3), scale.reshape(3))
        x[0], J_inv[0], is_valid[0] = ox, oJ, ov

    def precompute(lbs, tfs, voxel_d, voxel_J, offset, scale):
        voxel_J[0] = odef.precompute(lbs[0], tfs[0])
    ref_def.fuse_kernel = types.SimpleNamespace(fuse_broyden=fuse_broyden)
    ref_def.filter_cuda = types.SimpleNamespace(filter=lambda x, m: odef.filter_duplicates(x[0], m[0])[None])
    ref_def.precompute_cuda = types.SimpleNamespace(precompute=precompute)

    # ---- lib.nerfacc: real Python layer, kernels -> oracle C restatements (pinned on the GPU box)
    import lib.nerfacc as LN
    import lib.nerfacc.cuda as LC
    import lib.nerfacc.pack as LP

    def _rs(packed_info, starts, ends, weights, sdfs, n):
        return oops.ray_resampling(packed_info, starts, ends, weights, sdfs, n)
    LC.ray_resampling = _rs
    LC.ray_resampling_merge = lambda pi, v, il, ir, w, n: oops.ray_resampling_merge(pi, v, il, ir, w, n)
    LC.ray_resampling_fine = lambda pi, s, e, w, n: oops.ray_resampling_fine(pi, s, e, w, n)
    LC.ray_resampling_sdf_fine = lambda pi, s, e, a, sd, n: oops.ray_resampling_sdf_fine(pi, s, e, a, sd, n)
    LC.unpack_info = lambda pi, n: oops.unpack_info(pi, n)
    LC.unpack_data = lambda pi, d, n: oops.unpack_data(pi, d.float(), n).to(d.dtype)
    # pack.py refuses CPU tensors in two places (its CUDA branches are plain torch / the kernels above)
    LP.pack_info = lambda ray_indices, n_rays=None: oops.pack_info(
        ray_indices, int(ray_indices.max()) + 1 if n_rays is None else n_rays)
    LP.unpack_info = lambda packed_info, n_samples: oops.unpack_info(packed_info, n_samples)
    LN.pack_info, LN.unpack_info = LP.pack_info, LP.unpack_info

    # ---- nerfacc 0.5.3 (third party, absent) -> oracle restatements
    class RayIntervals:
        def __init__(self, vals=None, packed_info=None, ray_indices=None, is_left=None, is_right=None):
            self.vals, self.packed_info, self.ray_indices = vals, packed_info, ray_indices
            self.is_left, self.is_right = is_left, is_right

    class RaySamples:
        def __init__(self, vals=None, packed_info=None, ray_indices=None):
            self.vals, self.packed_info, self.ray_indices = vals, packed_info, ray_indices

    def traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=None, far_planes=None, step_size=1e-3, cone_angle=0.0):
        assert cone_angle == 0.0 and binaries.shape[0] == 1
        near = float(near_planes[0]) if near_planes.numel() else 0.0
        far = float(far_planes[0]) if far_planes.numel() else 1e10
        tg = oops.traverse_grid(rays_o, rays_d, binaries[0], aabbs[0], near, far, float(step_size))
        iv = RayIntervals(tg["vals"], tg["packed_info"].long(), tg["ray_indices"], tg["is_left"], tg["is_right"])
        sm = RaySamples((tg["t_starts"] + tg["t_ends"]) / 2, tg["sample_packed_info"].long(), tg["sample_ray_indices"])
        return iv, sm, None

    def render_weight_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
        if packed_info is None:
            packed_info = oops.pack_info(ray_indices, n_rays)
        return oops.render_weight_from_alpha(alphas, packed_info)

    def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
        return oops.accumulate_along_rays(weights, values, ray_indices, n_rays)

    class OccGridEstimator(torch.nn.Module):
        def __init__(self, roi_aabb, resolution=64, levels=1):
            super().__init__()
            assert levels == 1
            self.register_buffer("aabbs", torch.as_tensor(roi_aabb, dtype=torch.float32).reshape(1, 6))
            self.register_buffer("binaries", torch.zeros(1, resolution, resolution, resolution, dtype=torch.bool))
            self.register_buffer("occs", torch.zeros(resolution ** 3))

    def _na(*a, **k):
        raise NotImplementedError("not on the render path")
    na = sys.modules["nerfacc"]
    na.RayIntervals, na.OccGridEstimator, na.traverse_grids = RayIntervals, OccGridEstimator, traverse_grids
    na.render_weight_from_alpha, na.accumulate_along_rays = render_weight_from_alpha, accumulate_along_rays
    na.render_visibility_from_alpha = na.render_visibility_from_density = _na
    _stub("nerfacc.volrend", render_weight_from_density=_na, render_weight_from_alpha=render_weight_from_alpha,
          accumulate_along_rays=accumulate_along_rays)
    # the training-time grid is not on this path
    _pkg("models.occ_grid", os.path.join(REF, "models", "occ_grid"))
    _stub("models.occ_grid.temporal_occ_grid", TemporalOccGridEstimator=object)
    _pkg("models.pbr", os.path.join(REF, "models", "pbr"))
    import utils.misc as ref_misc
    ref_misc.get_rank = lambda: "cpu"


class RandTables:
    """Replaces torch.rand for the duration of a reference call: answers with queued tables whose shape matches, or with
    a provider registered for the shape's rank (``providers[ndim](shape)``)."""

    def __init__(self):
        self.queue = []
        self.providers = {}
        self._orig = None

    def push(self, t):
        self.queue.append(torch.as_tensor(t, dtype=torch.float32))

    def __enter__(self):
        self._orig = torch.rand

        def rand(*shape, **kw):
            shp = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else tuple(shape)
            for i, t in enumerate(self.queue):
                if tuple(t.shape) == shp:
                    return self.queue.pop(i).clone()
            if len(shp) in self.providers:
                return self.providers[len(shp)](shp)
            raise RuntimeError(f"reference drew torch.rand{shp} and no table of that shape was queued")
        torch.rand = rand
        return self

    def __exit__(self, *a):
        torch.rand = self._orig


def build_model(scene, frame, spp, gi=False, render_mode="light", add_emitter=False, binaries=None, env=None, u1=None, u2=None,
                material_feature="hybrid", secondary_importance_sample=True, zero_crossing_search=True):
    """The reference's IntrinsicAvatarModel with its own sub-modules, in the fully warmed-up test-time state, for one
    frame.  ``scene`` = tests/conftest.Scene (synthetic body, our state dict); ``frame`` = scene.frame(idx)."""
    install()
    from models import network_utils as ref_net
    ref_net.get_rank = lambda: "cpu"
    import models.base as ref_base
    ref_base.get_rank = lambda: "cpu"
    from models.rf import geometry as ref_geo, radiance as ref_rad, density as ref_den
    ref_geo.get_rank = lambda: "cpu"
    from models.pbr import material as ref_mat
    from models.deformers import deformer as ref_comp, non_rigid_deformer as ref_nr, snarf_deformer as ref_sd
    from models.deformers.fast_snarf import deformer_torch as ref_def
    from lib.torch_pbr import bxdf as ref_bxdf, light as ref_light
    import models.intrinsic_avatar as ref_ia

    gcfg = load_yaml("geometry/progressive_hash_grid.yaml", **{"${model.radius}": 1.0})
    gcfg["isosurface"] = None
    rcfg = load_yaml("radiance/progressive_hash_grid.yaml", **{"${add:${model.geometry.feature_dim}, 3}": 16})
    mat_in = {"hybrid": 48, "geometry": 13, "radiance": 35}[material_feature]      # models/intrinsic_avatar.py:1102-1113
    mcfg = load_yaml("material/shallow_mlp.yaml", **{"${add:${model.geometry.feature_dim}, 35}": mat_in})
    geo, rad, mat = ref_geo.VolumeSDF(gcfg), ref_rad.VolumeRefDirRadiance(rcfg), ref_mat.VolumeMaterial(mcfg)
    den = ref_den.LearnedLaplaceDensity(cfg({"params_init": {"beta": 0.1}, "beta_min": 0.0001}))
    sd = scene.state_dict
    sub = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    geo.load_state_dict(sub("geometry."), strict=True)
    rad.load_state_dict(sub("radiance."), strict=False)
    msd = sub("material.")
    if material_feature != "hybrid":
        # the material net then sees only the geometry feature (13) / only the radiance encoding (35): the matching
        # columns of the 48-wide first layer (weights.material_state_dict_for)
        from intrinsicavatar_b200.weights import material_state_dict_for
        msd = {k[len("material."):]: v for k, v in material_state_dict_for(sd, material_feature).items() if k.startswith("material.")}
    mat.load_state_dict(msd, strict=False)
    for i in range(3):
        mat.network.weights_per_layer[i].data.copy_(msd[f"network.layers.{i}.weight"])
        mat.network.biases_per_layer[i].data.copy_(msd[f"network.layers.{i}.bias"])
        mat.network.lipshitz_bound_per_layer[i].data.copy_(msd[f"network.lipshitz_bound_per_layer.{i}"].reshape(-1))
    den.beta.data.copy_(sd["density.beta"])
    geo.encoding.encoding.update_step(250, 25000)
    rad.xyz_encoding.encoding.update_step(250, 25000)
    rad.sh_mask[:] = 1.0

    # ---- deformer: the reference's classes, per-frame state from SnarfSetup (pinned separately)
    snarf = scene.snarf
    fd = ref_def.ForwardDeformer.__new__(ref_def.ForwardDeformer)
    torch.nn.Module.__init__(fd)
    fd.opt, fd.soft_blend, fd.global_scale, fd.version = {}, 20, 1.2, 1
    fd.init_bones = list(ref_def_init_bones())
    fd.init_bones_cuda = torch.tensor(fd.init_bones).int()
    fd.device = torch.device("cpu")
    body = snarf.body
    from intrinsicavatar_b200.body import a_pose
    cano = body(body_pose=a_pose())
    # switch_to_explicit defines the query_weights closure and the kernels; its KNN voxel grid is then replaced by the
    # SnarfSetup grid both sides use (the two differ only at KNN near-ties, see test_lbs_voxelisation_matches_reference)
    fd.switch_to_explicit(resolution=8, smpl_verts=torch.from_numpy(cano["vertices"]),
                          smpl_weights=torch.from_numpy(body.lbs_weights)[None], use_smpl=True)
    fd.resolution = snarf.resolution
    fd.lbs_voxel_final = torch.from_numpy(snarf.lbs_voxel)[None]
    assert torch.allclose(fd.offset_kernel.reshape(3), torch.from_numpy(snarf.offset_kernel), atol=1e-6)
    assert torch.allclose(fd.scale_kernel.reshape(3), torch.from_numpy(snarf.scale_kernel), rtol=1e-6)
    rigid = ref_sd.SNARFDeformer.__new__(ref_sd.SNARFDeformer)
    rigid.deformer = fd
    rigid.opt = types.SimpleNamespace(use_j_inv=False, cano_pose="a_pose", resolution=snarf.resolution, optimize_betas=False)
    rigid.initialized, rigid.dtype = True, torch.float32
    rigid.tfs = torch.from_numpy(frame["tfs"])[None]
    rigid.w2s = torch.from_numpy(frame["w2s"])[None]
    rigid.vertices = torch.from_numpy(frame["vertices"])[None]
    rigid.bbox = torch.from_numpy(snarf.bbox)
    rigid.smpl_outputs = types.SimpleNamespace(betas=torch.zeros(1, 10))
    rigid.rot_mats, rigid.basic_joints = None, None
    fd.precompute(rigid.tfs)
    comp = ref_comp.SNARFDeformer.__new__(ref_comp.SNARFDeformer)
    torch.nn.Module.__init__(comp)
    comp.rigid_deformer = rigid
    comp.non_rigid_deformer = ref_nr.DummyNonRigidDeformer.__new__(ref_nr.DummyNonRigidDeformer)
    torch.nn.Module.__init__(comp.non_rigid_deformer)

    # ---- the model
    m = ref_ia.IntrinsicAvatarModel.__new__(ref_ia.IntrinsicAvatarModel)
    torch.nn.Module.__init__(m)
    m.rank = "cpu"
    m.config = cfg({"global_illumination": gi, "render_mode": render_mode, "resample_light": True,
                    "grid_prune_occ_thre": 0.001, "light": {"name": "envlight-tensor"}, "add_emitter": add_emitter,
                    "learned_background": False, "ray_chunk": 4096})
    m.geometry, m.density, m.radiance, m.material, m.deformer = geo, den, rad, mat, comp
    m.pose_encoder = lambda *a, **k: None
    m.cond = None
    m.scatterer = ref_bxdf.MultiLobe(types.SimpleNamespace())
    m.scatterer.train(False)
    ecfg = types.SimpleNamespace(xyz2lonlat_mode=None,
                                 envlight_config=types.SimpleNamespace(scale=1.0, bias=0.0, base_res=8, hdr_filepath=None))
    m.emitter = ref_light.EnvironmentLightTensor(ecfg)
    m.emitter.train(False)
    m.material_feature = material_feature
    aabb = torch.tensor([-1.25, -1.55, -1.25, 1.25, 0.95, 1.25])
    m.register_buffer("scene_aabb", aabb)
    m.randomized, m.background_color, m.samples_per_pixel = False, torch.ones(3), spp
    m.render_step_size = torch.norm(aabb[3:] - aabb[:3]).item() / 128
    m.num_samples_per_secondary_ray, m.secondary_near_plane, m.secondary_far_plane = 64, 0.0, 1.5
    m.secondary_shader_chunk, m.secondary_importance_sample = 160000, secondary_importance_sample
    m.enable_phys, m.importance_sample, m.add_emitter, m.zero_crossing_search = True, True, add_emitter, zero_crossing_search
    m.albedo_only, m.t_idx = False, 0.0
    m.train(False)
    geo.prepare_bbox(rigid.bbox)
    rad.prepare_bbox(rigid.bbox)
    if binaries is not None:
        occ = sys.modules["nerfacc"].OccGridEstimator(roi_aabb=torch.as_tensor(frame["deformed_bbox"]),
                                                      resolution=binaries.shape[-1], levels=1)
        occ.binaries = torch.as_tensor(binaries).reshape(1, *binaries.shape[-3:]).bool()
        m.occupancy_grid_test = occ
    if env is not None:
        m.emitter.base = torch.nn.Parameter(torch.as_tensor(env, dtype=torch.float32))
        m.emitter.pdf_scale = (m.emitter.base.shape[0] * m.emitter.base.shape[1]) / (2 * np.pi * np.pi)
        m.emitter.update_pdf()
        with RandTables() as rt:
            rt.push(u1); rt.push(u2)
            m.secondary_rays_d = m.emitter.sample(spp)
    return m


def ref_def_init_bones():
    return [0, 1, 2, 4, 5, 10, 11, 12, 15, 16, 17, 18, 19]   # deformer_torch.py:27 (ForwardDeformer.__init__ needs .cuda())


def forward(m, rays, seed=0, ray_offset=0, public=False):
    """The reference's forward_ on [n,8] world-space rays with the randomness product and oracle use: the keyed per-ray
    light permutation (light / uniform_light) and the counter-based uniforms of MultiLobe.sample / emitter.sample (mats / mis)."""
    import models.intrinsic_avatar as ref_ia
    from oracle.pbr import kensler_permute, pixel_key, rng_uniform
    rays = torch.as_tensor(rays, dtype=torch.float32)
    n, spp = rays.shape[0], m.samples_per_pixel
    with RandTables() as rt:
        if m.config.render_mode in ("light", "uniform_light"):
            # argsort(table[r]) must be the keyed permutation of ray r (models/intrinsic_avatar.py:1355-1364)
            table = np.zeros((n, spp), np.float32)
            j = np.arange(spp, dtype=np.uint64)
            for r in range(n):
                perm = kensler_permute(j, spp, pixel_key(seed, np.full(spp, r + ray_offset, np.int64)))
                table[r, perm] = (np.arange(spp) + 0.5) / spp
            rt.push(table)
        else:
            # MultiLobe.sample draws torch.rand(N, 2), emitter.sample torch.rand(N) twice, N = foreground shading samples in
            # resampled order: sample j of ray r gets the stream (seed, r, j, dim) -- needs (r, j) of every foreground sample
            state = {"calls": 0}
            orig_svi = ref_ia.sample_volume_interaction

            def svi(*a, **k):
                res = orig_svi(*a, **k)
                resampled_packed_info, resampled_ray_indices, fg_indices = res[0], res[1], res[3]
                fg_ray = resampled_ray_indices[fg_indices]
                state["j"] = (fg_indices - resampled_packed_info[fg_ray, 0].long()).numpy().astype(np.uint64)
                state["key"] = pixel_key(seed, fg_ray.numpy() + ray_offset)
                return res
            ref_ia.sample_volume_interaction = svi
            rt.providers[2] = lambda shp: torch.stack([rng_uniform(state["key"], state["j"], 0),
                                                       rng_uniform(state["key"], state["j"], 1)], -1)

            def light_uniform(shp):
                state["calls"] += 1
                return rng_uniform(state["key"], state["j"], 1 + state["calls"])      # dims 2, 3
            rt.providers[1] = light_uniform
        try:
            with torch.no_grad():
                out = m.forward(rays) if public else m.forward_(rays)   # public: chunk_batch wrapper + "beta" (:1653-1666)
        finally:
            if m.config.render_mode not in ("light", "uniform_light"):
                ref_ia.sample_volume_interaction = orig_svi
    return out
