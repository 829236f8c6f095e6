#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz by running the REFERENCE's own Python modules on CPU.

Runs only in the build container (needs /root/reference); the .npz is committed and is what the
`-m "not gpu"` tests (tests/test_oracle_golden.py) hold the oracle to.  Nothing is copied from the
reference: its modules are imported from where they lie, with the absent third-party packages
(tinycudann, pytorch_lightning, omegaconf, nerfacc, imageio, pyexr) replaced by empty stub modules
and `device="cuda"` factory calls redirected to the CPU.

What is pinned (reference file:line):
  bxdf       MultiLobe.eval (Lambertian + GGX)                    lib/torch_pbr/bxdf.py:111-146,217-265,321-330
  envlight   EnvironmentLightTensor.update_pdf/sample/pdf/eval    lib/torch_pbr/light.py:221-446
  srgb       rgb_to_srgb                                          lib/torch_pbr/utils/nvdiffrecmc_util.py:94-102
  mlp        VanillaMLP (weight-norm, sphere init, softplus100),  models/network_utils.py:201-244
             VanillaMLP (ReLU), LipshitzMLP                       models/network_utils.py:360-428
  density    LearnedLaplaceDensity.density_func                   models/rf/density.py:25-34
  cc         max_connected_component                              models/utils.py:152-163
  lbs        batch_rodrigues, batch_rigid_transform               models/deformers/smplx/lbs.py:345-401, 152-248
  reflect    reflect(), get_activation                            models/utils.py
and, in tests/golden/reference_vectors_bsdf.npz (`make_golden.py bsdf`; the integrators of SURVEY 8f.1):
  bsdf       MultiLobe.sample (explicit `sample` uniforms), MultiLobe.pdf, eval mode
                                                                  lib/torch_pbr/bxdf.py:290-388
  sphere     EnvironmentLightBase.sample_uniform_sphere_stratified(1, 16, 32), eval mode
                                                                  lib/torch_pbr/light.py:161-217
and, in tests/golden/reference_vectors_voxel.npz (`make_golden.py voxel`):
  deform_train  ForwardDeformer.forward, training mode (implicit-differentiation correction) + autograd to the bone transforms
                                                                  models/deformers/fast_snarf/deformer_torch.py:57-76
  voxel      ForwardDeformer.switch_to_explicit + query_weights_smpl (skinning-weight voxel grid, offset / scale kernels)
                                                                  models/deformers/fast_snarf/deformer_torch.py:139-197, 234-253
and, in tests/golden/reference_vectors_e2e.npz (`make_golden.py e2e`, through scripts/ref_harness.py):
  e2e        IntrinsicAvatarModel.forward_ and _compute_occupancy_grid themselves, six frames: render_mode light / mats /
             mis / uniform_light, global illumination, add_emitter          models/intrinsic_avatar.py:307-362, 396-1651
and, in tests/golden/reference_vectors_fields.npz (`make_golden.py fields`):
  fields     VolumeSDF (autograd gradient), VolumeRefDirRadiance, VolumeMaterial with the reference's yaml configs and our
             state dict; tcnn.Encoding replaced by the oracle's hash grid / SH      models/rf/geometry.py:124-235,
                                                                  models/rf/radiance.py:82-135, models/pbr/material.py:31-51
and, in tests/golden/reference_vectors_snarf.npz (`make_golden.py snarf`, after `smpl`):
  snarf      SNARFDeformer.initialize + prepare_deformer, get_bbox_from_smpl    models/deformers/snarf_deformer.py:24-35, 46-126
and, in tests/golden/reference_vectors_smpl.npz (`make_golden.py smpl`):
  smpl       lbs() + SMPL.forward translation on a random model of SMPL's shapes
                                                                  models/deformers/smplx/lbs.py:152-248, body_models.py:342-358
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
OUT_BSDF = os.path.join(ROOT, "tests", "golden", "reference_vectors_bsdf.npz")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def install_stubs():
    _stub("tinycudann")
    _stub("imageio")
    _stub("pyexr")
    _stub("nerfacc")
    _stub("omegaconf", OmegaConf=type("OmegaConf", (), {"register_new_resolver": staticmethod(lambda *a, **k: None)}))
    _stub("pytorch_lightning")
    _stub("pytorch_lightning.utilities")
    _stub("pytorch_lightning.utilities.rank_zero", rank_zero_debug=lambda *a, **k: None,
          rank_zero_info=lambda *a, **k: None, rank_zero_only=lambda f: f)
    # packages whose __init__ pulls the whole framework in: expose only their directories
    models = _pkg("models", os.path.join(REF, "models"))
    models.register = lambda name: (lambda cls: cls)
    _pkg("models.deformers", os.path.join(REF, "models", "deformers"))
    _pkg("models.deformers.smplx", os.path.join(REF, "models", "deformers", "smplx"))
    _pkg("models.rf", os.path.join(REF, "models", "rf"))
    _pkg("utils", os.path.join(REF, "utils"))
    _pkg("systems", os.path.join(REF, "systems"))
    sys.path.insert(0, REF)
    # device="cuda" -> CPU for tensor factories and .cuda()
    for fn in ("rand", "tensor", "arange", "zeros", "ones", "empty", "full", "linspace", "randn", "eye"):
        orig = getattr(torch, fn)

        def wrap(*a, __o=orig, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return __o(*a, **k)
        setattr(torch, fn, wrap)
    torch.Tensor.cuda = lambda self, *a, **k: self


def main():
    install_stubs()
    torch.manual_seed(1234)
    g = {}

    # ------------------------------------------------------------------ bxdf (MultiLobe.eval)
    from lib.torch_pbr import bxdf as ref_bxdf
    cfg = types.SimpleNamespace(get=lambda k, d=None: d)
    N = 512
    n = torch.nn.functional.normalize(torch.randn(N, 3), dim=-1)
    wi = torch.nn.functional.normalize(n + 0.8 * torch.randn(N, 3), dim=-1)
    wo = torch.nn.functional.normalize(n + 0.8 * torch.randn(N, 3), dim=-1)
    rough = torch.rand(N, 1) * 0.9 + 0.09
    albedo = torch.rand(N, 3) * 0.77 + 0.03
    metal = torch.rand(N, 1)
    lobe = ref_bxdf.MultiLobe(types.SimpleNamespace())
    # called exactly as pbr_light_forward does (models/intrinsic_avatar.py:816-825)
    out = lobe.eval(wi=wi, n=n, wo=wo, alpha_x=rough.squeeze(-1), alpha_y=rough.squeeze(-1), albedo=albedo,
                    metallic=metal, attenuation=torch.zeros(N, 1))
    diff, spec = out[0], out[1]
    g.update(bxdf_wi=wi, bxdf_n=n, bxdf_wo=wo, bxdf_rough=rough, bxdf_albedo=albedo, bxdf_metal=metal,
             bxdf_diff=diff, bxdf_spec=spec)

    # ------------------------------------------------------------------ env light
    from lib.torch_pbr import light as ref_light
    ecfg = types.SimpleNamespace(xyz2lonlat_mode=None,
                                 envlight_config=types.SimpleNamespace(scale=1.0, bias=0.0, base_res=8, hdr_filepath=None))
    env = ref_light.EnvironmentLightTensor(ecfg)
    H, W = 24, 48
    base = torch.rand(H, W, 3) ** 4 * 20.0           # HDR-like dynamic range
    base[5, 7] = torch.tensor([900.0, 700.0, 500.0])  # a "sun"
    base[20:, :, :] = 0.0                              # rows with zero radiance (pdf floor 1e-6)
    env.base.data = base
    env.pdf_scale = (H * W) / (2 * np.pi * np.pi)
    env.update_pdf()
    env.train(False)
    n_s = 256
    torch.manual_seed(77)
    u1, u2 = torch.rand(n_s), torch.rand(n_s)
    torch.manual_seed(77)
    dirs = env.sample(n_s)
    more = torch.nn.functional.normalize(torch.randn(300, 3), dim=-1)
    alld = torch.cat([dirs, more])
    g.update(env_base=base, env_pdf_table=env._pdf, env_rows=env.rows, env_cols=env.cols, env_u1=u1, env_u2=u2,
             env_dirs=dirs, env_query_dirs=alld, env_pdf=env.pdf(alld), env_eval=ref_light.EnvironmentLightTensor.eval(env, alld))

    # ------------------------------------------------------------------ sRGB
    from lib.torch_pbr.utils import nvdiffrecmc_util as ref_util
    ramp = torch.cat([torch.linspace(-0.1, 0.01, 50), torch.linspace(0.0, 4.0, 200)])[:, None].repeat(1, 3)
    g.update(srgb_in=ramp, srgb_out=ref_util.rgb_to_srgb(ramp))

    # ------------------------------------------------------------------ MLPs with OUR random state dict
    sys.path.insert(0, ROOT)
    from intrinsicavatar_b200.weights import random_state_dict
    from models import network_utils as ref_net
    sd = random_state_dict(0)

    def sub(prefix):
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}

    geo = ref_net.VanillaMLP(35, 13, {"n_neurons": 64, "n_hidden_layers": 1, "sphere_init": True,
                                      "sphere_init_radius": 0.5, "weight_norm": True, "output_activation": "none"})
    geo.load_state_dict(sub("geometry.network."), strict=True)
    rad = ref_net.VanillaMLP(67, 3, {"n_neurons": 64, "n_hidden_layers": 2, "output_activation": "none"})
    rad.load_state_dict(sub("radiance.network."), strict=True)
    mat = ref_net.LipshitzMLP(48, 5, {"n_neurons": 64, "n_hidden_layers": 2, "output_activation": "none"})
    msd = sub("material.network.")
    mat.load_state_dict({k: v for k, v in msd.items()}, strict=False)
    # LipshitzMLP shares its Linear weights with weights_per_layer: make sure both views carry our values
    for i in range(3):
        mat.weights_per_layer[i].data.copy_(msd[f"layers.{i}.weight"])
        mat.biases_per_layer[i].data.copy_(msd[f"layers.{i}.bias"])
        mat.lipshitz_bound_per_layer[i].data.copy_(msd[f"lipshitz_bound_per_layer.{i}"].reshape(-1))
    x35 = torch.cat([torch.rand(400, 3) * 2 - 1, torch.randn(400, 32) * 0.1], 1)
    x67 = torch.cat([torch.rand(400, 3) * 2 - 1, torch.randn(400, 32) * 1e-2, torch.randn(400, 13) * 0.3,
                     torch.randn(400, 16), torch.nn.functional.normalize(torch.randn(400, 3), dim=-1)], 1)
    x48 = torch.cat([torch.rand(400, 3) * 2 - 1, torch.randn(400, 32) * 1e-2, torch.randn(400, 13) * 0.3], 1)
    with torch.no_grad():
        g.update(mlp_geo_in=x35, mlp_geo_out=geo(x35), mlp_rad_in=x67, mlp_rad_out=rad(x67),
                 mlp_mat_in=x48, mlp_mat_out=mat(x48))

    # ------------------------------------------------------------------ Laplace density
    import importlib.util
    models_base = _stub("models.base", BaseModel=type("BaseModel", (torch.nn.Module,), {}))
    spec_ = importlib.util.spec_from_file_location("ref_density", os.path.join(REF, "models/rf/density.py"))
    dens = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(dens)
    sdf = torch.cat([torch.linspace(-0.2, 0.2, 401), torch.tensor([0.0, 1e5, -1e-7, 1e-7])])
    beta = torch.tensor(0.01) + 1e-4
    sigma = dens.LearnedLaplaceDensity.density_func(None, sdf, beta=beta)
    g.update(density_sdf=sdf, density_beta=beta, density_sigma=sigma)

    # ------------------------------------------------------------------ connected components, reflect
    spec_u = importlib.util.spec_from_file_location("ref_models_utils", os.path.join(REF, "models/utils.py"))
    mu = importlib.util.module_from_spec(spec_u)
    spec_u.loader.exec_module(mu)
    grid = torch.zeros(16, 16, 16, dtype=torch.bool)
    grid[2:7, 3:9, 4:8] = True          # big blob
    grid[10:12, 10:12, 10:12] = True    # small blob
    grid[14, 14, 14] = True             # single voxel
    grid[7, 8, 7] = True                # diagonal neighbour of the big blob (26-connectivity joins it)
    g.update(cc_grid=grid, cc_labels=mu.max_connected_component(grid[None]))
    v = torch.nn.functional.normalize(torch.randn(64, 3), dim=-1)
    nn_ = torch.nn.functional.normalize(torch.randn(64, 3), dim=-1)
    g.update(reflect_v=v, reflect_n=nn_, reflect_out=mu.reflect(v, nn_))

    # ------------------------------------------------------------------ SMPL skeleton maths
    from models.deformers.smplx import lbs as ref_lbs
    rv = torch.randn(24, 3) * 0.6
    rv[3] = 0.0                          # zero rotation edge case
    R = ref_lbs.batch_rodrigues(rv)
    from intrinsicavatar_b200.body import SyntheticBody
    body = SyntheticBody()
    from intrinsicavatar_b200.body import PARENTS
    joints = torch.from_numpy(np.asarray(body.joints_rest, np.float32))[None]
    parents = torch.from_numpy(np.asarray(PARENTS, np.int64))
    posed_joints, A = ref_lbs.batch_rigid_transform(R[None], joints, parents)
    g.update(lbs_rvec=rv, lbs_rotmats=R, lbs_joints=joints[0], lbs_parents=parents, lbs_posed_joints=posed_joints[0],
             lbs_A=A[0])

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in g.items()})
    print("wrote", OUT, {k: tuple(np.asarray(v.detach() if torch.is_tensor(v) else v).shape) for k, v in g.items()})


def main_bsdf():
    """MultiLobe.sample / .pdf and the stratified sphere directions of render_mode = mats | mis | uniform_light."""
    install_stubs()
    torch.manual_seed(4321)
    g = {}
    from lib.torch_pbr import bxdf as ref_bxdf
    from lib.torch_pbr import light as ref_light
    N = 2048
    n = torch.nn.functional.normalize(torch.randn(N, 3), dim=-1)
    wi = torch.nn.functional.normalize(n + 0.8 * torch.randn(N, 3), dim=-1)
    wi[:64] = torch.nn.functional.normalize(-n[:64] + 0.3 * torch.randn(64, 3), dim=-1)   # viewer below the surface
    rough = torch.rand(N, 1) * 0.9 + 0.09
    albedo = torch.rand(N, 3) * 0.77 + 0.03
    metal = torch.rand(N, 1)
    metal[64:128] = 1.0
    metal[128:192] = 0.0
    sample = torch.rand(N, 2)
    lobe = ref_bxdf.MultiLobe(types.SimpleNamespace())
    lobe.train(False)   # (.eval is shadowed by the BRDF eval)
    kw = dict(alpha_x=rough.squeeze(-1), alpha_y=rough.squeeze(-1), albedo=albedo, metallic=metal,
              attenuation=torch.zeros(N, 1))
    # called as pbr_mats_forward does (models/intrinsic_avatar.py:880-903), uniforms made explicit
    wo = lobe.sample(n=n, wi=wi, sample=sample.clone(), **kw)
    pdf = lobe.pdf(n=n, wi=wi, wo=wo, **kw)
    wo2 = torch.nn.functional.normalize(n + 0.8 * torch.randn(N, 3), dim=-1)   # pdf of directions it did not sample (MIS)
    pdf2 = lobe.pdf(n=n, wi=wi, wo=wo2, **kw)
    g.update(bsdf_n=n, bsdf_wi=wi, bsdf_rough=rough, bsdf_albedo=albedo, bsdf_metal=metal, bsdf_sample=sample,
             bsdf_wo=wo, bsdf_pdf=pdf, bsdf_wo2=wo2, bsdf_pdf2=pdf2)
    ecfg = types.SimpleNamespace(xyz2lonlat_mode=None,
                                 envlight_config=types.SimpleNamespace(scale=1.0, bias=0.0, base_res=8, hdr_filepath=None))
    env = ref_light.EnvironmentLightTensor(ecfg)
    env.train(False)
    dirs, inv_pdf = env.sample_uniform_sphere_stratified(1, 16, 32, device="cpu")
    g.update(sphere_dirs=dirs, sphere_inv_pdf=inv_pdf)
    np.savez_compressed(OUT_BSDF, **{k: v.detach().cpu().numpy() for k, v in g.items()})
    print("wrote", OUT_BSDF, {k: tuple(v.shape) for k, v in g.items()})


def main_smpl():
    """The reference's own lbs() (models/deformers/smplx/lbs.py:152-248) + SMPL.forward's translation handling
    (body_models.py:342-358) on a RANDOM model with SMPL's array shapes (the licensed .pkl is absent): pins
    intrinsicavatar_b200.body.SMPLBody."""
    install_stubs()
    sys.path.insert(0, ROOT)
    torch.manual_seed(99)
    from models.deformers.smplx import lbs as ref_lbs
    from intrinsicavatar_b200.body import PARENTS, SyntheticBody
    V = 400
    body = SyntheticBody(n_verts=V)
    v_template = torch.from_numpy(body.v_template).double()
    shapedirs = (torch.randn(V, 3, 10) * 0.01).double()              # float32-representable: stored as float32
    posedirs = (torch.randn(207, V * 3) * 0.002).double()              # stored [P, V*3] as body_models.py:156-159 does
    # a joint regressor with the right flavour: convex weights over the vertices nearest to each rest joint
    d = torch.cdist(torch.from_numpy(body.joints_rest).double(), v_template)
    Jr = torch.softmax(-d / 0.03, dim=1).float().double()
    w = torch.from_numpy(body.lbs_weights).double()
    parents = torch.from_numpy(np.asarray(PARENTS, np.int64))
    betas = torch.randn(1, 10).double()
    pose = (torch.randn(1, 72) * 0.4).double()
    transl = torch.tensor([[0.1, -0.2, 3.0]]).double()
    verts, joints, A, T, so, po = ref_lbs.lbs(betas, pose, v_template[None], shapedirs, posedirs, Jr, parents, w, pose2rot=True)
    A = A.clone()
    A[..., :3, 3] += transl.unsqueeze(1)
    g = dict(smpl_v_template=v_template, smpl_shapedirs=shapedirs, smpl_posedirs=posedirs, smpl_J_regressor=Jr,
             smpl_weights=w, smpl_betas=betas, smpl_pose=pose, smpl_transl=transl,
             smpl_vertices=verts + transl.unsqueeze(1), smpl_joints=joints + transl.unsqueeze(1), smpl_A=A)
    out = os.path.join(ROOT, "tests", "golden", "reference_vectors_smpl.npz")
    big = ("smpl_shapedirs", "smpl_posedirs", "smpl_J_regressor", "smpl_weights", "smpl_v_template")
    np.savez_compressed(out, **{k: v.detach().cpu().numpy().astype(np.float32 if k in big else np.float64) for k, v in g.items()})
    print("wrote", out, {k: tuple(v.shape) for k, v in g.items()})


def main_voxel():
    """The reference's own skinning-weight voxelisation (ForwardDeformer.switch_to_explicit + query_weights_smpl,
    models/deformers/fast_snarf/deformer_torch.py:139-197, 234-253) on the synthetic body's A-pose vertices, at
    resolution 32.  Its two non-Python dependencies are replaced: the JIT-compiled CUDA extensions are not needed on
    this path (torch.utils.cpp_extension.load is stubbed), and pytorch3d's knn_points (third party, compiled) by a
    brute-force torch K-nearest-neighbours with the same return convention (squared distances, indices)."""
    install_stubs()
    sys.path.insert(0, ROOT)
    import torch.utils.cpp_extension as cpp
    cpp.load = lambda *a, **k: types.SimpleNamespace()

    def knn_points(x, y, K=1):
        d2 = torch.cdist(x.double(), y.double()) ** 2
        val, idx = torch.topk(d2, K, dim=-1, largest=False)
        return val.to(x.dtype), idx, None
    _pkg("lib", os.path.join(REF, "lib"))
    _pkg("lib.pytorch3d", os.path.join(REF, "lib", "pytorch3d"))
    _stub("lib.pytorch3d.ops", knn_points=knn_points)
    sys.modules["lib.pytorch3d"].ops = sys.modules["lib.pytorch3d.ops"]
    _pkg("models.deformers.fast_snarf", os.path.join(REF, "models", "deformers", "fast_snarf"))
    from models.deformers.fast_snarf import deformer_torch as ref_def
    from intrinsicavatar_b200.body import SyntheticBody, a_pose
    body = SyntheticBody()
    cano = body(body_pose=a_pose())
    verts = torch.from_numpy(cano["vertices"])                   # [1,V,3]
    weights = torch.from_numpy(body.lbs_weights)[None]           # [1,V,24]
    fd = ref_def.ForwardDeformer.__new__(ref_def.ForwardDeformer)
    torch.nn.Module.__init__(fd)
    fd.global_scale = 1.2                                        # deformer_torch.py:32
    fd.device = torch.device("cpu")
    fd.switch_to_explicit(resolution=32, smpl_verts=verts, smpl_weights=weights, use_smpl=True)
    g = dict(voxel_res=np.int64(32), voxel_lbs=fd.lbs_voxel_final[0], voxel_offset_kernel=fd.offset_kernel.reshape(3),
             voxel_scale_kernel=fd.scale_kernel.reshape(3), voxel_bbox=fd.bbox)
    out = os.path.join(ROOT, "tests", "golden", "reference_vectors_voxel.npz")
    np.savez_compressed(out, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in g.items()})
    print("wrote", out, {k: tuple(np.shape(v)) for k, v in g.items()})


def main_deform_train():
    """The reference's own training-mode ForwardDeformer.forward (version 1: implicit-differentiation correction,
    models/deformers/fast_snarf/deformer_torch.py:57-76, with forward_skinning / skinning_mask / query_weights :127-137,
    199-227) and autograd through it, on the voxel grid of main_voxel.  Only ``search`` -- the CUDA Broyden kernel -- is
    replaced: it returns prescribed roots, validity flags and inverse Jacobians, which is all the correction reads."""
    install_stubs()
    sys.path.insert(0, ROOT)
    import torch.utils.cpp_extension as cpp
    cpp.load = lambda *a, **k: types.SimpleNamespace()

    def knn_points(x, y, K=1):
        d2 = torch.cdist(x.double(), y.double()) ** 2
        val, idx = torch.topk(d2, K, dim=-1, largest=False)
        return val.to(x.dtype), idx, None
    _pkg("lib", os.path.join(REF, "lib"))
    _pkg("lib.pytorch3d", os.path.join(REF, "lib", "pytorch3d"))
    _stub("lib.pytorch3d.ops", knn_points=knn_points)
    sys.modules["lib.pytorch3d"].ops = sys.modules["lib.pytorch3d.ops"]
    _pkg("models.deformers.fast_snarf", os.path.join(REF, "models", "deformers", "fast_snarf"))
    from models.deformers.fast_snarf import deformer_torch as ref_def
    from intrinsicavatar_b200.body import SyntheticBody, a_pose
    body = SyntheticBody()
    cano = body(body_pose=a_pose())
    verts = torch.from_numpy(cano["vertices"])
    weights = torch.from_numpy(body.lbs_weights)[None]
    fd = ref_def.ForwardDeformer.__new__(ref_def.ForwardDeformer)
    torch.nn.Module.__init__(fd)
    fd.global_scale = 1.2
    fd.version = 1
    fd.device = torch.device("cpu")
    fd.init_bones = [0, 1, 2, 4, 5, 10, 11, 12, 15, 16, 17, 18, 19]
    fd.switch_to_explicit(resolution=32, smpl_verts=verts, smpl_weights=weights, use_smpl=True)
    g = torch.Generator().manual_seed(21)
    n = 300
    bb = fd.bbox.reshape(2, 3) if torch.is_tensor(fd.bbox) else torch.as_tensor(np.asarray(fd.bbox), dtype=torch.float32).reshape(2, 3)
    lo, hi = bb[0] - 0.1 * (bb[1] - bb[0]), bb[1] + 0.1 * (bb[1] - bb[0])      # some roots outside the grid: border padding
    xc_opt = lo + torch.rand(1, n, 13, 3, generator=g) * (hi - lo)
    valid = torch.rand(1, n, 13, generator=g) < 0.4
    J_inv = torch.eye(3)[None, None, None] + 0.3 * torch.randn(1, n, 13, 3, 3, generator=g)
    # 24 rigid-ish bone transforms
    A = torch.randn(24, 3, 3, generator=g)
    Q, _ = torch.linalg.qr(A)
    tfs = torch.zeros(1, 24, 4, 4)
    tfs[0, :, :3, :3] = Q
    tfs[0, :, :3, 3] = 0.2 * torch.randn(24, 3, generator=g)
    tfs[0, :, 3, 3] = 1.0
    tfs.requires_grad_(True)
    G = torch.randn(1, n, 13, 3, generator=g)
    fd.search = lambda xd, cond, tfs_, eval_mode=False: (xc_opt.clone(), {"result": xc_opt.clone(), "valid_ids": valid.clone(),
                                                                         "J_inv": J_inv.clone()})
    xd = torch.zeros(1, n, 3)
    xc, others = fd.forward(xd, None, tfs, eval_mode=False)
    (xc * G).sum().backward()
    out = dict(xc_opt=xc_opt[0], valid=valid[0], J_inv=J_inv[0], tfs=tfs.detach()[0], g_xc=G[0], xc=xc.detach()[0],
               g_tfs=tfs.grad[0], fwd_tfs=others["fwd_tfs"].detach()[0])
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors_deform_train.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("wrote", path, {k: tuple(v.shape) for k, v in out.items()}, "valid roots", int(valid.sum()),
          "|g_tfs|", float(tfs.grad.norm()))


def main_fields():
    """The reference's own field modules -- VolumeSDF (models/rf/geometry.py:124-235, autograd gradient),
    VolumeRefDirRadiance (models/rf/radiance.py:82-135), VolumeMaterial (models/pbr/material.py:31-51) with the
    'hybrid' feature assembly of rgb_normal_mats_alpha_fn (models/intrinsic_avatar.py:1100-1112) -- built from the
    reference's own yaml configs and loaded with OUR state dict.  tiny-cuda-nn is absent, so `tcnn.Encoding` is a torch
    module around the oracle's hashgrid / sh4 (the only third-party arithmetic): what this pins is
    everything first-party around it -- input scaling, include_xyz, the order in which features are concatenated into
    each MLP (which a real checkpoint depends on), level / SH masks, activations, material scales, the autograd gradient."""
    import contextlib
    import yaml
    install_stubs()
    sys.path.insert(0, ROOT)
    from oracle import fields as OF
    from intrinsicavatar_b200.weights import hashgrid_layout, random_state_dict
    layout = hashgrid_layout()

    class Encoding(torch.nn.Module):           # stand-in for tcnn.Encoding(n_input_dims, config, dtype)
        def __init__(self, n_input_dims, config, dtype=torch.float32):
            super().__init__()
            self.n_input_dims, self.otype = n_input_dims, config["otype"]
            if self.otype == "HashGrid":
                assert (config["n_levels"], config["n_features_per_level"], config["log2_hashmap_size"],
                        config["base_resolution"]) == (16, 2, 19, 16) and config["interpolation"] == "Linear"
                self.params = torch.nn.Parameter(torch.zeros(layout["total"] * 2))
                self.n_output_dims = 32
            else:
                assert self.otype == "SphericalHarmonics" and config["degree"] == 4
                self.n_output_dims = 16

        def forward(self, x):
            if self.otype == "HashGrid":
                return OF.hashgrid(x, self.params, layout)
            return OF.sh4(x * 2.0 - 1.0)     # tcnn maps its [0,1] input to [-1,1]
    sys.modules["tinycudann"].Encoding = Encoding
    sys.modules["tinycudann"].free_temporary_memory = lambda: None
    sys.modules["omegaconf"].OmegaConf.to_container = staticmethod(lambda c, resolve=True: {k: v for k, v in c.items()})
    torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    _stub("cv2")
    import utils.misc as ref_misc
    ref_misc.get_rank = lambda: "cpu"
    models_mod = sys.modules["models"]
    _stub("lib"); sys.modules["lib"].__path__ = [os.path.join(REF, "lib")]
    from models import network_utils as ref_net
    ref_net.get_rank = lambda: "cpu"
    import models.base as ref_base
    ref_base.get_rank = lambda: "cpu"
    _pkg("models.pbr", os.path.join(REF, "models", "pbr"))
    from models.rf import geometry as ref_geo, radiance as ref_rad
    ref_geo.get_rank = lambda: "cpu"
    from models.pbr import material as ref_mat

    class Cfg(dict):
        __getattr__ = dict.__getitem__
        def copy(self):
            return Cfg(self)
    def cfg(d):
        return Cfg({k: cfg(v) if isinstance(v, dict) else v for k, v in d.items()})
    def load(path, **subst):
        txt = open(os.path.join(REF, "configs", path)).read()
        for k, v in subst.items():
            txt = txt.replace(k, str(v))
        return cfg(yaml.safe_load(txt))
    gcfg = load("geometry/progressive_hash_grid.yaml", **{"${model.radius}": 1.0})
    gcfg["isosurface"] = None
    rcfg = load("radiance/progressive_hash_grid.yaml", **{"${add:${model.geometry.feature_dim}, 3}": 16})
    mcfg = load("material/shallow_mlp.yaml", **{"${add:${model.geometry.feature_dim}, 35}": 48})
    geo, rad, mat = ref_geo.VolumeSDF(gcfg), ref_rad.VolumeRefDirRadiance(rcfg), ref_mat.VolumeMaterial(mcfg)
    sd = random_state_dict(0)
    sub = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    print("geometry:", geo.load_state_dict(sub("geometry."), strict=False))
    print("radiance:", rad.load_state_dict(sub("radiance."), strict=False))
    msd = sub("material.")
    print("material:", mat.load_state_dict(msd, strict=False))
    for i in range(3):
        mat.network.weights_per_layer[i].data.copy_(msd[f"network.layers.{i}.weight"])
        mat.network.biases_per_layer[i].data.copy_(msd[f"network.layers.{i}.bias"])
        mat.network.lipshitz_bound_per_layer[i].data.copy_(msd[f"network.lipshitz_bound_per_layer.{i}"].reshape(-1))
    for m in (geo, rad, mat):
        m.train(False)
    # the fully warmed-up test-time state: all hash levels and SH bands on (update_step(250, 25000))
    geo.encoding.encoding.update_step(250, 25000)
    rad.xyz_encoding.encoding.update_step(250, 25000)
    rad.sh_mask[:] = 1.0
    from intrinsicavatar_b200.snarf import SnarfSetup
    bbox = torch.from_numpy(SnarfSetup().bbox)
    geo.prepare_bbox(bbox)
    rad.center, rad.scale = geo.center, geo.scale          # models/intrinsic_avatar.py prepare: radiance shares the bbox
    torch.manual_seed(5)
    n = 256
    pts = geo.center + (torch.rand(n, 3) - 0.5) * geo.scale * 0.7
    sdf, grad, feat = geo(pts, with_grad=True, with_feature=True)
    view = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
    nrm = torch.nn.functional.normalize(grad + 0.05 * torch.randn(n, 3), dim=-1)
    with torch.no_grad():
        rgb, rgb_feature = rad(pts, feat, view, nrm)
        materials = mat(torch.cat([rgb_feature, feat], dim=-1))      # material_feature == "hybrid"
    g = dict(fields_bbox=bbox, fields_points=pts, fields_sdf=sdf, fields_grad=grad, fields_feature=feat, fields_view=view,
             fields_normal=nrm, fields_rgb=rgb, fields_xyz_embd=rgb_feature, fields_materials=materials)
    out = os.path.join(ROOT, "tests", "golden", "reference_vectors_fields.npz")
    np.savez_compressed(out, **{k: v.detach().cpu().numpy().astype(np.float32) for k, v in g.items()})
    print("wrote", out, {k: tuple(v.shape) for k, v in g.items()})


E2E_CASES = [
    # name, frame (None = neutral pose), image side, spp, render_mode, global_illumination, add_emitter
    ("light_neutral", None, 20, 4, "light", False, False),
    ("light_gi_posed", 0, 20, 8, "light", True, False),
    ("light_emitter", 0, 16, 4, "light", False, True),
    ("mats", 0, 16, 8, "mats", False, False),
    ("mis_gi", 0, 16, 4, "mis", True, False),
    ("uniform_light", 0, 8, 512, "uniform_light", False, False),
    # externally set test-time attributes (systems/base.py:112-119, systems/intrinsic_avatar.py:601-617): flags after add_emitter
    ("albedo_only", 0, 16, 4, "light", False, False, {"albedo_only": True}),
    ("black_bg_albedo_ratio", 0, 16, 4, "light", False, False, {"background": (0.0, 0.0, 0.0), "albedo_align_ratio": (1.2, 0.9, 0.8)}),
]
E2E_KEYS = ("comp_rgb", "comp_normal", "opacity", "depth", "comp_albedo", "comp_roughness", "comp_metallic", "comp_rgb_phys",
            "comp_demod_phys", "comp_rgb_full", "comp_rgb_phys_full", "comp_albedo_full", "comp_roughness_full")


def main_e2e():
    """The reference's OWN IntrinsicAvatarModel.forward_ (+ _compute_occupancy_grid) executed on CPU through
    scripts/ref_harness.py (third-party / CUDA ops replaced by the oracle's restatements of exactly those ops; everything
    else -- control flow and glue of forward_, compute_indirect_radiance, pbr_*_forward, volrend, sample_volume_interaction,
    the deformer classes, the field modules, torch_pbr -- is the reference's code).  Six small frames covering the four
    render modes, global illumination and add_emitter."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import ref_harness as H
    from conftest import Scene
    H.install()
    sc = Scene()
    env = sc.syn.load_envmap()
    g = {}
    grids = {}
    for name, frame, side, spp, mode, gi, emit, *extra in E2E_CASES:
        extra = extra[0] if extra else {}
        fr = sc.frame(frame)
        tabs = sc.syn.random_tables(spp, 32, seed=0)
        if frame not in grids:
            # the reference's own test-grid construction (models/intrinsic_avatar.py:307-362) at resolution 32 with the
            # jitter table product and oracle use in place of torch.rand_like
            m0 = H.build_model(sc, fr, spp)
            import models.intrinsic_avatar as ref_ia
            coords = ref_ia._meshgrid3d(torch.tensor([32, 32, 32])).reshape(-1, 3)
            orig = torch.rand_like
            torch.rand_like = lambda t, **k: torch.from_numpy(tabs["jitter"]).reshape(t.shape)
            try:
                _, binaries, aabb = m0._compute_occupancy_grid(coords, resolution=32)
            finally:
                torch.rand_like = orig
            assert torch.allclose(aabb, torch.as_tensor(fr["deformed_bbox"]), atol=1e-5)
            grids[frame] = binaries[0]
            g[f"grid_{'neutral' if frame is None else frame}"] = np.packbits(binaries[0].numpy().reshape(-1))
            print("grid", frame, int(binaries.sum()), "occupied cells")
        m = H.build_model(sc, fr, spp, gi=gi, render_mode=mode, add_emitter=emit, binaries=grids[frame], env=env,
                          u1=tabs["u1"], u2=tabs["u2"])
        if "background" in extra:
            m.background_color = torch.tensor(extra["background"])
        if "albedo_align_ratio" in extra:
            m.albedo_align_ratio = torch.tensor(extra["albedo_align_ratio"])
        m.albedo_only = bool(extra.get("albedo_only", False))
        rays = torch.from_numpy(sc.syn.make_rays(side, side, fr["transl"]))
        out = H.forward(m, rays, seed=0)
        if name in ("light_neutral", "uniform_light"):
            # output contract of the public forward(): key -> (shape, dtype) for n rays (models/intrinsic_avatar.py:1653-1666)
            pub = H.forward(m, rays, seed=0, public=True)
            g[f"{name}/contract"] = np.array(sorted(f"{k}|{tuple(v.shape)}|{str(v.dtype).replace('torch.', '')}|{v.device.type}"
                                                    for k, v in pub.items()))
        for k in E2E_KEYS + (("visibility",) if mode == "uniform_light" else ()):
            g[f"{name}/{k}"] = out[k].detach().numpy().astype(np.float32)
        print(name, "hit rays", int((out["opacity"] > 0.5).sum()), "of", rays.shape[0],
              "mean rgb_phys over hits", float(out["comp_rgb_phys"][out["opacity"][:, 0] > 0.5].mean()))
    out_path = os.path.join(ROOT, "tests", "golden", "reference_vectors_e2e.npz")
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, len(g), "arrays")


def main_e2e_hi():
    """The reference's own forward_ (scripts/ref_harness.py) at the sample counts the bench runs: 64 / 256 / 1024 spp,
    the last one with global illumination, AIST frame 0, nonzero ray-index offsets, the real city.hdr.
    -> tests/golden/reference_vectors_e2e_hi.npz"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import time
    import e2e_cases as E2E
    import ref_harness as H
    from conftest import Scene
    H.install()
    sc = Scene()
    env = sc.syn.load_envmap_full()
    g = {}
    grid = None
    only = sys.argv[2:]
    out_path = os.path.join(ROOT, "tests", "golden", "reference_vectors_e2e_hi.npz")
    if only and os.path.exists(out_path):
        z = np.load(out_path)
        g = {k: z[k] for k in z.files}
    for name, frame, side, spp, mode, gi, offset in E2E.HI_CASES:
        if only and name not in only:
            continue
        t0 = time.time()
        fr = sc.frame(frame)
        tabs = sc.syn.random_tables(spp, 32, seed=0)
        if grid is None:
            m0 = H.build_model(sc, fr, spp)
            import models.intrinsic_avatar as ref_ia
            coords = ref_ia._meshgrid3d(torch.tensor([32, 32, 32])).reshape(-1, 3)
            orig = torch.rand_like
            torch.rand_like = lambda t, **k: torch.from_numpy(tabs["jitter"]).reshape(t.shape)
            try:
                _, binaries, aabb = m0._compute_occupancy_grid(coords, resolution=32)
            finally:
                torch.rand_like = orig
            grid = binaries[0]
            g["grid_0"] = np.packbits(grid.numpy().reshape(-1))
        m = H.build_model(sc, fr, spp, gi=gi, render_mode=mode, binaries=grid, env=env, u1=tabs["u1"], u2=tabs["u2"])
        rays = E2E.hi_rays(sc.syn, fr["transl"], side)
        out = H.forward(m, rays, seed=0, ray_offset=offset)
        for k in E2E_KEYS:
            g[f"{name}/{k}"] = out[k].detach().numpy().astype(np.float32)
        print(name, "hit rays", int((out["opacity"] > 0.5).sum()), "of", rays.shape[0], "mean rgb_phys over hits",
              float(out["comp_rgb_phys"][out["opacity"][:, 0] > 0.5].mean()), "%.0f s" % (time.time() - t0), flush=True)
        np.savez_compressed(out_path, **g)
    print("wrote", out_path, len(g), "arrays")


def main_config():
    """The reference's ``config.model`` node for the relighting test run (configs/config.yaml with its defaults list, dataset
    animation/male-3-casual, light envlight_tensor as README.md:84-95 selects), Hydra interpolations resolved by hand ->
    tests/golden/reference_model_config.json.  tests/test_host_logic.py and tests/test_gpu_render.py build
    IntrinsicAvatarModel from it through a models.register / models.make registry like the reference's."""
    import json
    import yaml
    C = os.path.join(REF, "configs")
    load = lambda rel: yaml.safe_load(open(os.path.join(C, rel)))
    root = load("config.yaml")
    choice = {"dataset": "animation/male-3-casual", "light": "envlight_tensor"}
    groups = {}
    for d in root["defaults"]:
        if isinstance(d, dict):
            for g, name in d.items():
                if not g.startswith("override"):
                    groups[g] = choice.get(g, name)
    nodes = {g: load(f"{g}/{name}.yaml") for g, name in groups.items()}
    model = root["model"]
    subst = {"${dataset.scene_aabb}": nodes["dataset"]["scene_aabb"], "${dataset.gender}": nodes["dataset"]["gender"],
             "${trainer.precision}": 32, "${add:${model.geometry.feature_dim}, 3}": 16,
             "${add:${model.geometry.feature_dim}, 35}": 48}

    def resolve(v):
        if isinstance(v, dict):
            return {k: resolve(x) for k, x in v.items()}
        if isinstance(v, list):
            return [resolve(x) for x in v]
        if isinstance(v, str):
            if v in subst:
                return subst[v]
            if v.startswith("${") and v[2:-1] in nodes:
                return resolve(nodes[v[2:-1]])
        return v
    out = resolve(model)
    # what the README's relight command overrides (README.md:84-95)
    out.update({"render_mode": "light", "global_illumination": False, "samples_per_pixel": 1024, "resample_light": False,
                "add_emitter": True})
    path = os.path.join(ROOT, "tests", "golden", "reference_model_config.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, sorted(out))


def main_e2e_switch():
    """The reference's own forward_ with the non-default switches of config.model: zero_crossing_search = false
    (ray_resampling_fine), secondary_importance_sample = false (coarse samples rendered), material_feature = geometry /
    radiance.  -> tests/golden/reference_vectors_e2e_switch.npz"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import e2e_cases as E2E
    import ref_harness as H
    from conftest import Scene
    H.install()
    sc = Scene()
    env = sc.syn.load_envmap()
    base = E2E.load()                      # the frame-0 occupancy grid of the main golden (the reference's own build)
    grid = E2E.grid(base, 0)
    g = {}
    for name, frame, side, spp, gi, opts in E2E.SWITCH_CASES:
        fr = sc.frame(frame)
        tabs = sc.syn.random_tables(spp, E2E.GRID_RES, seed=0)
        m = H.build_model(sc, fr, spp, gi=gi, render_mode="light", binaries=grid, env=env, u1=tabs["u1"], u2=tabs["u2"], **opts)
        rays = torch.from_numpy(sc.syn.make_rays(side, side, fr["transl"]))
        out = H.forward(m, rays, seed=0)
        for k in E2E_KEYS:
            g[f"{name}/{k}"] = out[k].detach().numpy().astype(np.float32)
        print(name, "hit rays", int((out["opacity"] > 0.5).sum()), "of", rays.shape[0], "mean rgb_phys over hits",
              float(out["comp_rgb_phys"][out["opacity"][:, 0] > 0.5].mean()), flush=True)
    out_path = os.path.join(ROOT, "tests", "golden", "reference_vectors_e2e_switch.npz")
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, len(g), "arrays")


def main_snarf():
    """The reference's own SNARFDeformer.initialize + prepare_deformer (models/deformers/snarf_deformer.py:46-126) driven
    by a body model that calls the reference's lbs() on the random SMPL-shaped arrays of reference_vectors_smpl.npz (the
    licensed .pkl is absent) -- pins SnarfSetup.__init__ / .frame (tfs, w2s, root-frame vertices, bboxes)."""
    install_stubs()
    sys.path.insert(0, ROOT)
    import torch.utils.cpp_extension as cpp
    cpp.load = lambda *a, **k: types.SimpleNamespace()

    def knn_points(x, y, K=1):
        d2 = torch.cdist(x.double(), y.double()) ** 2
        val, idx = torch.topk(d2, K, dim=-1, largest=False)
        return val.to(x.dtype), idx, None
    _pkg("lib", os.path.join(REF, "lib"))
    _pkg("lib.pytorch3d", os.path.join(REF, "lib", "pytorch3d"))
    _stub("lib.pytorch3d.ops", knn_points=knn_points)
    sys.modules["lib.pytorch3d"].ops = sys.modules["lib.pytorch3d.ops"]
    _pkg("models.deformers.fast_snarf", os.path.join(REF, "models", "deformers", "fast_snarf"))
    _stub("torchgeometry")
    _stub("torchgeometry.core")
    _stub("torchgeometry.core.conversions",
          angle_axis_to_rotation_matrix=lambda aa: torch.eye(4)[None].repeat(aa.shape[0], 1, 1))
    sys.modules["torchgeometry.core"].conversions = sys.modules["torchgeometry.core.conversions"]
    sys.modules["models.deformers.smplx"].SMPL = object      # (the .pkl-backed class; replaced by FakeSMPL below)
    from models.deformers import snarf_deformer as ref_sd
    from models.deformers.smplx import lbs as ref_lbs
    from intrinsicavatar_b200.body import PARENTS
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_smpl.npz"))
    T = lambda k: torch.from_numpy(z[k]).double()

    class FakeSMPL(torch.nn.Module):
        """SMPL.forward (body_models.py:288-370) with the arrays above instead of the .pkl."""
        def __init__(self):
            super().__init__()
            self.register_buffer("lbs_weights", T("smpl_weights").float())
            self.register_buffer("faces_tensor", torch.zeros(1, 3, dtype=torch.long))
            self.dummy = torch.nn.Parameter(torch.zeros(1))

        def forward(self, betas=None, body_pose=None, global_orient=None, transl=None):
            go = torch.zeros(1, 3) if global_orient is None else global_orient
            pose = torch.cat([go, body_pose], 1).double()
            v, j, A, Tm, so, po = ref_lbs.lbs(betas.double(), pose, T("smpl_v_template")[None], T("smpl_shapedirs"),
                                             T("smpl_posedirs"), T("smpl_J_regressor"),
                                             torch.from_numpy(np.asarray(PARENTS, np.int64)), T("smpl_weights"))
            if transl is not None:
                v, j = v + transl.double()[:, None], j + transl.double()[:, None]
                A = A.clone()
                A[..., :3, 3] += transl.double()[:, None]
            return types.SimpleNamespace(vertices=v.float(), joints=j.float(), A=A.float())   # SMPL runs in float32

    sd = ref_sd.SNARFDeformer.__new__(ref_sd.SNARFDeformer)
    sd.body_model = FakeSMPL()
    fd = ref_sd.ForwardDeformer.__new__(ref_sd.ForwardDeformer)
    torch.nn.Module.__init__(fd)
    fd.global_scale = 1.2
    seen = {}
    fd.precompute = lambda tfs: seen.__setitem__("tfs", tfs)
    sd.deformer = fd
    sd.initialized = False
    sd.opt = types.SimpleNamespace(cano_pose="a_pose", resolution=32, optimize_betas=False)
    pose = T("smpl_pose").float()
    zero = lambda *s_: torch.zeros(*s_)
    params = {"betas": T("smpl_betas").float(), "body_pose": pose[:, 3:], "global_orient": pose[:, :3],
              "transl": T("smpl_transl").float(), "pose_correction": zero(1, 69), "global_orient_correction": zero(1, 3),
              "transl_correction": zero(1, 3)}
    sd.prepare_deformer(params)
    g = dict(snarf_tfs=sd.tfs[0], snarf_w2s=sd.w2s[0], snarf_vertices=sd.vertices[0], snarf_cano_bbox=sd.bbox,
             snarf_tfs_inv_t=sd.tfs_inv_t[0], snarf_lbs_voxel=fd.lbs_voxel_final[0], snarf_offset_kernel=fd.offset_kernel.reshape(3),
             snarf_scale_kernel=fd.scale_kernel.reshape(3),
             snarf_deformed_bbox=ref_sd.get_bbox_from_smpl(sd.vertices.float()))
    out = os.path.join(ROOT, "tests", "golden", "reference_vectors_snarf.npz")
    np.savez_compressed(out, **{k: v.detach().cpu().numpy().astype(np.float32) for k, v in g.items()})
    print("wrote", out, {k: tuple(v.shape) for k, v in g.items()})


def main_saver():
    """The image grid of test_step as the reference's own SaverMixin builds it (utils/mixins.py:43-53, 87-101, 116-144):
    random columns of the kinds test_step uses (systems/intrinsic_avatar.py:721-841) -> the uint8 grid in the channel order
    of the written file (get_image_grid_ returns BGR for cv2.imwrite) -> tests/golden/reference_vectors_saver.npz."""
    install_stubs()
    _stub("matplotlib", cm=types.SimpleNamespace())
    _stub("matplotlib.cm")
    _stub("matplotlib.colors", LinearSegmentedColormap=object)
    _stub("utils.obj", write_obj=None)
    from utils.mixins import SaverMixin
    g = torch.Generator().manual_seed(11)
    H, W = 12, 16
    rgb = torch.rand(H, W, 3, generator=g) * 1.5 - 0.25            # out of range on both sides
    chw = torch.rand(3, H, W, generator=g)
    normal = torch.rand(H, W, 3, generator=g) * 2.4 - 1.2
    rough = torch.rand(H, W, generator=g) * 1.2 - 0.1
    depth = torch.rand(H, W, generator=g) * 4 + 1
    depth[0, 0] = float("nan")
    depth[1, 1] = float("inf")
    two = torch.rand(H, W, 2, generator=g)                          # fewer than three channels: zero-padded
    imgs = [
        {"type": "rgb", "img": rgb, "kwargs": {"data_format": "HWC"}},
        {"type": "rgb", "img": chw, "kwargs": {}},
        {"type": "grayscale", "img": rough, "kwargs": {"data_range": (0, 1), "cmap": None}},
        {"type": "grayscale", "img": depth.clone().nan_to_num(posinf=5.0), "kwargs": {}},
        {"type": "rgb", "img": normal, "kwargs": {"data_format": "HWC", "data_range": (-1, 1)}},
        {"type": "rgb", "img": two, "kwargs": {"data_format": "HWC"}},
        {"type": "grayscale", "img": depth, "kwargs": {"data_range": (0, 6), "cmap": "jet"}},
    ]
    grid, cols = SaverMixin().get_image_grid_(imgs)
    out = os.path.join(ROOT, "tests", "golden", "reference_vectors_saver.npz")
    np.savez_compressed(out, rgb=rgb.numpy(), chw=chw.numpy(), normal=normal.numpy(), rough=rough.numpy(), depth=depth.numpy(),
                        two=two.numpy(), grid_file_rgb=np.ascontiguousarray(grid[..., ::-1]))
    print("wrote", out, grid.shape, grid.dtype)


def main_occ_ema():
    """The training-time occupancy update as the reference's own TemporalOccGridEstimator._update runs it
    (models/occ_grid/temporal_occ_grid.py:369-411): two successive updates of a 16^3 level with a synthetic occupancy
    function (the grid logic is what is pinned: one jittered point per cell, EMA with max, max-pool, threshold
    min(mean, occ_thre), largest connected component) -> tests/golden/reference_vectors_occ_ema.npz."""
    install_stubs()
    sys.modules["nerfacc"].traverse_grids = None
    sys.modules["nerfacc"].render_visibility_from_alpha = None
    sys.modules["nerfacc"].render_visibility_from_density = None
    _pkg("models.occ_grid", os.path.join(REF, "models", "occ_grid"))
    from models.occ_grid.temporal_occ_grid import TemporalOccGridEstimator
    R = 16
    aabb = torch.tensor([[-0.9, -1.1, -0.4, 0.8, 0.7, 0.5]])
    est = TemporalOccGridEstimator(roi_aabb=aabb, resolution=R, levels=1)

    def occ_fn(x):     # two blobs (the smaller one must be dropped by the connected-component step) with smooth fall-off
        a = torch.exp(-((x - torch.tensor([0.0, -0.2, 0.0])) ** 2).sum(-1) / 0.08)
        b = 0.6 * torch.exp(-((x - torch.tensor([0.55, 0.45, 0.3])) ** 2).sum(-1) / 0.004)
        return (a + b).clamp(max=1.0)[:, None] * 0.05
    out = {"aabb": aabb[0].numpy(), "res": np.int32(R)}
    for k, (seed, decay, thre) in enumerate(((3, 0.8, 0.001), (4, 0.8, 0.001), (5, 0.95, 0.01))):
        torch.manual_seed(seed)
        jitter = torch.rand(R ** 3, 3)
        x = (est.grid_coords + jitter) / est.resolution
        x = est.aabbs[0, :3] + x * (est.aabbs[0, 3:] - est.aabbs[0, :3])
        out[f"occ_in_{k}"] = occ_fn(x)[:, 0].numpy()
        out[f"state_in_{k}"] = est.occs.clone().numpy()
        torch.manual_seed(seed)                  # _update draws torch.rand_like(grid_coords): the same numbers
        est._update(step=0, t_idx=0, occ_eval_fn=occ_fn, occ_thre=thre, ema_decay=decay)
        out[f"jitter_{k}"] = jitter.numpy()
        out[f"state_out_{k}"] = est.occs.clone().numpy()
        out[f"binaries_{k}"] = est.binaries[0].clone().numpy()
        out[f"params_{k}"] = np.array([decay, thre], np.float32)
        print(k, "occupied", int(est.binaries.sum()))
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors_occ_ema.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "occ_ema":
        main_occ_ema()
    elif len(sys.argv) > 1 and sys.argv[1] == "saver":
        main_saver()
    elif len(sys.argv) > 1 and sys.argv[1] == "e2e":
        main_e2e()
    elif len(sys.argv) > 1 and sys.argv[1] == "e2e_switch":
        main_e2e_switch()
    elif len(sys.argv) > 1 and sys.argv[1] == "config":
        main_config()
    elif len(sys.argv) > 1 and sys.argv[1] == "e2e_hi":
        main_e2e_hi()
    elif len(sys.argv) > 1 and sys.argv[1] == "fields":
        main_fields()
    elif len(sys.argv) > 1 and sys.argv[1] == "snarf":
        main_snarf()
    elif len(sys.argv) > 1 and sys.argv[1] == "voxel":
        main_voxel()
    elif len(sys.argv) > 1 and sys.argv[1] == "deform_train":
        main_deform_train()
    elif len(sys.argv) > 1 and sys.argv[1] == "smpl":
        main_smpl()
    elif len(sys.argv) > 1 and sys.argv[1] == "bsdf":
        main_bsdf()
    else:
        main()
