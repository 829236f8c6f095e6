"""Network parameters of the render path: reference ``state_dict`` layout, random init, folding.

The reference's checkpoint keys (SURVEY.md section 5 "checkpoint / resume") are kept so a
Lightning checkpoint can be ingested with ``strict=False`` (launch.py:110-124):

  geometry.encoding.encoding.encoding.params            hash grid, [6 299 960 * 2] f32
  geometry.network.layers.{0,2}.{weight_g,weight_v,bias} weight-normed VanillaMLP 35->64->13
  radiance.xyz_encoding.encoding.encoding.params        second hash grid
  radiance.network.layers.{0,2,4}.{weight,bias}         VanillaMLP 67->64->64->3
  material.network.layers.{0,1,2}.{weight,bias}         LipshitzMLP 48->64->64->5
  material.network.lipshitz_bound_per_layer.{0,1,2}
  density.beta

``fold`` turns them into the plain dense matrices the kernels consume:
  * weight norm  W = g * v / ||v||_row           (torch.nn.utils.weight_norm, dim=0)
  * Lipschitz    W = W * min(1, softplus(c) / sum_j |W_ij|)   (models/network_utils.py:391-397)
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

N_LEVELS = 16
N_FEAT = 2
LOG2_T = 19
BASE_RES = 16
PER_LEVEL_SCALE = 1.447269237440378


def hashgrid_layout(n_levels=N_LEVELS, log2_t=LOG2_T, base=BASE_RES, per_level_scale=PER_LEVEL_SCALE):
    """Per-level (scale, resolution, size, offset) of the tiny-cuda-nn HashGrid.

    float32 arithmetic as in tcnn's ``grid_scale`` / ``grid_resolution``; offsets in entries.
    """
    log2s = np.log2(np.float32(per_level_scale)).astype(np.float32)
    scales, ress, sizes, offsets = [], [], [], []
    off = 0
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2s) * np.float32(base) - np.float32(1.0))
        res = int(np.ceil(scale)) + 1
        size = min(res ** 3, 2 ** 62)
        size = (size + 7) // 8 * 8
        size = min(size, 1 << log2_t)
        scales.append(float(scale))
        ress.append(res)
        sizes.append(size)
        offsets.append(off)
        off += size
    return {
        "scale": np.asarray(scales, np.float32),
        "res": np.asarray(ress, np.int32),
        "size": np.asarray(sizes, np.int32),
        "offset": np.asarray(offsets, np.int32),
        "total": off,
    }


def random_state_dict(seed: int = 0, beta: float = 0.01, geo_hash_amp: float = 0.2,
                      geo_feat_std: float = 0.05) -> dict:
    """Synthetic weights: sphere-init geometry MLP, Kaiming radiance MLP,
    default-init Lipschitz material MLP.  The geometry hash grid is U(-a_l, a_l) with
    a_l = geo_hash_amp * (16 / res_l)^2: detail whose gradient contribution decays with the level,
    like a trained, eikonal-regularised SDF (a bumpy sphere with |grad| ~ 0.8 +- 0.17) instead of
    white noise whose normals would flip on every 0.5 mm cell face, and the first geometry layer gets
    small weights on the hash features so the grid actually shapes the surface (the reference's
    sphere init leaves them at 0).  Keys follow the reference state_dict."""
    g = torch.Generator().manual_seed(seed)
    lay = hashgrid_layout()
    geo = torch.empty(lay["total"], N_FEAT)
    for l in range(N_LEVELS):
        o, n = int(lay["offset"][l]), int(lay["size"][l])
        geo[o:o + n] = (torch.rand(n, N_FEAT, generator=g) * 2 - 1) * geo_hash_amp * (16.0 / float(lay["res"][l])) ** 2
    sd = {}
    sd["geometry.encoding.encoding.encoding.params"] = geo.reshape(-1)
    sd["radiance.xyz_encoding.encoding.encoding.params"] = (
        (torch.rand(lay["total"] * N_FEAT, generator=g) * 2 - 1) * 0.1)

    # geometry VanillaMLP 35 -> 64 -> 13, sphere init (network_utils.py:219-238), weight-normed
    d_in, n, d_out = 35, 64, 13
    w0 = torch.zeros(n, d_in)
    w0[:, :3] = torch.randn(n, 3, generator=g) * (math.sqrt(2) / math.sqrt(n))
    w0[:, 3:] = torch.randn(n, d_in - 3, generator=g) * geo_feat_std  # synthetic detail (0 in the reference init)
    w2 = torch.randn(d_out, n, generator=g) * 0.0001 + math.sqrt(math.pi) / math.sqrt(n)
    # rows 1..12 are features: give them variety so radiance/material inputs are not constant
    w2[1:] = torch.randn(d_out - 1, n, generator=g) * (1.0 / math.sqrt(n))
    for i, (w, b) in enumerate([(w0, torch.zeros(n)), (w2, torch.full((d_out,), -0.5))]):
        k = f"geometry.network.layers.{2 * i}"
        sd[k + ".weight_v"] = w
        sd[k + ".weight_g"] = w.norm(dim=1, keepdim=True)
        sd[k + ".bias"] = b
    sd["geometry.network.layers.2.bias"][1:] = 0.0

    def kaiming_uniform(o, i):
        bound = math.sqrt(6.0 / i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * bound

    dims = [67, 64, 64, 3]
    for li in range(3):
        sd[f"radiance.network.layers.{2 * li}.weight"] = kaiming_uniform(dims[li + 1], dims[li])
        sd[f"radiance.network.layers.{2 * li}.bias"] = torch.zeros(dims[li + 1])

    def linear_default(o, i):
        bound = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * bound, (torch.rand(o, generator=g) * 2 - 1) * bound

    dims = [48, 64, 64, 5]
    for li in range(3):
        w, b = linear_default(dims[li + 1], dims[li])
        sd[f"material.network.layers.{li}.weight"] = w
        sd[f"material.network.layers.{li}.bias"] = b
        sd[f"material.network.lipshitz_bound_per_layer.{li}"] = (w.abs().sum(1).max() * 2).reshape(1)
    sd["density.beta"] = torch.tensor(beta)
    return sd


MATERIAL_IN = {"hybrid": 48, "geometry": 13, "radiance": 35}     # config.model.material_feature -> material MLP input width


def material_state_dict_for(sd: dict, material_feature: str) -> dict:
    """The state dict of a model whose material net sees only the geometry feature / only the radiance encoding
    (models/intrinsic_avatar.py:1102-1113), cut out of a hybrid one: the first material layer keeps the columns of that
    input (hybrid input = cat[xyz_embd 35, feature 13])."""
    out = dict(sd)
    w = sd["material.network.layers.0.weight"]
    assert w.shape[1] == 48
    if material_feature == "geometry":
        out["material.network.layers.0.weight"] = w[:, 35:].clone()
    elif material_feature == "radiance":
        out["material.network.layers.0.weight"] = w[:, :35].clone()
    return out


def fold(sd: dict, material_feature: str = "hybrid", keep_graph: bool = False) -> dict:
    """Reference state_dict -> dense fp32 matrices (row-major [out,in]) + scalars for the kernels.  The kernels evaluate the
    material net on the hybrid input cat[xyz_embd 35, feature 13]; a model with material_feature = geometry | radiance has a
    13- / 35-wide first layer, which is the hybrid layer with zero columns for the input it does not see.
    ``keep_graph``: plain tensor code throughout, so with parameters that require grad the result stays attached to them
    (``beta`` is then a tensor): the training seam differentiates weight norm, Lipschitz bound and beta through this function."""
    if not keep_graph:
        sd = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in sd.items()}
    out = {}
    out["geo_hash"] = sd["geometry.encoding.encoding.encoding.params"].float().contiguous()
    out["rad_hash"] = sd["radiance.xyz_encoding.encoding.encoding.params"].float().contiguous()
    for i, name in enumerate(["geo_w1", "geo_w2"]):
        k = f"geometry.network.layers.{2 * i}"
        v, gg = sd[k + ".weight_v"].float(), sd[k + ".weight_g"].float()
        out[name] = (gg * v / v.norm(dim=1, keepdim=True)).contiguous()
        out[name.replace("w", "b")] = sd[k + ".bias"].float().contiguous()
    for li in range(3):
        out[f"rad_w{li + 1}"] = sd[f"radiance.network.layers.{2 * li}.weight"].float().contiguous()
        out[f"rad_b{li + 1}"] = sd[f"radiance.network.layers.{2 * li}.bias"].float().contiguous()
    for li in range(3):
        w = sd[f"material.network.layers.{li}.weight"].float()
        c = F.softplus(sd[f"material.network.lipshitz_bound_per_layer.{li}"].float())
        s = torch.clamp(c / w.abs().sum(1), max=1.0)
        wf = w * s[:, None]
        if li == 0 and wf.shape[1] != 48:
            assert wf.shape[1] == MATERIAL_IN[material_feature], (tuple(wf.shape), material_feature)
            pad = torch.zeros(wf.shape[0], 48, device=wf.device)
            if material_feature == "geometry":
                pad[:, 35:] = wf
            else:
                pad[:, :35] = wf
            wf = pad
        out[f"mat_w{li + 1}"] = wf.contiguous()
        out[f"mat_b{li + 1}"] = sd[f"material.network.layers.{li}.bias"].float().contiguous()
    beta = sd["density.beta"].float().abs() + 1e-4        # LearnedLaplaceDensity.get_beta, density.py:32-34
    out["beta"] = beta if keep_graph else float(beta.detach())
    return out


RENDER_PATH_PREFIXES = ("geometry.", "radiance.", "material.", "density.")


def load_lightning_checkpoint(path: str, map_location="cpu", material_feature: str = "hybrid") -> dict:
    """Reference-keyed state dict of the render path out of a Lightning checkpoint (launch.py:110-124 loads
    ``ckpt['state_dict']`` into the system with ``strict=False``; the model's parameters live under ``model.``).
    Keys outside the render path (pose correction, non-rigid, occupancy grids, emitter, loss state) are dropped;
    the shapes of what remains are checked against the network layout the kernels are built for."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt.get("state_dict", ckpt)
    out = {}
    for k, v in sd.items():
        if k.startswith("model."):
            k = k[len("model."):]
        if k.startswith(RENDER_PATH_PREFIXES) and torch.is_tensor(v):
            out[k] = v.detach().float()
    expect = {k: tuple(v.shape) for k, v in random_state_dict_shapes(material_feature).items()}
    missing = sorted(set(expect) - set(out))
    if missing:
        raise KeyError(f"checkpoint {path} lacks render-path parameters: {missing[:4]}{'...' if len(missing) > 4 else ''}")
    for k, shp in expect.items():
        if tuple(out[k].shape) != shp and out[k].numel() != int(np.prod(shp)):
            raise ValueError(f"{k}: checkpoint shape {tuple(out[k].shape)} does not match the built network {shp}")
        out[k] = out[k].reshape(shp)
    return out


def random_state_dict_shapes(material_feature: str = "hybrid") -> dict:
    """Shapes of every render-path parameter (meta tensors, no allocation)."""
    lay = hashgrid_layout()
    n = lay["total"] * N_FEAT
    m = lambda *s: torch.empty(*s, device="meta")
    sd = {"geometry.encoding.encoding.encoding.params": m(n), "radiance.xyz_encoding.encoding.encoding.params": m(n),
          "density.beta": m(())}
    for i, (o, i_) in enumerate([(64, 35), (13, 64)]):
        k = f"geometry.network.layers.{2 * i}"
        sd[k + ".weight_v"], sd[k + ".weight_g"], sd[k + ".bias"] = m(o, i_), m(o, 1), m(o)
    for li, (o, i_) in enumerate([(64, 67), (64, 64), (3, 64)]):
        sd[f"radiance.network.layers.{2 * li}.weight"], sd[f"radiance.network.layers.{2 * li}.bias"] = m(o, i_), m(o)
    for li, (o, i_) in enumerate([(64, MATERIAL_IN[material_feature]), (64, 64), (5, 64)]):
        sd[f"material.network.layers.{li}.weight"], sd[f"material.network.layers.{li}.bias"] = m(o, i_), m(o)
        sd[f"material.network.lipshitz_bound_per_layer.{li}"] = m(1)
    return sd
