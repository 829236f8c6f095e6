"""ORACLE (test infrastructure): environment light, BRDF, sRGB -- CPU PyTorch fp32.

Restates (reference file:line):
  * ``EnvironmentLightTensor.update_pdf / sample / pdf / eval``  lib/torch_pbr/light.py:259-446
    with ``xyz2lonlat_mode = null`` conventions  lib/torch_pbr/utils/light_utils.py:6-63,
    ``pixel_grid``  lib/torch_pbr/utils/nvdiffrecmc_util.py:61-65.  (The reference class cannot be
    constructed on CPU -- ``device="cuda"`` is hard-coded, light.py:230,239 -- hence a restatement;
    ``sample`` takes its uniforms as explicit inputs instead of ``torch.rand``.)
  * ``MultiLobe.eval`` = Lambertian + GGX  lib/torch_pbr/bxdf.py:111-146, 217-265, 321-330 and
    warp_utils helpers  lib/torch_pbr/utils/warp_utils.py:62-101, 693-702, 730-747, 782-794.
    Pinned: tests/golden/bxdf_multilobe.npz is produced by the reference's own module
    (scripts/make_golden_torch_pbr.py).
  * ``rgb_to_srgb``  lib/torch_pbr/utils/nvdiffrecmc_util.py:94-102.
  * light-index shuffle (models/intrinsic_avatar.py:1355-1378): the reference draws
    ``argsort(rand(n_rays, spp))`` from torch's global RNG on the CPU; product and oracle share the
    stateless keyed permutation ``kensler_permute`` instead (statistically equivalent, exactly
    reproducible).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------ env light ----
class EnvLight:
    def __init__(self, base: torch.Tensor):
        self.base = base.float()                      # [H,W,3]
        H, W = base.shape[:2]
        self.H, self.W = H, W
        self.pdf_scale = (H * W) / (2 * np.pi * np.pi)
        self.update_pdf()

    def update_pdf(self):
        H, W = self.H, self.W
        Y = ((torch.arange(H, dtype=torch.float32) + 0.5) / H)[:, None].expand(H, W)
        pdf = torch.max(self.base, dim=-1)[0] * torch.sin(Y * np.pi)
        pdf = torch.where(pdf <= 0, torch.full_like(pdf, 1e-6), pdf)
        pdf = pdf / torch.sum(pdf)
        self._pdf = pdf
        cols = torch.cumsum(pdf, dim=1)
        rows = torch.cumsum(cols[:, -1:].repeat(1, W), dim=0)
        cols = cols / torch.where(cols[:, -1:] > 0, cols[:, -1:], torch.ones_like(cols))
        rows = rows / torch.where(rows[-1:, :] > 0, rows[-1:, :], torch.ones_like(rows))
        self.cols = torch.cat([torch.zeros_like(cols[:, :1]), cols], dim=1)      # [H, W+1]
        self.rows = torch.cat([torch.zeros_like(rows[:1, :]), rows], dim=0)      # [H+1, W]

    def sample(self, u1: torch.Tensor, u2: torch.Tensor) -> torch.Tensor:
        """Inverse-CDF light directions from explicit uniforms (light.py:341-412)."""
        rows0 = self.rows[:, 0].contiguous()
        ri = torch.searchsorted(rows0, u1, right=True)
        below = torch.clamp(ri - 1, min=0)
        above = torch.clamp(ri, max=self.rows.shape[0] - 1)
        rfrac = (u1 - rows0[below]) / (rows0[above] - rows0[below])
        ri = below
        ci = torch.searchsorted(self.cols[ri], u2[:, None].contiguous(), right=True).squeeze(-1)
        below = torch.clamp(ci - 1, min=0)
        above = torch.clamp(ci, max=self.cols.shape[-1] - 1)
        cfrac = (u2 - self.cols[ri, below]) / (self.cols[ri, above] - self.cols[ri, below])
        ci = below
        u = (ci + cfrac) / self.W
        v = (ri + rfrac) / self.H
        lon = (u - 0.5) * 2 * np.pi
        lat = (v - 0.5) * np.pi
        d = torch.stack([torch.cos(lat) * torch.sin(lon), torch.sin(lat), torch.cos(lat) * torch.cos(lon)], -1)
        return F.normalize(d, dim=1)

    @staticmethod
    def _uv(d):
        lon = torch.atan2(d[:, 0], d[:, 2])
        lat = torch.asin(d[:, 1] / torch.linalg.norm(d, dim=-1))
        return lon / (2 * np.pi) + 0.5, lat / np.pi + 0.5, lat

    def pdf(self, d: torch.Tensor) -> torch.Tensor:
        u, v, lat = self._uv(d)
        col = torch.clamp(torch.floor(u * self.W), min=0, max=self.W - 1).long()
        row = torch.clamp(torch.floor(v * self.H), min=0, max=self.H - 1).long()
        sin_theta = torch.sin(np.pi / 2.0 - lat)
        val = torch.where(sin_theta > 0, self._pdf[row, col] * self.pdf_scale / sin_theta,
                          torch.zeros_like(sin_theta))
        return val[:, None]

    def eval(self, d: torch.Tensor) -> torch.Tensor:
        u, v, _ = self._uv(d)
        grid = torch.stack([u * 2 - 1, v * 2 - 1], dim=-1).reshape(1, 1, -1, 2)
        img = self.base[None].permute(0, 3, 1, 2)
        out = F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=True)
        return out.reshape(3, -1).t()


# ------------------------------------------------------------------------------------ BRDF ----
def multilobe_eval(wi, n, wo, roughness, albedo, metallic):
    """MultiLobe.eval with attenuation = 0: returns diff [N,1], spec [N,3], both incl. cos.

    wi = direction towards the viewer (= -ray dir), wo = light direction, n = unit normal."""
    eps = 1e-6
    F0 = 0.04 * (1.0 - metallic) + albedo * metallic
    diff = F.relu((wo * n).sum(-1, keepdim=True)) / np.pi
    # local frame (coordinate_system): only z and dot products matter for the isotropic lobe,
    # but build it like the reference to keep its rounding.
    a = n
    cond = a[:, 0].abs() > a[:, 1].abs()
    inv1 = 1.0 / torch.sqrt(a[:, 0] ** 2 + a[:, 2] ** 2)
    inv2 = 1.0 / torch.sqrt(a[:, 1] ** 2 + a[:, 2] ** 2)
    c1 = torch.stack([a[:, 2] * inv1, torch.zeros_like(inv1), -a[:, 0] * inv1], -1)
    c2 = torch.stack([torch.zeros_like(inv2), a[:, 2] * inv2, -a[:, 1] * inv2], -1)
    c = torch.where(cond[:, None], c1, c2)
    b = torch.cross(c, a, dim=-1)

    def to_local(v):
        return torch.stack([(v * b).sum(-1), (v * c).sum(-1), (v * a).sum(-1)], -1)

    wo_l, wi_l = to_local(wo), to_local(wi)
    wh = F.normalize(wi_l + wo_l, dim=-1)
    alpha = roughness
    k = (alpha ** 2 + 2 * alpha + 1) / 8.0
    cos2 = wh[:, 2] ** 2
    alpha2 = alpha ** 2
    ndf = alpha2 / (np.pi * (cos2 * (alpha2 - 1) + 1) ** 2 + eps)

    def g1(v):
        nom = v[:, 2]
        den = nom * (1.0 - k) + k
        return torch.where(den > eps, nom / (den + eps), torch.zeros_like(nom))

    cos_ih = (wi_l * wh).sum(-1, keepdim=True).abs()
    fres = F0 + (1.0 - F0) * 2 ** ((-5.55473 * cos_ih - 6.98316) * cos_ih)
    val = (ndf * g1(wi_l) * g1(wo_l))[:, None] * fres / (4 * wi_l[:, 2:] + eps)
    ok = (wi_l[:, 2:] > eps) & (wo_l[:, 2:] > eps)
    spec = torch.where(ok, val, torch.zeros_like(val))
    return diff, spec


def _frame(n):
    """coordinate_system (warp_utils.py:62-101): returns (t, b) with t = cross(b, n)."""
    a = n
    cond = a[:, 0].abs() > a[:, 1].abs()
    inv1 = 1.0 / torch.sqrt(a[:, 0] ** 2 + a[:, 2] ** 2)
    inv2 = 1.0 / torch.sqrt(a[:, 1] ** 2 + a[:, 2] ** 2)
    c1 = torch.stack([a[:, 2] * inv1, torch.zeros_like(inv1), -a[:, 0] * inv1], -1)
    c2 = torch.stack([torch.zeros_like(inv2), a[:, 2] * inv2, -a[:, 1] * inv2], -1)
    c = torch.where(cond[:, None], c1, c2)
    return torch.cross(c, a, dim=-1), c


def _to_local(v, t, b, n):
    return torch.stack([(v * t).sum(-1), (v * b).sum(-1), (v * n).sum(-1)], -1)


def _to_world(v, t, b, n):
    return F.normalize(v[:, 0:1] * t + v[:, 1:2] * b + v[:, 2:3] * n, dim=-1)


def luminance(x):
    return x[..., :1] * 0.212671 + x[..., 1:2] * 0.715160 + x[..., 2:3] * 0.072169


def _lobe_weights(wi, n, albedo, metallic):
    """Lobe selection weights shared by MultiLobe.pdf / .sample (bxdf.py:297-312, 343-358): note the
    Fresnel term uses the albedo itself as F0."""
    wd = (1.0 - metallic) * luminance(albedo)
    cos_t = (wi * n).sum(-1, keepdim=True)
    fres = albedo + (1.0 - albedo) * 2 ** ((-5.55473 * cos_t - 6.98316) * cos_t)
    ws = torch.where(cos_t > 0, luminance(fres), torch.zeros_like(cos_t))
    return wd, ws


def multilobe_pdf(wi, n, wo, roughness, albedo, metallic):
    """MultiLobe.pdf in eval mode (bxdf.py:290-317) = p_d * Lambertian.pdf (:117-123, cosine / pi)
    + (1 - p_d) * GGX.pdf (:222-236, VNDF / (4 |wo.wh| + eps)); roughness [N], albedo [N,3], metallic [N,1]."""
    eps = 1e-6
    wd, ws = _lobe_weights(wi, n, albedo, metallic)
    p_d = torch.where(wd + ws > eps, wd / (wd + ws + eps), torch.ones_like(wd))
    pdf_d = (F.relu((n * wo).sum(-1)) / np.pi)[:, None]
    t, b = _frame(n)
    wo_l, wi_l = _to_local(wo, t, b, n), _to_local(wi, t, b, n)
    wh = F.normalize(wi_l + wo_l, dim=-1)
    alpha = roughness
    k = (alpha ** 2 + 2 * alpha + 1) / 8.0
    nom = wi_l[:, 2]
    den = nom * (1.0 - k) + k
    g1 = torch.where(den > eps, nom / (den + eps), torch.zeros_like(nom))
    alpha2 = alpha ** 2
    ndf = alpha2 * torch.reciprocal(np.pi * (wh[:, 2] ** 2 * (alpha2 - 1) + 1) ** 2 + eps)
    vndf = torch.where((wh[:, 2] > eps) & (wi_l[:, 2] > eps),
                       g1 * torch.clamp((wh * wi_l).sum(-1), min=0.0) * ndf / (wi_l[:, 2] + eps),
                       torch.zeros_like(nom))
    pdf_s = torch.where(4 * (wi_l * wh).sum(-1).abs() > eps, vndf / (4 * (wo_l * wh).sum(-1).abs() + eps),
                        torch.zeros_like(nom))[:, None]
    return p_d * pdf_d + (1 - p_d) * pdf_s


def multilobe_sample(n, wi, roughness, albedo, metallic, sample):
    """MultiLobe.sample in eval mode (bxdf.py:332-388) with explicit uniforms ``sample`` [N,2]:
    lobe pick on sample[:,0] (rescaled afterwards), GGX VNDF sampling (warp_utils.py:632-690) or
    cosine-weighted hemisphere via the concentric disk map (warp_utils.py:139-172, 599-616)."""
    eps = 1e-6
    wd, ws = _lobe_weights(wi, n, albedo, metallic)
    p_s = torch.where(wd + ws > eps, ws / (wd + ws + eps), torch.zeros_like(wd)).squeeze(-1)
    spec_mask = p_s > sample[:, 0]
    s0 = torch.where(spec_mask, sample[:, 0] / p_s, (sample[:, 0] - p_s) / (1 - p_s))
    s1 = sample[:, 1]
    t, b = _frame(n)
    # --- specular: VNDF
    wi_l = _to_local(wi, t, b, n)
    a = roughness
    vh = F.normalize(torch.stack([a * wi_l[:, 0], a * wi_l[:, 1], wi_l[:, 2]], -1), dim=-1)
    lensq = vh[:, 0] * vh[:, 0] + vh[:, 1] * vh[:, 1]
    T1 = torch.where(lensq[:, None] > eps,
                     torch.stack([-vh[:, 1] / torch.sqrt(lensq + eps), vh[:, 0] / torch.sqrt(lensq + eps),
                                  torch.zeros_like(lensq)], -1),
                     torch.tensor([[1.0, 0.0, 0.0]]).expand(len(lensq), 3))
    T2 = torch.cross(vh, T1, dim=-1)
    r = torch.sqrt(s0)
    phi = 2.0 * np.pi * s1
    t1 = r * torch.cos(phi)
    t2 = r * torch.sin(phi)
    sv = 0.5 * (1.0 + vh[:, 2])
    t2 = (1.0 - sv) * torch.sqrt(torch.clamp(1.0 - t1 * t1, min=0.0)) + sv * t2
    nh = t1[:, None] * T1 + t2[:, None] * T2 + torch.sqrt(torch.clamp(1.0 - t1 * t1 - t2 * t2, min=0.0))[:, None] * vh
    wh = F.normalize(torch.stack([a * nh[:, 0], a * nh[:, 1], torch.clamp(nh[:, 2], min=0.0)], -1), dim=-1)
    wo_s = _to_world(2 * (wi_l * wh).sum(-1, keepdim=True) * wh - wi_l, t, b, n)
    # --- diffuse: concentric disk -> hemisphere
    ox, oy = 2.0 * s0 - 1.0, 2.0 * s1 - 1.0
    big = ox.abs() > oy.abs()
    rr = torch.where(big, ox, oy)
    th = torch.where(big, np.pi / 4.0 * (oy / ox), np.pi / 2.0 - np.pi / 4.0 * (ox / oy))
    x, y = rr * torch.cos(th), rr * torch.sin(th)
    z = torch.sqrt((1.0 - x ** 2 - y ** 2).clamp(min=0.0))
    wo_d = _to_world(torch.stack([x, y, z], -1), t, b, n)
    return torch.where(spec_mask[:, None], wo_s, wo_d)


def uniform_sphere_stratified(n_rows=16, n_cols=32):
    """EnvironmentLightBase.sample_uniform_sphere_stratified in eval mode (light.py:161-217; no jitter):
    cell centres of an n_rows x n_cols grid mapped by sample_uniform_sphere (warp_utils.py:174-198;
    the first column -> z).  Returns directions [n_rows*n_cols, 3]; inv_pdf is 4 pi for all of them."""
    v, u = torch.meshgrid(torch.arange(0, n_rows, dtype=torch.float32) + 0.5,
                          torch.arange(0, n_cols, dtype=torch.float32) + 0.5, indexing="ij")
    u = u / n_cols
    v = v / n_rows
    s = torch.stack([v, u], -1).reshape(-1, 2)
    z = s[:, 0] * 2.0 - 1.0
    phi = 2.0 * np.pi * s[:, 1]
    r = torch.sqrt(torch.clamp(1.0 - z ** 2, min=0.0))
    return F.normalize(torch.stack([torch.cos(phi) * r, torch.sin(phi) * r, z], -1), dim=-1)


def rgb_to_srgb(f: torch.Tensor) -> torch.Tensor:
    return torch.where(f <= 0.0031308, f * 12.92,
                       torch.pow(torch.clamp(f, 0.0031308), 1.0 / 2.4) * 1.055 - 0.055)


# ------------------------------------------------------------------ stateless permutation ----
def _u32(x):
    return x & np.uint64(0xFFFFFFFF)


def kensler_permute(i: np.ndarray, l: int, p: np.ndarray) -> np.ndarray:
    """Kensler, "Correlated Multi-Jittered Sampling" (2013), listing ``permute(i, l, p)``:
    a keyed bijection on [0, l).  i, p uint32 arrays (broadcastable)."""
    i = np.asarray(i, np.uint64)
    p = np.asarray(p, np.uint64)
    i, p = np.broadcast_arrays(i, p)
    i = i.copy()
    w = np.uint64(l - 1)
    w |= w >> np.uint64(1)
    w |= w >> np.uint64(2)
    w |= w >> np.uint64(4)
    w |= w >> np.uint64(8)
    w |= w >> np.uint64(16)
    todo = np.ones(i.shape, bool)
    while todo.any():
        x = i.copy()
        x ^= p
        x = _u32(x * np.uint64(0xe170893d))
        x ^= p >> np.uint64(16)
        x ^= (x & w) >> np.uint64(4)
        x ^= p >> np.uint64(8)
        x = _u32(x * np.uint64(0x0929eb3f))
        x ^= p >> np.uint64(23)
        x ^= (x & w) >> np.uint64(1)
        x = _u32(x * (np.uint64(1) | (p >> np.uint64(27))))
        x = _u32(x * np.uint64(0x6935fa69))
        x ^= (x & w) >> np.uint64(11)
        x = _u32(x * np.uint64(0x74dcb303))
        x ^= (x & w) >> np.uint64(2)
        x = _u32(x * np.uint64(0x9e501cc3))
        x ^= (x & w) >> np.uint64(2)
        x = _u32(x * np.uint64(0xc860a3df))
        x &= w
        x ^= x >> np.uint64(5)
        i = np.where(todo, x, i)
        todo = todo & (i >= np.uint64(l))
    return (_u32(i + p) % np.uint64(l)).astype(np.int64)


def pixel_key(seed: int, ray_index: np.ndarray) -> np.ndarray:
    """32-bit mix of (seed, ray index) used as the permutation key (lowbias32 finaliser)."""
    x = (np.asarray(ray_index, np.uint64) * np.uint64(0x9E3779B1) + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = _u32(x * np.uint64(0x7feb352d))
    x ^= x >> np.uint64(15)
    x = _u32(x * np.uint64(0x846ca68b))
    x ^= x >> np.uint64(16)
    return x


def _mix32(x):
    x = np.asarray(x, np.uint64)
    x = x ^ (x >> np.uint64(16))
    x = _u32(x * np.uint64(0x7feb352d))
    x = x ^ (x >> np.uint64(15))
    x = _u32(x * np.uint64(0x846ca68b))
    x = x ^ (x >> np.uint64(16))
    return x


def rng_uniform(key: np.ndarray, j: np.ndarray, dim: int) -> torch.Tensor:
    """Counter-based uniform in [0,1) (24 bits) standing in for the reference's torch.rand in
    MultiLobe.sample / emitter.sample: stream ``dim`` of shading sample ``j`` of the pixel with key ``key``
    (= pixel_key(seed, ray index)).  Same integer recipe as ia_rng_uniform (csrc/ia_pbr.cuh)."""
    key = np.asarray(key, np.uint64)
    j = np.asarray(j, np.uint64)
    inner = _u32(j * np.uint64(0x9E3779B9) + np.uint64(dim) * np.uint64(0x85EBCA6B) + np.uint64(0x6A09E667))
    x = _mix32(key ^ _mix32(inner))
    return torch.from_numpy(((x >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)))
