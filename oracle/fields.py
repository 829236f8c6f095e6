"""ORACLE (test infrastructure, not product): canonical fields on CPU, plain PyTorch fp32.

Restates, for the render path only:
  * tiny-cuda-nn ``HashGrid`` forward (NOT in /root/reference; unpinned git master, README.md:48;
    algorithm restated from the published tcnn kernel as recalled in SURVEY.md Appendix B;
    call site models/network_utils.py:58-79)                                    -> ``hashgrid``
  * tcnn ``SphericalHarmonics`` degree 4 (configs/radiance/progressive_hash_grid.yaml:17-19) -> ``sh4``
  * ``VolumeSDF.forward`` incl. autograd gradient (models/rf/geometry.py:124-172)      -> ``geometry``
  * ``VolumeRefDirRadiance.forward`` (models/rf/radiance.py:111-135)                    -> ``radiance``
  * ``VolumeMaterial.forward`` (models/pbr/material.py:31-51)                           -> ``material``
  * ``LearnedLaplaceDensity.density_func`` + ``get_alpha`` (models/rf/density.py:25-34;
    models/intrinsic_avatar.py:390-394)                                                 -> ``alpha_from_sdf``

Parity status: hash grid / SH are "parity unpinned" (third-party source absent); everything
else follows first-party reference code line by line.  The gradient is taken with autograd,
like the reference, so it is an independent check of the kernels' analytic derivative.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

PRIMES = (1, 2654435761, 805459861)


def hashgrid(x: torch.Tensor, table: torch.Tensor, layout: dict) -> torch.Tensor:
    """x [N,3] in [0,1] -> [N, 32] (level-major, feature-minor). Differentiable w.r.t. x."""
    N = x.shape[0]
    feats = []
    tab = table.view(-1, 2)
    for l in range(len(layout["res"])):
        scale = float(layout["scale"][l])
        res = int(layout["res"][l])
        size = int(layout["size"][l])
        off = int(layout["offset"][l])
        pos = x * scale + 0.5
        g = torch.floor(pos).detach()
        w = pos - g
        g = g.to(torch.int64)
        acc = torch.zeros(N, 2, dtype=x.dtype)
        for corner in range(8):
            idx3, wt = [], torch.ones(N, dtype=x.dtype)
            for d in range(3):
                bit = (corner >> d) & 1
                idx3.append(g[:, d] + bit)
                wt = wt * (w[:, d] if bit else (1.0 - w[:, d]))
            # dense index with the tcnn stride rule, else coherent prime hash (uint32 wrap)
            stride, index, d = 1, torch.zeros(N, dtype=torch.int64), 0
            while d < 3 and stride <= size:
                index = index + (idx3[d] & 0xFFFFFFFF) * stride
                stride *= res
                d += 1
            if size < stride:
                index = torch.zeros(N, dtype=torch.int64)
                for d in range(3):
                    index = index ^ (((idx3[d] & 0xFFFFFFFF) * PRIMES[d]) & 0xFFFFFFFF)
            index = (index & 0xFFFFFFFF) % size
            acc = acc + wt[:, None] * tab[off + index]
        feats.append(acc)
    return torch.cat(feats, dim=-1)


def sh4(d: torch.Tensor) -> torch.Tensor:
    """16 real SH basis values of unit direction d [N,3] (tcnn SphericalHarmonics degree 4)."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    out = [
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y,
        0.48860251190291987 * z,
        -0.48860251190291987 * x,
        1.0925484305920792 * xy,
        -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ]
    return torch.stack(out, dim=-1)


class Fields:
    """Folded weights (intrinsicavatar_b200.weights.fold) + canonical bbox -> field queries."""

    def __init__(self, folded: dict, layout: dict, bbox):
        self.w = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in folded.items()}
        self.layout = layout
        bbox = torch.as_tensor(bbox, dtype=torch.float32)
        self.center = (bbox[0] + bbox[1]) / 2          # geometry.py:61-68 prepare_bbox
        self.scale = bbox[1] - bbox[0]
        self.beta = float(folded["beta"])
        # material scales/biases, configs/material/shallow_mlp.yaml
        self.mat_scale = torch.tensor([0.77, 0.77, 0.77, 0.9, 1.0])
        self.mat_bias = torch.tensor([0.03, 0.03, 0.03, 0.09, 0.0])

    # -- geometry ---------------------------------------------------------------------------
    def _geo_net(self, xn):
        enc = hashgrid(xn, self.w["geo_hash"], self.layout)
        inp = torch.cat([xn * 2.0 - 1.0, enc], dim=-1)                      # CompositeEncoding include_xyz
        h = F.softplus(F.linear(inp, self.w["geo_w1"], self.w["geo_b1"]), beta=100)
        return F.linear(h, self.w["geo_w2"], self.w["geo_b2"])

    def geometry(self, x, with_grad=False):
        """x [M,3] canonical metric -> sdf [M], feature [M,13] (, grad [M,3])."""
        if x.shape[0] == 0:
            e = torch.zeros(0)
            return (e, torch.zeros(0, 13), torch.zeros(0, 3)) if with_grad else (e, torch.zeros(0, 13))
        if with_grad:
            with torch.enable_grad():
                xg = x.clone().requires_grad_(True)
                xn = (xg - self.center) / self.scale + 0.5
                out = self._geo_net(xn)
                sdf = out[:, 0]
                (grad,) = torch.autograd.grad(sdf, xg, torch.ones_like(sdf))
            return sdf.detach(), out.detach(), grad.detach()
        with torch.no_grad():
            xn = (x - self.center) / self.scale + 0.5
            out = self._geo_net(xn)
        return out[:, 0], out

    # -- radiance / material ----------------------------------------------------------------
    @torch.no_grad()
    def radiance(self, x, feature, view_dir_world, normal_world):
        """-> rgb [M,3], xyz_embd [M,35] (radiance.py:111-135; all SH bands on)."""
        xn = (x - self.center) / self.scale + 0.5
        enc = hashgrid(xn, self.w["rad_hash"], self.layout)
        xyz_embd = torch.cat([xn * 2.0 - 1.0, enc], dim=-1)
        v = -view_dir_world
        refl = 2.0 * (v * normal_world).sum(-1, keepdim=True) * normal_world - v   # models/utils.py:115
        d01 = (refl + 1.0) / 2.0
        sh = sh4(d01 * 2.0 - 1.0)                                                  # tcnn maps [0,1]->[-1,1]
        inp = torch.cat([xyz_embd, feature, sh, normal_world], dim=-1)            # 35+13+16+3 = 67
        h = F.relu(F.linear(inp, self.w["rad_w1"], self.w["rad_b1"]))
        h = F.relu(F.linear(h, self.w["rad_w2"], self.w["rad_b2"]))
        rgb = torch.sigmoid(F.linear(h, self.w["rad_w3"], self.w["rad_b3"]))
        return rgb, xyz_embd

    @torch.no_grad()
    def material(self, xyz_embd, feature):
        """hybrid material feature = cat[xyz_embd(35), feature(13)] -> [M,5] (material.py:31-51)."""
        inp = torch.cat([xyz_embd, feature], dim=-1)
        h = F.relu(F.linear(inp, self.w["mat_w1"], self.w["mat_b1"]))
        h = F.relu(F.linear(h, self.w["mat_w2"], self.w["mat_b2"]))
        m = torch.sigmoid(F.linear(h, self.w["mat_w3"], self.w["mat_b3"]))
        return m * self.mat_scale + self.mat_bias

    # -- density ----------------------------------------------------------------------------
    def alpha_from_sdf(self, sdf, dists):
        """Laplace-CDF density then alpha (density.py:25-30; intrinsic_avatar.py:390-394)."""
        beta = self.beta
        sigma = (1.0 / beta) * (0.5 + 0.5 * torch.sign(sdf) * torch.expm1(-sdf.abs() / beta))
        return 1.0 - torch.exp(-sigma * dists)
