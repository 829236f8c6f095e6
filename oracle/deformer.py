"""ORACLE (test infrastructure, not product): fast-SNARF posed->canonical search on CPU.

Restates the reference's first-party CUDA kernels and their torch glue in vectorised fp32 PyTorch:
  * ``precompute_kernel``  models/deformers/fast_snarf/cuda/precompute/precompute.cu:22-71
  * ``broyden_kernel``     .../cuda/fuse_kernel/fuse_cuda_kernel_fast.cu:250-413
        (trilinear sampler with zero padding, align_corners=True: :110-248; rank-1 update :22-55)
  * ``filter``             .../cuda/filter/filter.cu:10-54
  * ``ForwardDeformer.forward/forward_skinning/query_weights``  deformer_torch.py:35-55,127-137,199-227
  * ``SNARFDeformer.deform``  models/deformers/snarf_deformer.py:187-261

The reference kernels are fp32 CUDA compiled with default nvcc flags (FMA contraction on); this
restatement uses separate mul/add, so converged roots agree to ~1e-6 and the discrete ``valid``
flags can differ on a measure-zero set of threshold cases (budgeted in the tests).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

INIT_BONES = [0, 1, 2, 4, 5, 10, 11, 12, 15, 16, 17, 18, 19]
CVG = 1e-5
DVG = 1e-1


def precompute(lbs_voxel: torch.Tensor, tfs: torch.Tensor) -> torch.Tensor:
    """lbs_voxel [24,D,H,W], tfs [24,4,4] -> voxel_J [12,D,H,W] (rows 0..2 of the blended 4x4)."""
    J = torch.einsum("jdhw,jab->abdhw", lbs_voxel, tfs[:, :3, :])
    return J.reshape(12, *lbs_voxel.shape[1:]).contiguous()


def _trilinear_zero_pad(vol: torch.Tensor, gx, gy, gz):
    """vol [C,D,H,W]; normalised coords in [-1,1] (x->W, y->H, z->D), align_corners=True,
    zero outside (fuse_cuda_kernel_fast.cu:110-232). Returns [N,C]."""
    C, D, H, W = vol.shape
    ix = (gx + 1.0) / 2 * (W - 1)
    iy = (gy + 1.0) / 2 * (H - 1)
    iz = (gz + 1.0) / 2 * (D - 1)
    bad = ~(torch.isfinite(ix) & torch.isfinite(iy) & torch.isfinite(iz))
    ix = torch.where(bad, torch.full_like(ix, -100.0), ix)
    iy = torch.where(bad, torch.full_like(iy, -100.0), iy)
    iz = torch.where(bad, torch.full_like(iz, -100.0), iz)
    ix = ix.clamp(-1e6, 1e6)
    iy = iy.clamp(-1e6, 1e6)
    iz = iz.clamp(-1e6, 1e6)
    x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    flat = vol.reshape(C, -1)
    out = torch.zeros(ix.shape[0], C, dtype=vol.dtype)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                wx = (ix - x0) if dx else (x0 + 1 - ix)
                wy = (iy - y0) if dy else (y0 + 1 - iy)
                wz = (iz - z0) if dz else (z0 + 1 - iz)
                inb = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H) & (zi >= 0) & (zi < D)
                lin = (zi.clamp(0, D - 1) * H + yi.clamp(0, H - 1)) * W + xi.clamp(0, W - 1)
                v = flat[:, lin.long()].t()
                out = out + torch.where(inb[:, None], v * (wx * wy * wz)[:, None], torch.zeros_like(v))
    return out


def broyden(xd: torch.Tensor, voxel_J: torch.Tensor, tfs: torch.Tensor, offset, scale):
    """xd [N,3] -> x [N,13,3], J_inv [N,13,3,3], valid [N,13] (un-filtered).

    offset/scale are the reference's ``offset_kernel`` / ``scale_kernel`` [3]."""
    N = xd.shape[0]
    I = len(INIT_BONES)
    bones = torch.tensor(INIT_BONES)
    tgt = xd[:, None, :].expand(N, I, 3).reshape(-1, 3)
    R = tfs[bones, :3, :3]
    t = tfs[bones, :3, 3]
    rel = xd[:, None, :] - t[None]                                  # [N,I,3]
    xl = torch.einsum("nia,iab->nib", rel, R).reshape(-1, 3)       # R^T (x - t)

    def fetch(x):
        g = scale[None] * (x + offset[None])
        J = _trilinear_zero_pad(voxel_J, g[:, 0], g[:, 1], g[:, 2]).reshape(-1, 3, 4)
        return J, g

    def apply(J, x):
        return (J[:, :, :3] * x[:, None, :]).sum(-1) + J[:, :, 3]

    P = N * I
    J, _ = fetch(xl)
    Jinv = J[:, :, :3].transpose(1, 2).clone()                      # init: transpose of blended rotation
    gx = apply(J, xl) - tgt
    out_x = torch.zeros(P, 3)
    out_J = torch.zeros(P, 3, 3)
    valid = torch.zeros(P, dtype=torch.bool)
    ids = torch.arange(P)                                           # still-iterating chains (compacted)
    for _ in range(10):
        if ids.numel() == 0:
            break
        u = -(Jinv * gx[:, None, :]).sum(-1)
        xl = xl + u
        Jn, g = fetch(xl)
        gx_new = apply(Jn, xl) - tgt
        nrm = (gx_new * gx_new).sum(-1)
        conv = nrm < CVG * CVG
        inb = (g >= -1).all(-1) & (g <= 1).all(-1)
        ok = conv & inb
        out_x[ids[ok]] = xl[ok]
        out_J[ids[ok]] = Jinv[ok]
        valid[ids[ok]] = True
        keep = ~conv & ~(nrm > DVG * DVG)
        # rank-1 inverse-Jacobian update (fuse_J_inv_update)
        dx, dg = u[keep], (gx_new - gx)[keep]
        Jk = Jinv[keep]
        c = torch.einsum("pji,pj->pi", Jk, dx)                      # J^T dx
        s = (c * dg).sum(-1, keepdim=True)
        r = -torch.einsum("pij,pj->pi", Jk, dg)
        Jinv = Jk + (r + dx)[:, :, None] * c[:, None, :] / s[:, :, None]
        gx, xl, tgt, ids = gx_new[keep], xl[keep], tgt[keep], ids[keep]
    return out_x.reshape(N, I, 3), out_J.reshape(N, I, 3, 3), valid.reshape(N, I)


def filter_duplicates(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Drop candidate i if a later valid candidate j lies within 1e-4 (filter.cu:25-50)."""
    N, I, _ = x.shape
    d2 = ((x[:, :, None, :] - x[:, None, :, :]) ** 2).sum(-1)       # [N,I,I]
    later = torch.triu(torch.ones(I, I, dtype=torch.bool), diagonal=1)[None]
    dup = (d2 < 0.0001 * 0.0001) & later & mask[:, None, :]
    return mask & ~dup.any(-1)


def forward_rotation(xc: torch.Tensor, lbs_voxel: torch.Tensor, tfs: torch.Tensor, offset, scale):
    """Blended forward LBS 3x3 at canonical points (deformer_torch.py:127-137,199-227): border padding."""
    g = scale[None] * (xc + offset[None])
    w = F.grid_sample(lbs_voxel[None], g[None, :, None, None, :], align_corners=True, mode="bilinear",
                      padding_mode="border")[0, :, :, 0, 0].t()      # [M,24]
    return torch.einsum("pn,nij->pij", w, tfs)[:, :3, :3]


def implicit_correction(xc_opt: torch.Tensor, valid: torch.Tensor, J_inv: torch.Tensor, lbs_voxel: torch.Tensor,
                        tfs: torch.Tensor, offset, scale) -> torch.Tensor:
    """Training-mode ``ForwardDeformer.forward`` (version 1, deformer_torch.py:57-76) after the search: the value is the root,
    the gradient reaches ``tfs`` through  x_c = x_c* - J_inv (x_d(x_c*) - stopgrad(x_d(x_c*)))  with
    x_d = skinning_mask(x_c*, w(x_c*), tfs) (:127-137, 199-227).  xc_opt [N,13,3], valid [N,13], J_inv [N,13,3,3],
    tfs [24,4,4] (requires_grad) -> x_c [N,13,3]."""
    xc_opt = xc_opt.detach().clone()
    xc_opt[~valid] = 0
    pts = xc_opt[valid]
    g = scale[None] * (pts + offset[None])
    w = F.grid_sample(lbs_voxel[None], g[None, :, None, None, :], align_corners=True, mode="bilinear",
                      padding_mode="border")[0, :, :, 0, 0].t()      # [M,24]
    w_tf = torch.einsum("pn,nij->pij", w, tfs)
    x_h = F.pad(pts, (0, 1), value=1.0)
    xd_opt = (w_tf * x_h[:, None, :]).sum(-1)[:, :3]
    correction = xd_opt - xd_opt.detach()
    correction = torch.einsum("pij,pj->pi", -J_inv[valid], correction)
    xc = xc_opt.clone()
    xc[valid] = xc[valid] + correction
    return xc


def deform(xd: torch.Tensor, fields, voxel_J, lbs_voxel, tfs, offset, scale, with_grad: bool):
    """SNARFDeformer.deform (snarf_deformer.py:187-261) with the dummy non-rigid deformer.

    Returns dict: x_c [N,3], sdf [N], valid [N] and, if with_grad, grad (posed), grad_cano, feature."""
    N = xd.shape[0]
    x, _, valid = broyden(xd, voxel_J, tfs, offset, scale)
    valid = filter_duplicates(x, valid)
    x = torch.where(valid[..., None], x, torch.zeros_like(x))
    sdf = torch.full((N, 13), 1e5)
    res = {}
    flat_valid = valid.reshape(-1)
    pts = x.reshape(-1, 3)[flat_valid]
    if with_grad:
        s, feat, g = fields.geometry(pts, with_grad=True)
        Rf = forward_rotation(pts, lbs_voxel, tfs, offset, scale)
        grad_cano = torch.tensor([0.0, 0.0, 1.0]).repeat(N * 13, 1)
        grad = grad_cano.clone()
        feature = torch.zeros(N * 13, 13)
        grad_cano[flat_valid] = g
        grad[flat_valid] = torch.einsum("bij,bj->bi", Rf, g)
        feature[flat_valid] = feat
    else:
        s, _ = fields.geometry(pts, with_grad=False)
    sdf.view(-1)[flat_valid] = s
    sdf_min, idx = torch.min(sdf, dim=-1)
    ar = torch.arange(N)
    res["x_c"] = x[ar, idx]
    res["sdf"] = sdf_min
    res["valid"] = valid.any(-1)
    if with_grad:
        res["grad"] = grad.view(N, 13, 3)[ar, idx]
        res["grad_cano"] = grad_cano.view(N, 13, 3)[ar, idx]
        res["feature"] = feature.view(N, 13, 13)[ar, idx]
    res["n_valid"] = valid.sum(-1)
    return res
