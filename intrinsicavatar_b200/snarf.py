"""Host-side fast-SNARF set-up: what the reference does once per subject and once per frame
*before* any kernel runs.

Mirrors (reference file:line):
  * ``ForwardDeformer.switch_to_explicit`` + ``query_weights_smpl``
    (models/deformers/fast_snarf/deformer_torch.py:139-197, 234-253): voxelise the body's
    skinning weights on a 128x128x32 grid (K=30 inverse-distance + 30 smoothing passes).
  * ``SNARFDeformer.initialize`` / ``prepare_deformer``
    (models/deformers/snarf_deformer.py:46-126): canonical A-pose inverse transforms,
    per-frame ``tfs = w2s . A . A_cano^-1``, root-frame vertices, canonical bbox.
  * ``get_bbox_from_smpl`` (snarf_deformer.py:24-35).

Init-only / per-frame 24-matrix host math; the per-voxel work (``precompute``) is a CUDA
kernel behind the C-ABI (csrc/deform.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from .body import SyntheticBody, a_pose

INIT_BONES = [0, 1, 2, 4, 5, 10, 11, 12, 15, 16, 17, 18, 19]  # deformer_torch.py:27
GLOBAL_SCALE = 1.2                                              # deformer_torch.py:32


def get_bbox_from_verts(vs: np.ndarray, factor: float = 1.2) -> np.ndarray:
    """Cube bbox around vertices [V,3] -> [2,3] (snarf_deformer.py:24-35)."""
    mn, mx = vs.min(0), vs.max(0)
    c = (mx + mn) / 2
    s = ((mx - mn) / 2).max() * factor
    return np.stack([c - s, c + s]).astype(np.float32)


def voxelize_lbs_weights(verts: np.ndarray, weights: np.ndarray, resolution: int = 128):
    """Skinning-weight voxel grid, reference recipe (deformer_torch.py:139-197, 234-253).

    Returns dict with ``lbs_voxel`` [24, D, H, W] float32 (D = resolution // 4),
    ``offset_kernel`` [3] (= -centre), ``scale_kernel`` [3] (= 1/scale, z multiplied by 4).
    """
    from scipy.spatial import cKDTree

    d, h, w = resolution // 4, resolution, resolution
    ratio = h / d
    mn, mx = verts.min(0), verts.max(0)
    offset = (mn + mx) * 0.5
    scale = float((mx - mn).max() / 2 * GLOBAL_SCALE)

    xs = np.linspace(-1, 1, w, dtype=np.float32)
    ys = np.linspace(-1, 1, h, dtype=np.float32)
    zs = np.linspace(-1, 1, d, dtype=np.float32)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")  # [d,h,w], w fastest
    grid = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float64)
    grid[:, 2] /= ratio
    grid = grid * scale + offset  # denormalize

    tree = cKDTree(verts.astype(np.float64))
    dist, idx = tree.query(grid, k=30)
    dist = np.clip(dist, 0.0001, 1.0)
    ws = 1.0 / dist
    ws = ws / ws.sum(-1, keepdims=True)
    vox = np.einsum("pk,pkj->pj", ws, weights[idx].astype(np.float64)).astype(np.float32)  # [P,24]
    vox = torch.from_numpy(vox).t().reshape(1, 24, d, h, w).contiguous()
    for _ in range(30):
        mean = (vox[:, :, 2:, 1:-1, 1:-1] + vox[:, :, :-2, 1:-1, 1:-1]
                + vox[:, :, 1:-1, 2:, 1:-1] + vox[:, :, 1:-1, :-2, 1:-1]
                + vox[:, :, 1:-1, 1:-1, 2:] + vox[:, :, 1:-1, 1:-1, :-2]) / 6.0
        vox[:, :, 1:-1, 1:-1, 1:-1] = (vox[:, :, 1:-1, 1:-1, 1:-1] - mean) * 0.7 + mean
        vox = vox / vox.sum(1, keepdim=True)
    scale_kernel = np.full(3, 1.0 / scale, dtype=np.float32)
    scale_kernel[2] *= ratio
    return {
        "lbs_voxel": vox[0].numpy().astype(np.float32),
        "offset_kernel": (-offset).astype(np.float32),
        "scale_kernel": scale_kernel,
    }


class SnarfSetup:
    """Subject-level state + per-frame bone transforms (host)."""

    def __init__(self, body: SyntheticBody | None = None, resolution: int = 128, engine=None):
        """``engine`` (a RenderEngine): voxelise on the device (ia_voxelize_lbs: 0.1 s instead of ~5 s of cKDTree queries on
        the host at resolution 128; the same recipe, held to the same reference golden); ``lbs_voxel`` is then a CUDA
        tensor."""
        self.body = body if body is not None else SyntheticBody()
        cano = self.body(body_pose=a_pose())
        self.tfs_inv_t = np.linalg.inv(cano["A"][0].astype(np.float64))
        if engine is not None:
            self.lbs_voxel, self.offset_kernel, self.scale_kernel = engine.voxelize_lbs(cano["vertices"][0],
                                                                                       self.body.lbs_weights, resolution)
        else:
            vox = voxelize_lbs_weights(cano["vertices"][0], self.body.lbs_weights, resolution)
            self.lbs_voxel = vox["lbs_voxel"]
            self.offset_kernel = vox["offset_kernel"]
            self.scale_kernel = vox["scale_kernel"]
        self.bbox = get_bbox_from_verts(cano["vertices"][0])  # canonical bbox -> field normalisation
        self.resolution = resolution

    def frame(self, body_pose, global_orient, transl):
        """Per-frame quantities of ``prepare_deformer`` (snarf_deformer.py:81-126)."""
        out = self.body(body_pose=body_pose, global_orient=global_orient, transl=transl)
        A = out["A"][0].astype(np.float64)
        s2w = A[0]
        w2s = np.linalg.inv(s2w)
        tfs = (w2s[None] @ A @ self.tfs_inv_t).astype(np.float32)
        verts = out["vertices"][0].astype(np.float64) @ w2s[:3, :3].T + w2s[:3, 3]
        return {
            "tfs": tfs,                                   # [24,4,4]
            "w2s": w2s.astype(np.float32),                # [4,4]
            "vertices": verts.astype(np.float32),         # root frame
            "deformed_bbox": get_bbox_from_verts(verts).reshape(-1),  # aabb[6] of the test occupancy grid
        }
