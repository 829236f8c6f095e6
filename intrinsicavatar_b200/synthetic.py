"""Synthetic inputs of BASELINE.json's configs: camera rays, SMPL-style
pose stream, environment map, explicit random tables.  Host-side, numpy.

  * camera: K = [[1000,0,256],[0,1000,256],[0,0,1]] at 512x512 (AnimationDataset's f=2000 at
    downscale 2, reference datasets/animation.py:72-87), c2w = I, rays as ``make_rays``
    (datasets/animation.py:19-27); near/far = |transl| -/+ 1 (:185-189).
  * pose: frames of ``load/animation/aist/poses.npz`` (first 32 frames shipped in data/), transl
    re-based to (0, 0.15, 5) (datasets/animation.py:127-131); ``neutral`` = zero pose.
  * light: ``data/city_1024x2048_f16.npz`` = the reference's hdri_images/city.hdr as AnimationDataset hands it
    to the model (cv2.imread ANYDEPTH|COLOR -> RGB -> cv2.resize(2048, 1024, INTER_AREA), datasets/animation.py:196-204),
    stored as fp16 -- exact: RGBE texels have 8-bit mantissas, the 52 736-nit sun fits fp16 (``load_envmap_full``; the
    bench and the high-spp goldens use it).  ``data/city_128x256_f16.npy`` = the same image area-downsampled 8x and
    bilinearly re-expanded to 1024x2048 (``load_envmap``; the stand-in of round 1, kept because the round-1 goldens
    were generated with it).
"""
from __future__ import annotations

import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def make_rays(H: int, W: int, transl) -> np.ndarray:
    """[H*W, 8] = o(3), d(3), near, far in world space."""
    f = 1000.0 * W / 512.0
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], dtype=np.float64)
    x, y = np.meshgrid(np.arange(W), np.arange(H), indexing="xy")
    xy = np.stack([x, y, np.ones_like(x)], axis=-1).reshape(-1, 3).astype(np.float32)
    d = xy @ np.linalg.inv(K).T
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    o = np.zeros_like(d)
    dist = float(np.sqrt(np.square(np.asarray(transl, np.float64)).sum()))
    near = np.full((len(d), 1), dist - 1.0)
    far = np.full((len(d), 1), dist + 1.0)
    return np.concatenate([o, d, near, far], axis=1).astype(np.float32)


N_FRAMES = 32   # frames of the AIST sequence shipped in data/


def load_pose(frame: int | None):
    """-> body_pose[69], global_orient[3], transl[3].  frame=None: neutral (zero) pose."""
    if frame is None:
        return np.zeros(69, np.float32), np.zeros(3, np.float32), np.array([0, 0.15, 5], np.float32)
    z = np.load(os.path.join(_DATA, "aist_poses_0_32.npz"))
    poses, trans = z["poses"], z["trans"]
    t = trans[frame] - trans[0] + np.array([0, 0.15, 5], np.float32)
    return poses[frame, 3:].astype(np.float32), poses[frame, :3].astype(np.float32), t.astype(np.float32)


def load_envmap(H: int = 1024, W: int = 2048) -> np.ndarray:
    small = np.load(os.path.join(_DATA, "city_128x256_f16.npy")).astype(np.float32)
    h, w = small.shape[:2]
    # bilinear, pixel-centre aligned, wrap in longitude / clamp in latitude
    ys = (np.arange(H) + 0.5) * h / H - 0.5
    xs = (np.arange(W) + 0.5) * w / W - 0.5
    y0 = np.floor(ys).astype(int)
    x0 = np.floor(xs).astype(int)
    wy = (ys - y0)[:, None, None].astype(np.float32)
    wx = (xs - x0)[None, :, None].astype(np.float32)
    y0c, y1c = np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)
    x0c, x1c = x0 % w, (x0 + 1) % w
    top = small[y0c][:, x0c] * (1 - wx) + small[y0c][:, x1c] * wx
    bot = small[y1c][:, x0c] * (1 - wx) + small[y1c][:, x1c] * wx
    return np.ascontiguousarray(top * (1 - wy) + bot * wy, dtype=np.float32)


def load_envmap_full() -> np.ndarray:
    """The reference's city.hdr at 1024x2048, fp32 RGB, exactly as ``datum["hdri"]`` (datasets/animation.py:196-204)."""
    return np.ascontiguousarray(np.load(os.path.join(_DATA, "city_1024x2048_f16.npz"))["env"].astype(np.float32))


def random_tables(spp: int, grid_res: int = 64, seed: int = 0):
    """Explicit randomness shared by product and oracle: occupancy jitter, light uniforms."""
    rng = np.random.RandomState(seed)
    return {
        "jitter": rng.rand(grid_res ** 3, 3, 3).astype(np.float32),
        "u1": rng.rand(spp).astype(np.float32),
        "u2": rng.rand(spp).astype(np.float32),
        "seed": seed,
    }
