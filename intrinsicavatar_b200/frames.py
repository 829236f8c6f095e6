"""Frame producer / consumer either side of the render path.

The reference builds every test frame on the host -- ``AnimationDataset.__getitem__`` (datasets/animation.py:163-206)
makes the ray tensors with numpy, reloads and resizes the HDRI with cv2 for *every* frame, a DataLoader worker
pickles ~35 MB per frame to the main process, Lightning copies it to the GPU -- and converts every output image to
uint8 on the host (utils/mixins.py:43-58).  Once a frame renders in under a second that host work dominates.
Here the per-frame inputs are produced on the device (``ia_make_rays``), the envmap is uploaded once, and the
images are clipped / scaled to uint8 on the device (``ia_pack_rgb8``) so a frame leaves as 1 byte per channel.

Same batch contract as the reference's ``preprocess_data`` (systems/intrinsic_avatar.py:84-116): ``rays [H*W,8]``,
``body_pose [1,69]``, ``global_orient [1,3]``, ``transl [1,3]``, ``betas``, ``hdri [1024,2048,3]``, ``index``.
"""
from __future__ import annotations

import numpy as np
import torch


class AnimationFrames:
    """Device-side stand-in for AnimationDataset(split="test") + preprocess_data.

    poses [F,72] (global_orient + body_pose), trans [F,3]: ``load/animation/<seq>/poses.npz``;
    K [3,3]: cameras.npz intrinsic already divided by ``downscale``; w2c [4,4] or [F,4,4]: cameras.npz extrinsic.
    """

    def __init__(self, engine, poses, trans, K, H, W, w2c=None, hdri=None, betas=None, near=None, far=None):
        self.engine = engine
        self.poses = np.asarray(poses, np.float32)
        trans = np.asarray(trans, np.float32)
        # datasets/animation.py:127-131: the sequence is re-based so that frame 0 stands at (0, 0.15, 5)
        self.trans = trans - trans[0] + np.array([0, 0.15, 5], np.float32)
        self.K, self.H, self.W = np.asarray(K, np.float64), int(H), int(W)
        self.w2c = None if w2c is None else np.asarray(w2c, np.float32)
        self.near, self.far = near, far
        self.betas = np.zeros((1, 10), np.float32) if betas is None else np.asarray(betas, np.float32).reshape(1, 10)
        # uploaded once; the reference re-reads and re-sizes the .hdr file for every frame (:191-201)
        self.hdri = None if hdri is None else torch.as_tensor(hdri, dtype=torch.float32).to(engine.dev).contiguous()
        self._rays = torch.empty(self.H * self.W, 8, device=engine.dev)

    def __len__(self):
        return len(self.poses)

    def __getitem__(self, idx):
        transl = self.trans[idx]
        if self.near is not None and self.far is not None:
            near, far = self.near, self.far
        else:
            dist = float(np.sqrt(np.square(transl).sum(-1)))      # distance from the camera to the mid-hip (:185-189)
            near, far = dist - 1, dist + 1
        w2c = None
        if self.w2c is not None:
            w2c = self.w2c[idx] if self.w2c.ndim == 3 else self.w2c
        rays = self.engine.make_rays(self.K, self.H, self.W, near, far, w2c=w2c, out=self._rays)
        batch = {
            "rays": rays,
            "betas": torch.from_numpy(self.betas[0]),
            "global_orient": torch.from_numpy(self.poses[idx, :3][None]),
            "body_pose": torch.from_numpy(self.poses[idx, 3:][None]),
            "transl": torch.from_numpy(transl[None]),
            "index": idx,
        }
        if w2c is not None:
            batch["w2c"] = torch.from_numpy(np.asarray(w2c))
        if self.hdri is not None:
            batch["hdri"] = self.hdri
        return batch

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def images_to_uint8(engine, out: dict, H: int, W: int, keys=("comp_rgb_full", "comp_rgb_phys_full", "comp_albedo_full"),
                    bgr=True) -> dict:
    """The validation/test image columns of the reference (systems/intrinsic_avatar.py:436-530 through
    SaverMixin.get_rgb_image_) as uint8 [H,W,3] host arrays ready for cv2.imwrite: conversion on the device, one
    3-byte-per-pixel copy per image."""
    res = {}
    for k in keys:
        img = out[k]
        if not img.is_cuda:
            img = img.to(engine.dev, non_blocking=True)
        res[k] = engine.pack_rgb8(img.reshape(-1, img.shape[-1]), (0.0, 1.0), bgr=bgr).reshape(H, W, -1).cpu().numpy()
    return res
