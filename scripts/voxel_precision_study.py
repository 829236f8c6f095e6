#!/usr/bin/env python
"""CPU study (oracle only): what would lower-precision voxel_J storage do to the rendered buffers?

The Broyden phase is bound by the BYTES it gathers: 8 corners x 48 B per fetch.  Candidates:
  fp16      : all 12 entries in fp16 (24 B / voxel) -- SURVEY 8d's format
  centre16  : y_c = R_c c_c + t_c (deformed voxel centre) in fp32 + R_c in fp16 = 30 B -> one 32-byte sector per corner;
              the kernel then evaluates  sum_c w_c (y_c + R_c (x - c_c))  -- algebraically the reference's  J(x) [x;1]
Each variant perturbs the field the root finder sees; this script renders the E2E golden cases with the oracle under each
and reports the relative L2 against the fp32 oracle (north_star tolerance: 1e-3).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import e2e_cases as E2E
from conftest import Scene
from oracle import deformer as odef
from oracle.render import OracleRenderer


def voxel_centres(R):
    """canonical position of every voxel centre [D,H,W,3] (inverse of g = scale * (x + offset))."""
    D, H, W = R.lbs_voxel.shape[1:]
    zs, ys, xs = torch.linspace(-1, 1, D), torch.linspace(-1, 1, H), torch.linspace(-1, 1, W)
    Z, Y, X = torch.meshgrid(zs, ys, xs, indexing="ij")
    g = torch.stack([X, Y, Z], -1)
    return g / R.scale - R.offset


def perturb(R, variant):
    J = odef.precompute(R.lbs_voxel, R.tfs)                      # [12,D,H,W] fp32
    if variant == "fp32":
        return J
    if variant == "fp16":
        return J.half().float()
    if variant == "centre16":
        D, H, W = J.shape[1:]
        M = J.reshape(3, 4, D, H, W)
        Rm, t = M[:, :3], M[:, 3]
        c = voxel_centres(R).permute(3, 0, 1, 2)                 # [3,D,H,W]
        y = torch.einsum("abdhw,bdhw->adhw", Rm, c) + t          # fp32
        R16 = Rm.half().float()
        # equivalent 3x4 the (unchanged) oracle sampler can use: t' = y - R16 c  (evaluated in fp64 to isolate storage error)
        t2 = (y.double() - torch.einsum("abdhw,bdhw->adhw", R16.double(), c.double())).float()
        return torch.cat([R16, t2[:, None]], 1).reshape(12, D, H, W)
    raise ValueError(variant)


def main():
    sc = Scene()
    gold = E2E.load()
    env = sc.syn.load_envmap()
    for case in E2E.CASES[:5]:
        name, frame, side, spp, mode, gi, emit, *extra = case
        fr = sc.frame(frame)
        outs = {}
        for variant in ("fp32", "fp16", "centre16"):
            R = OracleRenderer(sc.fields, sc.snarf.lbs_voxel, sc.snarf.offset_kernel, sc.snarf.scale_kernel, samples_per_pixel=spp,
                               global_illumination=gi, grid_res=E2E.GRID_RES, render_mode=mode, add_emitter=emit)
            R.set_pose(fr["tfs"], fr["w2s"])
            R.voxel_J = perturb(R, variant)
            R.binaries = E2E.grid(gold, frame)
            R.grid_aabb = torch.as_tensor(fr["deformed_bbox"], dtype=torch.float32)
            tabs = sc.syn.random_tables(spp, E2E.GRID_RES, seed=0)
            R.set_light(env, tabs["u1"], tabs["u2"])
            outs[variant] = R.forward(torch.from_numpy(sc.syn.make_rays(side, side, fr["transl"])), seed=0)
        for variant in ("fp16", "centre16"):
            print(f"{name:16s} {variant:9s} " + "  ".join(
                f"{k} {E2E.rel_l2(outs[variant][k], outs['fp32'][k]):.1e}" for k in ("comp_rgb_phys", "comp_albedo", "comp_normal", "opacity", "depth")))


if __name__ == "__main__":
    main()
