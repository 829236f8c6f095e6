#!/usr/bin/env python
"""Tiny frame in every render mode (for compute-sanitizer): 20x20 rays, 8 spp, light / light+GI / uniform_light /
mats / mis+GI / add_emitter, the non-default secondary-sampling switches, primary only, plus the frame producer / consumer
kernels, the tensor-core geometry op, the geometry backward and the device voxelisation.  Prints one line per case."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from intrinsicavatar_b200 import synthetic as syn
from intrinsicavatar_b200.engine import RenderEngine
from intrinsicavatar_b200.snarf import SnarfSetup
from intrinsicavatar_b200.weights import fold, hashgrid_layout, random_state_dict

snarf = SnarfSetup()
e = RenderEngine(0)
e.set_fields(fold(random_state_dict(0)), hashgrid_layout(), snarf.bbox)
e.set_lbs_voxels(snarf.lbs_voxel, snarf.offset_kernel, snarf.scale_kernel)
e.set_render_config([-1.25, -1.55, -1.25, 1.25, 0.95, 1.25])
bp, go, tr = syn.load_pose(0)
fr = snarf.frame(bp, go, tr)
e.set_pose(fr["tfs"], fr["w2s"])
tabs = syn.random_tables(8, 32, seed=0)
e.build_occupancy(fr["deformed_bbox"], tabs["jitter"], 32)
env = syn.load_envmap()
H = 20
f = 1000.0 * H / 512.0
dist = float(np.sqrt(np.square(tr).sum()))
rays = e.make_rays(np.array([[f, 0, H / 2.0], [0, f, H / 2.0], [0, 0, 1]]), H, H, dist - 1, dist + 1)
for mode, gi, emit in (("light", False, False), ("light", True, False), ("uniform_light", False, False), ("mats", False, False),
                       ("mis", True, False), ("light", False, True)):
    if mode == "uniform_light":
        e.set_light_uniform(env, 2, 4)
    else:
        e.set_light(env, tabs["u1"], tabs["u2"])
    o = e.render(rays, gi=gi, seed=0, render_mode=mode, add_emitter=emit)
    img = e.pack_rgb8(o["comp_rgb_phys_full"], bgr=True)
    torch.cuda.synchronize()
    print(mode, "gi" if gi else "", "emitter" if emit else "", "mean rgb_phys", float(o["comp_rgb_phys"].mean()),
          "uint8 mean", float(img.float().mean()), "rays", e.counters()["secondary_rays"])

# round 2: the secondary-sampling switches, primary only, image grid, tensor-core geometry op, geometry backward, voxelisation
e.set_light(env, tabs["u1"], tabs["u2"])
for imp, zc in ((False, True), (True, False), (False, False)):
    e.set_secondary_sampling(importance_sample=imp, zero_crossing_search=zc)
    o = e.render(rays, gi=True, seed=0, render_mode="light")
    torch.cuda.synchronize()
    print("switches", imp, zc, "mean rgb_phys", float(o["comp_rgb_phys"].mean()))
e.set_secondary_sampling(True, True)
o = e.render(rays, primary_only=True, seed=0)
grid = torch.empty(H, 3 * H, 3, dtype=torch.uint8, device="cuda")
e.pack_grid8(grid, 0, o["comp_rgb_full"].reshape(H, H, 3))
e.pack_grid8(grid, H, o["depth"].reshape(H, H), kind="grayscale", data_range=None)
e.pack_grid8(grid, 2 * H, o["comp_normal"].reshape(H, H, 3), data_range=(-1, 1))
bb = torch.as_tensor(snarf.bbox, dtype=torch.float32).reshape(2, 3)
xc = (bb[0] + torch.rand(1000, 3) * (bb[1] - bb[0])).cuda()
sd = e.op_geometry(xc)
gb = e.op_geometry_backward(xc, torch.randn(1000, 13).cuda())
torch.cuda.synchronize()
print("primary only / grid / geometry ops", float(grid.float().mean()), float(sd.mean()))
