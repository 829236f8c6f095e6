#!/usr/bin/env python
"""Tiny frame in every render mode (for compute-sanitizer): 20x20 rays, 8 spp, light / light+GI / uniform_light /
mats / mis+GI / add_emitter, plus the frame producer / consumer kernels.  Prints one line per mode."""
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
