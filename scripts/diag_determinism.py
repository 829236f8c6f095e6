#!/usr/bin/env python
"""Render the smoke frame (primary-only and relit) twice and print the largest difference per output buffer."""
import sys
import torch
sys.path.insert(0, ".")
from intrinsicavatar_b200 import synthetic as syn
from intrinsicavatar_b200.engine import RenderEngine
from intrinsicavatar_b200.snarf import SnarfSetup
from intrinsicavatar_b200.weights import fold, hashgrid_layout, random_state_dict

snarf = SnarfSetup()
folded, layout = fold(random_state_dict(0)), hashgrid_layout()
bp, go, tr = syn.load_pose(0)
fr = snarf.frame(bp, go, tr)
spp, H, res = 4, 64, 32
tabs = syn.random_tables(spp, res, seed=0)
rays = torch.from_numpy(syn.make_rays(H, H, tr)).cuda()
e = RenderEngine(0)
e.set_fields(folded, layout, snarf.bbox)
e.set_lbs_voxels(snarf.lbs_voxel, snarf.offset_kernel, snarf.scale_kernel)
e.set_render_config([-1.25, -1.55, -1.25, 1.25, 0.95, 1.25])
e.set_pose(fr["tfs"], fr["w2s"])
e.build_occupancy(fr["deformed_bbox"], tabs["jitter"], res)
e.set_light(syn.load_envmap(), tabs["u1"], tabs["u2"])
outs = []
for it in range(3):
    o = e.render(rays, seed=0)
    torch.cuda.synchronize()
    outs.append({k: v.clone() for k, v in o.items() if torch.is_tensor(v)})
for k in outs[0]:
    d1 = float((outs[0][k].float() - outs[1][k].float()).abs().max())
    d2 = float((outs[1][k].float() - outs[2][k].float()).abs().max())
    print(f"{k:24s} max|r0-r1| = {d1:.3e}  max|r1-r2| = {d2:.3e}  max = {float(outs[0][k].float().abs().max()):.3e}")
