#!/usr/bin/env python
"""The training-mode ops on a few hundred points (for compute-sanitizer memcheck / racecheck): fused query forward / backward,
shading networks backward (shared-memory staging + per-thread accumulators), compositing forward / backward with a gradient on
the weights, the BRDF combine and the environment backward.  Prints one line per op."""
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
bp, go, tr = syn.load_pose(0)
fr = snarf.frame(bp, go, tr)
e.set_pose(fr["tfs"], fr["w2s"])
g = torch.Generator().manual_seed(0)
n = 777                                                   # not a multiple of the 32 points a CTA stages
bb = torch.as_tensor(fr["deformed_bbox"])
c, h = (bb[:3] + bb[3:]) / 2, (bb[3:] - bb[:3]) / 2
xd = c + (torch.rand(n, 3, generator=g) * 2 - 1) * h * 0.45
fwd = e.op_query_train(xd)
gq = e.op_query_backward(fwd, torch.randn(n, 13, generator=g))
print("query_train / backward: valid", int(fwd["valid"].sum()), "|g_tfs|", float(gq["tfs"].abs().sum()))
v = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
nw = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
gs = e.op_shade_fields_backward(fwd["x_c"], fwd["feature"], v, nw, torch.randn(n, 3, generator=g), torch.randn(n, 5, generator=g))
print("shade_fields_backward: |g_w1|", float(gs["rad"]["w1"].abs().sum()), float(gs["mat"]["w3"].abs().sum()))
counts = torch.randint(0, 30, (60,), generator=g)
pi = torch.stack([torch.cumsum(counts, 0) - counts, counts], 1).int()
m = int(counts.sum())
sdf, dists, vals = (torch.rand(m, generator=g) - 0.35) * 0.08, torch.full((m,), 0.01), torch.randn(m, 12, generator=g)
w, comp, op = e.op_volrend(pi, sdf, dists, vals, 0.01)
gv = e.op_volrend_backward(pi, sdf, dists, vals, 0.01, torch.randn(60, 12, generator=g), torch.randn(60, generator=g),
                           torch.randn(m, generator=g))
print("volrend / backward: opacity", float(op.mean()), "|g_sdf|", float(gv[0].abs().sum()))
wi = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
nn_ = torch.nn.functional.normalize(wi + 0.7 * torch.randn(n, 3, generator=g), dim=-1)
wo = torch.nn.functional.normalize(nn_ + 0.9 * torch.randn(n, 3, generator=g), dim=-1)
a = (wi, nn_, wo, torch.rand(n, generator=g) * 0.9 + 0.05, torch.rand(n, 3, generator=g), torch.rand(n, generator=g),
     torch.rand(n, 3, generator=g), torch.full((n,), 4 * np.pi))
Lo = e.op_pbr_shade(*a)[0]
gp = e.op_pbr_shade_backward(*a, torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g), None)
print("pbr_shade / backward: mean Lo", float(Lo.mean()), "|g_rough|", float(gp["rough"].abs().sum()))
e.set_light_uniform(torch.rand(8, 16, 3, generator=g), 2, 4)
ge = e.op_env_backward(wo, torch.randn(n, 3, generator=g), (8, 16))
torch.cuda.synchronize()
print("env_backward: |g_env|", float(ge.abs().sum()))
