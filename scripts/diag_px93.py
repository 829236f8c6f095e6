#!/usr/bin/env python
"""Diagnostic (GPU box): secondary rays of one pixel (build/px93_rays.npz, written on the CPU side together with the
oracle's transmittance) through ia_op_secondary; prints where product and oracle disagree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import e2e_cases as E2E
from conftest import Scene

z = np.load(os.path.join(ROOT, "build", "px93_rays.npz"))
scene = Scene(); gold = E2E.load_hi()
fr = scene.frame(0)
e = scene.engine()
e.set_pose(fr["tfs"], fr["w2s"])
e.set_occupancy(fr["deformed_bbox"], E2E.grid(gold, 0))
o, d = torch.from_numpy(z["o"]), torch.from_numpy(z["d"])
T, rgb = e.op_secondary(o, d, gi=True)
T, rgb = T.cpu().numpy(), rgb.cpu().numpy()
q = e.op_query(o[:4])
print("sdf at P: product", q["sdf"].cpu().numpy(), "oracle", z["sdfP"])
dT = np.abs(T - z["T"][:, 0])
print("rays", len(T), "mean T product %.5f oracle %.5f" % (T.mean(), z["T"].mean()), "n |dT| > 1e-3:", int((dT > 1e-3).sum()))
bad = np.argsort(-dT)[:12]
for i in bad:
    print(i, "T product %.5f oracle %.5f" % (T[i], z["T"][i, 0]), "d", z["d"][i], "rgb product", rgb[i], "oracle", z["rgb"][i])
np.savez(os.path.join(ROOT, "gpurun_out", "diag_px93.npz"), T=T, rgb=rgb)
