#!/usr/bin/env python
"""Diagnostic (GPU box): render the high-spp golden frames with the loaded library, print the relative L2 of every buffer
against the reference's forward_, and store the buffers in gpurun_out/ for analysis on the CPU side.
usage: python scripts/diag_hi.py [tag]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import e2e_cases as E2E
from conftest import Scene

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
scene = Scene()
gold = E2E.load_hi()
dump = {}
for name, frame, side, spp, mode, gi, offset in E2E.HI_CASES:
    fr = scene.frame(frame)
    e = scene.engine()
    e.set_pose(fr["tfs"], fr["w2s"])
    e.set_occupancy(fr["deformed_bbox"], E2E.grid(gold, frame))
    tabs = scene.syn.random_tables(spp, E2E.GRID_RES, seed=0)
    e.set_light(scene.syn.load_envmap_full(), tabs["u1"], tabs["u2"])
    rays = E2E.hi_rays(scene.syn, fr["transl"], side).cuda()
    got = e.render(rays, gi=gi, seed=0, render_mode=mode, ray_index_base=offset)
    torch.cuda.synchronize()
    line = []
    for k in E2E.KEYS:
        r = torch.from_numpy(gold[f"{name}/{k}"])
        g = got[k].cpu()
        dump[f"{name}/{k}"] = g.numpy()
        d = (g - r).abs().max(-1).values
        line.append(f"{k}={E2E.rel_l2(g, r):.2e}(px{int(d.argmax())}:{float(d.max()):.1e})")
    dump[f"{name}/num_samples"] = got["num_samples"].cpu().numpy()
    print(tag, name, " ".join(line), "counters", {k: v for k, v in e.counters().items() if k != "primary"}, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"diag_hi_{tag}.npz"), **dump)
