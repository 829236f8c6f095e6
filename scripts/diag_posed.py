#!/usr/bin/env python
"""Per-buffer / per-pixel differences of the AIST frame 0 parity case (tests/test_gpu_render.py::test_posed_frame_parity_spp16)."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_gpu_render as T
from e2e_cases import *  # noqa


def main():
    import conftest
    scene = conftest.Scene()
    fr, R, e, tabs, rays = T._setup(scene, 0, 16, 48)
    e.set_occupancy(fr["deformed_bbox"], R.binaries)
    ref = R.forward(rays, seed=0)
    got = e.render(rays.cuda(), seed=0)
    for k in ("comp_rgb", "comp_normal", "comp_albedo", "opacity", "depth", "comp_rgb_phys", "comp_demod_phys"):
        a, b = got[k].cpu().float(), ref[k].float()
        d = (a - b).abs().reshape(a.shape[0], -1).max(1).values
        top = torch.topk(d, 5)
        print(f"{k:18s} rel_l2 {T.rel_l2(got[k], ref[k]):.3e}  worst px {top.indices.tolist()} {[f'{v:.2e}' for v in top.values.tolist()]}")


main()
