"""Timing of the training-mode ops on one B200: CUDA events on the launching stream, 3 warm-up + 10 timed
launches per op, inputs resident in HBM.  Product code only (no oracle).  Writes gpurun_out/train_ops.json.

    python scripts/bench_train.py [--points 1048576]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from intrinsicavatar_b200 import synthetic as syn                      # noqa: E402
from intrinsicavatar_b200.engine import RenderEngine                   # noqa: E402
from intrinsicavatar_b200.snarf import SnarfSetup                      # noqa: E402
from intrinsicavatar_b200.train import SHADE_PARAMS, render_radiance   # noqa: E402
from intrinsicavatar_b200.weights import fold, hashgrid_layout, random_state_dict  # noqa: E402


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 20)
    args = ap.parse_args()
    n = args.points
    snarf = SnarfSetup()
    folded, layout = fold(random_state_dict(0)), hashgrid_layout()
    e = RenderEngine()
    e.set_fields(folded, layout, snarf.bbox)
    e.set_lbs_voxels(snarf.lbs_voxel, snarf.offset_kernel, snarf.scale_kernel)
    e.set_render_config([-1.25, -1.55, -1.25, 1.25, 0.95, 1.25])
    bp, go, tr = syn.load_pose(0)
    fr = snarf.frame(bp, go, tr)
    e.set_pose(fr["tfs"], fr["w2s"])
    g = torch.Generator().manual_seed(0)
    bb = torch.as_tensor(fr["deformed_bbox"])
    c, h = (bb[:3] + bb[3:]) / 2, (bb[3:] - bb[:3]) / 2
    xd = (c + (torch.rand(n, 3, generator=g) * 2 - 1) * h * 0.45).cuda()        # the region around the body
    out = {"points": n, "device": torch.cuda.get_device_name(0), "ops": {}}

    def rec(name, ms, units, bytes_per_unit=None):
        r = {"ms": round(ms, 4), "units": units, "Munits_per_s": round(units / ms / 1e3, 2)}
        if bytes_per_unit:
            r["GB_per_s"] = round(units * bytes_per_unit / ms / 1e6, 1)
            r["bytes_per_unit"] = bytes_per_unit
        out["ops"][name] = r
        print(name, r, flush=True)

    fwd = e.op_query_train(xd)
    out["valid_fraction"] = float(fwd["valid"].float().mean())
    d_out = torch.randn(n, 13, device="cuda")
    rec("query_eval (ia_op_query, with_grad)", timed(lambda: e.op_query(xd, with_grad=True)), n)
    rec("query_train", timed(lambda: e.op_query_train(xd)), n)
    rec("query_backward", timed(lambda: e.op_query_backward(fwd, d_out)), n)
    xc, feat = fwd["x_c"], fwd["feature"]
    v = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=-1)
    nw = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=-1)
    d_rgb, d_mat = torch.randn(n, 3, device="cuda"), torch.randn(n, 5, device="cuda")
    rec("shade_fields", timed(lambda: e.op_shade_fields(xc, feat, v, nw)), n)
    rec("shade_fields_backward", timed(lambda: e.op_shade_fields_backward(xc, feat, v, nw, d_rgb, d_mat)), n)
    # compositing: rays of 32 samples
    spr = 32
    n_rays = n // spr
    pi = torch.stack([torch.arange(n_rays) * spr, torch.full((n_rays,), spr)], 1).int().cuda()
    sdf = ((torch.rand(n, device="cuda") - 0.35) * 0.08)
    dists = torch.full((n,), 0.01, device="cuda")
    vals = torch.randn(n, 12, device="cuda")
    d_comp, d_op = torch.randn(n_rays, 12, device="cuda"), torch.randn(n_rays, device="cuda")
    rec("volrend (12 channels)", timed(lambda: e.op_volrend(pi, sdf, dists, vals, 0.01)), n, 4 * (2 + 12 + 1))
    rec("volrend_backward", timed(lambda: e.op_volrend_backward(pi, sdf, dists, vals, 0.01, d_comp, d_op)), n, 4 * (2 + 12 + 1 + 12))
    # physically based combine: a streaming kernel (72 B in, 36 B out; backward 108 B in, 44 B out)
    m = 4 * n
    wi = torch.nn.functional.normalize(torch.randn(m, 3, device="cuda"), dim=-1)
    nn_ = torch.nn.functional.normalize(wi + 0.7 * torch.randn(m, 3, device="cuda"), dim=-1)
    wo = torch.nn.functional.normalize(nn_ + 0.9 * torch.randn(m, 3, device="cuda"), dim=-1)
    rough, metal = torch.rand(m, device="cuda") * 0.9 + 0.05, torch.rand(m, device="cuda")
    alb, Li = torch.rand(m, 3, device="cuda"), torch.rand(m, 3, device="cuda")
    ip = torch.full((m,), 4 * np.pi, device="cuda")
    ups = [torch.randn(m, 3, device="cuda") for _ in range(3)]
    rec("pbr_shade", timed(lambda: e.op_pbr_shade(wi, nn_, wo, rough, alb, metal, Li, ip)), m, 72 + 36)
    rec("pbr_shade_backward", timed(lambda: e.op_pbr_shade_backward(wi, nn_, wo, rough, alb, metal, Li, ip, *ups)), m, 108 + 44)
    # the radiance-field branch end to end: forward + backward through the three autograd nodes
    names = ("geo_hash", "geo_w1", "geo_b1", "geo_w2", "geo_b2") + SHADE_PARAMS
    P = {k: torch.as_tensor(folded[k], dtype=torch.float32).cuda().requires_grad_(True) for k in names}
    tfs = torch.as_tensor(fr["tfs"], dtype=torch.float32).cuda().requires_grad_(True)
    beta = torch.tensor(float(folded["beta"]), requires_grad=True)
    spr, n_rays = 16, n // 16
    pi16 = torch.stack([torch.arange(n_rays) * spr, torch.full((n_rays,), spr)], 1).int().cuda()
    p0 = xd[:n_rays]
    rd = torch.nn.functional.normalize(torch.randn(n_rays, 3, device="cuda"), dim=-1)
    ro = p0 - rd * (spr * 0.01 / 2)
    t0 = (torch.arange(spr, device="cuda") * 0.01).repeat(n_rays)
    t1 = t0 + 0.01

    def step():
        for t in list(P.values()) + [tfs, beta]:
            t.grad = None
        o = render_radiance(e, P, tfs, fr["w2s"], ro, rd, pi16, t0, t1, beta)
        (o["comp_rgb"].sum() + o["opacity"].sum() + o["comp_mats"].sum()).backward()

    rec("render_radiance fwd+bwd (16 samples per ray)", timed(step, reps=5), n_rays * spr)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "train_ops.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
