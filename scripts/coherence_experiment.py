"""Experiment (run by hand on the GPU box, not a test): how much does the ORDER of the secondary rays matter to the
wavefront integrator?  Same rays through ia_op_secondary in several orders:
  * pixel-major, random light per consecutive sample (round-1 feed below 64 spp)
  * pixel-major, sorted by light direction within each pixel (round-1 feed at >= 64 spp)
  * light-major: for each light direction all pixels in raster order / in 8x8-tile (Morton) order -- bundles of PARALLEL
    rays from neighbouring surface points
  * fully shuffled
usage: python scripts/coherence_experiment.py [image side = 256] [spp = 64]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from intrinsicavatar_b200 import synthetic as syn
from intrinsicavatar_b200.engine import RenderEngine
from intrinsicavatar_b200.snarf import SnarfSetup
from intrinsicavatar_b200.weights import fold, hashgrid_layout, random_state_dict

torch.cuda.set_device(0)
snarf = SnarfSetup()
e = RenderEngine(0)
e.set_fields(fold(random_state_dict(0)), hashgrid_layout(), snarf.bbox)
e.set_lbs_voxels(snarf.lbs_voxel, snarf.offset_kernel, snarf.scale_kernel)
e.set_render_config([-1.25, -1.55, -1.25, 1.25, 0.95, 1.25])
bp, go, tr = syn.load_pose(0)
fr = snarf.frame(bp, go, tr)
e.set_pose(fr["tfs"], fr["w2s"])
SIDE = int(sys.argv[1]) if len(sys.argv) > 1 else 256
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tabs = syn.random_tables(spp, 64, seed=0)
e.build_occupancy(fr["deformed_bbox"], tabs["jitter"], 64)
dirs_w, em, pdf = e.set_light(syn.load_envmap_full(), tabs["u1"], tabs["u2"], return_tables=True)
w2s = torch.from_numpy(fr["w2s"]).cuda()
dirs = torch.nn.functional.normalize(dirs_w @ w2s[:3, :3].t(), dim=-1)         # SMPL-root frame
# shading points: primary hits of a 64x64 image
rays = torch.from_numpy(syn.make_rays(SIDE, SIDE, tr)).cuda()
out = e.render(rays, primary_only=True)
hit = out["opacity"][:, 0] > 0.9
o_w = rays[hit, :3] + rays[hit, 3:6] * out["depth"][hit]
o_s = o_w @ w2s[:3, :3].t() + w2s[:3, 3]
n = out["comp_normal"][hit] @ w2s[:3, :3].t()
P = o_s.shape[0]
print("pixels", P)
g = torch.Generator(device="cuda").manual_seed(0)
# per pixel: a random permutation of the light set
perm = torch.argsort(torch.rand(P, spp, device="cuda", generator=g), dim=1)
# direction sort key: Morton-ish on (lon, lat)
lon = torch.atan2(dirs[:, 0], dirs[:, 2]); lat = torch.asin(dirs[:, 1].clamp(-1, 1))
qx = ((lon / (2 * np.pi) + 0.5) * 1023).long().clamp(0, 1023); qy = ((lat / np.pi + 0.5) * 1023).long().clamp(0, 1023)
def part(v):
    v = (v | (v << 8)) & 0x00FF00FF; v = (v | (v << 4)) & 0x0F0F0F0F; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555
    return v
key = part(qx) | (part(qy) << 1)
rank = torch.argsort(torch.argsort(key))
def build(order):   # order: [P, spp] light index per consecutive sample
    d = dirs[order]                                  # [P, spp, 3]
    o = o_s[:, None, :].expand(P, spp, 3)
    cos = (n[:, None, :] * d).sum(-1) > 1e-6
    return o[cos].contiguous(), d[cos].contiguous()
orders = {"random (today)": perm, "sorted by direction": torch.gather(perm, 1, torch.argsort(rank[perm], dim=1))}
res = {}
for name, order in orders.items():
    o, d = build(order)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        T, _ = e.op_secondary(o, d)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[name] = (dt, float(T.mean()), o.shape[0])
    print(f"{name:24s} rays={o.shape[0]} time={dt*1e3:.1f} ms  mean T={float(T.mean()):.6f}")
# light-major: ray (pixel p, light k) for k outer, p inner -- every pixel uses every light exactly once, so the SET of rays
# is the one above
hit_idx = torch.nonzero(hit)[:, 0]
py, px = hit_idx // SIDE, hit_idx % SIDE
mort = part(px) | (part(py) << 1)
for name, porder in (("light-major, raster pixels", torch.arange(P, device="cuda")), ("light-major, Morton pixels", torch.argsort(mort))):
    for lname, lorder in (("", torch.arange(spp, device="cuda")), (" + lights in Morton order", torch.argsort(rank))):
        oo = o_s[porder][None, :, :].expand(spp, P, 3)
        dd = dirs[lorder][:, None, :].expand(spp, P, 3)
        cos = (n[porder][None, :, :] * dd).sum(-1) > 1e-6
        o, d = oo[cos].contiguous(), dd[cos].contiguous()
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            T, _ = e.op_secondary(o, d)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name + lname:48s} rays={o.shape[0]} time={dt*1e3:.1f} ms  mean T={float(T.mean()):.6f}")
o, d = build(perm)
idx = torch.randperm(o.shape[0], device="cuda", generator=g)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); T, _ = e.op_secondary(o[idx].contiguous(), d[idx].contiguous()); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{'all rays shuffled':24s} rays={o.shape[0]} time={dt*1e3:.1f} ms  mean T={float(T.mean()):.6f}")
