import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import Scene
from oracle import deformer as odef
sc = Scene()
fr = sc.frame(0)
R = sc.oracle_renderer(4)
R.set_pose(fr["tfs"], fr["w2s"])
e = sc.engine(); e.set_pose(fr["tfs"], fr["w2s"])
g = torch.Generator().manual_seed(2)
bb = torch.as_tensor(fr["deformed_bbox"]); c=(bb[:3]+bb[3:])/2; h=(bb[3:]-bb[:3])/2
xd = c + (torch.rand(20000,3,generator=g)*2-1)*h*0.45
got = e.op_query(xd, with_grad=True); ref = R._deform(xd, with_grad=True)
both = got["valid"].cpu() & ref["valid"]
same = (got["x_c"].cpu()-ref["x_c"]).abs().max(-1).values < 1e-4
m = both & same
print("n", m.sum().item(), "both", both.sum().item())
for k in ("sdf","grad_cano","grad","feature","x_c"):
    d = (got[k].cpu()[m]-ref[k][m]).abs()
    if d.dim()>1: d = d.max(-1).values
    print(k, "mean %.3e median %.3e p99 %.3e max %.3e" % (d.mean(), d.median(), d.quantile(0.99), d.max()))
# rotation check: grad = R grad_cano -> compare R implied
gc = ref["grad_cano"][m]; 
Rf = odef.forward_rotation(ref["x_c"][m], R.lbs_voxel, R.tfs, R.offset, R.scale)
gpu_gc = got["grad_cano"].cpu()[m]
pred = torch.einsum("bij,bj->bi", Rf, gpu_gc)
d = (pred - got["grad"].cpu()[m]).abs().max(-1).values
print("rotation consistency: mean %.3e p99 %.3e max %.3e" % (d.mean(), d.quantile(0.99), d.max()))
# where is grad_cano error large?
d = (got["grad_cano"].cpu()[m]-ref["grad_cano"][m]).abs().max(-1).values
idx = torch.argsort(-d)[:5]
print("worst grad_cano", d[idx], ref["grad_cano"][m][idx], got["grad_cano"].cpu()[m][idx])
