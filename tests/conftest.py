import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


class Scene:
    """Synthetic subject + weights shared by oracle and product (built once per session)."""

    def __init__(self):
        from intrinsicavatar_b200 import synthetic as syn
        from intrinsicavatar_b200.snarf import SnarfSetup
        from intrinsicavatar_b200.weights import fold, hashgrid_layout, random_state_dict
        from oracle.fields import Fields

        self.syn = syn
        self.snarf = SnarfSetup()
        self.state_dict = random_state_dict(0)
        self.folded = fold(self.state_dict)
        self.layout = hashgrid_layout()
        self.fields = Fields(self.folded, self.layout, self.snarf.bbox)

    def frame(self, idx):
        bp, go, tr = self.syn.load_pose(idx)
        fr = self.snarf.frame(bp, go, tr)
        fr["transl"] = tr
        return fr

    def oracle_renderer(self, spp=4, gi=False, grid_res=64):
        from oracle.render import OracleRenderer
        return OracleRenderer(self.fields, self.snarf.lbs_voxel, self.snarf.offset_kernel, self.snarf.scale_kernel,
                              samples_per_pixel=spp, global_illumination=gi, grid_res=grid_res)

    def engine(self):
        from intrinsicavatar_b200.engine import RenderEngine
        e = RenderEngine()
        e.set_fields(self.folded, self.layout, self.snarf.bbox)
        e.set_lbs_voxels(self.snarf.lbs_voxel, self.snarf.offset_kernel, self.snarf.scale_kernel)
        e.set_render_config([-1.25, -1.55, -1.25, 1.25, 0.95, 1.25])
        return e


@pytest.fixture(scope="session")
def scene():
    return Scene()


@pytest.fixture(scope="session")
def posed(scene):
    """Frame 0 of the AIST sequence: oracle renderer state + engine with the ORACLE's occupancy grid,
    so that downstream comparisons are not affected by grid-cell flips."""
    fr = scene.frame(0)
    R = scene.oracle_renderer(spp=4)
    R.set_pose(fr["tfs"], fr["w2s"])
    tabs = scene.syn.random_tables(4, 64, seed=0)
    cache = os.path.join(ROOT, "tests", "golden", "_cache_occ_frame0.npy")
    if os.path.exists(cache):
        R.binaries = torch.from_numpy(np.load(cache))
        R.grid_aabb = torch.as_tensor(fr["deformed_bbox"], dtype=torch.float32)
    else:
        R.build_occupancy(fr["deformed_bbox"], tabs["jitter"])
        np.save(cache, R.binaries.numpy())
    return {"frame": fr, "oracle": R, "tabs": tabs}
