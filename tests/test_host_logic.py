"""CPU (-m "not gpu"): host-side logic and the oracle's own invariants / edge cases (empty and ragged
inputs, sign-change snapping, permutations), at sizes that run in seconds."""
import os

import numpy as np
import pytest
import torch

from oracle import ops as oops
from oracle.pbr import kensler_permute, pixel_key


# ------------------------------------------------------------------------------ host logic ----
def test_hashgrid_layout_matches_tcnn_spec():
    """16 levels, base 16, scale 2^(8/15), T = 2^19."""
    from intrinsicavatar_b200.weights import hashgrid_layout
    L = hashgrid_layout()
    assert list(L["res"]) == [16, 24, 34, 49, 71, 102, 148, 213, 308, 446, 646, 934, 1352, 1956, 2831, 4096]
    assert list(L["size"][:5]) == [4096, 13824, 39304, 117656, 357912]
    assert all(s == 1 << 19 for s in L["size"][5:])
    assert L["total"] == 6299960 and list(L["offset"]) == list(np.cumsum([0] + list(L["size"][:-1])))


def test_fold_weight_norm_and_lipschitz():
    from intrinsicavatar_b200.weights import fold, random_state_dict
    sd = random_state_dict(0)
    w = fold(sd)
    g, v = sd["geometry.network.layers.0.weight_g"], sd["geometry.network.layers.0.weight_v"]
    assert torch.allclose(w["geo_w1"].norm(dim=1), g.reshape(-1), rtol=1e-5)      # ||W_row|| = g
    assert torch.allclose(F_normalize(w["geo_w1"]), F_normalize(v), atol=1e-6)
    for i, name in enumerate(("mat_w1", "mat_w2", "mat_w3")):
        c = torch.nn.functional.softplus(sd[f"material.network.lipshitz_bound_per_layer.{i}"]).reshape(-1)
        assert (w[name].abs().sum(1) <= c + 1e-5).all()                              # row L1 norm bounded by softplus(c)
    assert set(k.split(".")[0] for k in sd) == {"geometry", "radiance", "material", "density"}
    assert "geometry.encoding.encoding.encoding.params" in sd and "density.beta" in sd


def F_normalize(x):
    return torch.nn.functional.normalize(x, dim=1)


def test_synthetic_rays_and_pose_stream():
    from intrinsicavatar_b200 import synthetic as syn
    bp, go, tr = syn.load_pose(0)
    assert bp.shape == (69,) and go.shape == (3,) and np.allclose(tr, [0, 0.15, 5], atol=1e-6)
    rays = syn.make_rays(32, 32, tr)
    assert rays.shape == (1024, 8) and rays.dtype == np.float32
    assert np.allclose(np.linalg.norm(rays[:, 3:6], axis=1), 1, atol=1e-6)
    assert np.allclose(rays[:, 7] - rays[:, 6], 2.0, atol=1e-5)
    env = syn.load_envmap(64, 128)
    assert env.shape == (64, 128, 3) and env.min() >= 0 and env.max() > 50       # HDR sun survives the down-sampling
    t = syn.random_tables(8, 4, seed=3)
    assert t["jitter"].shape == (64, 3, 3) and t["u1"].shape == (8,)


def test_snarf_setup_frame(scene):
    s = scene.snarf
    assert s.lbs_voxel.shape == (24, 32, 128, 128)
    assert np.allclose(s.lbs_voxel.sum(0), 1, atol=1e-4)                          # skinning weights are convex
    fr = scene.frame(None)
    assert fr["tfs"].shape == (24, 4, 4) and fr["w2s"].shape == (4, 4)
    assert np.allclose(fr["tfs"][:, 3], [0, 0, 0, 1], atol=1e-6)
    R = fr["tfs"][:, :3, :3]
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-5)            # rigid bones
    bb = fr["deformed_bbox"]
    assert np.allclose(bb[3:] - bb[:3], (bb[3:] - bb[:3])[0])                     # cube bbox (get_bbox_from_smpl)


def test_model_surface_without_gpu():
    """The drop-in class refuses to construct without a device (no silent CPU path)."""
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    from intrinsicavatar_b200.model import IntrinsicAvatarModel
    with pytest.raises(RuntimeError):
        IntrinsicAvatarModel()


# ---------------------------------------------------------------------- oracle edge cases ----
def test_pack_unpack_empty_and_ragged():
    ri = torch.tensor([0, 0, 0, 2, 2, 5], dtype=torch.int64)
    packed = oops.pack_info(ri, 7)
    assert packed.tolist() == [[0, 3], [3, 0], [3, 2], [5, 0], [5, 0], [5, 1], [6, 0]]
    assert torch.equal(oops.unpack_info(packed, 6), ri)
    empty = oops.pack_info(torch.zeros(0, dtype=torch.int64), 4)
    assert empty[:, 1].sum() == 0 and oops.unpack_info(empty, 0).numel() == 0
    data = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    padded = oops.unpack_data(packed, data, 4)
    assert padded.shape == (7, 4, 2) and torch.equal(padded[2, :2], data[3:5]) and (padded[1] == 0).all()


def test_render_weight_and_accumulate():
    alphas = torch.tensor([0.5, 0.5, 1.0, 0.25, 0.0])
    packed = torch.tensor([[0, 3], [3, 0], [3, 2]], dtype=torch.int32)
    w, T = oops.render_weight_from_alpha(alphas, packed)
    assert torch.allclose(w, torch.tensor([0.5, 0.25, 0.25, 0.25, 0.0]))
    assert torch.allclose(T, torch.tensor([1.0, 0.5, 0.25, 1.0, 0.75]))
    ri = oops.unpack_info(packed, 5)
    acc = oops.accumulate_along_rays(w, None, ri, 3)
    assert torch.allclose(acc[:, 0], torch.tensor([1.0, 0.0, 0.25]))
    vals = torch.ones(5, 2) * torch.tensor([1.0, 2.0])
    assert torch.allclose(oops.accumulate_along_rays(w, vals, ri, 3)[0], torch.tensor([1.0, 2.0]))


@pytest.mark.parametrize("spp", [2, 5, 64])
def test_ray_resampling_invariants(spp):
    """cdf.cu:9-149: every hit ray gets exactly spp samples; fg counts + bg = spp; t is non-decreasing;
    after the first +/- SDF crossing all later samples repeat one t (zero-crossing snap)."""
    g = torch.Generator().manual_seed(spp)
    n_rays = 40
    steps = torch.randint(0, 12, (n_rays,), generator=g)
    steps[3] = 0
    steps[7] = 1
    base = torch.cumsum(steps, 0) - steps
    packed = torch.stack([base, steps], 1).int()
    n = int(steps.sum())
    starts = torch.zeros(n)
    ends = torch.zeros(n)
    sdfs = torch.zeros(n)
    alphas = torch.rand(n, generator=g) * 0.6
    for r in range(n_rays):
        b, s = int(base[r]), int(steps[r])
        t = 1.0 + torch.arange(s + 1) * 0.1
        starts[b:b + s], ends[b:b + s] = t[:-1], t[1:]
        sdfs[b:b + s] = torch.linspace(0.3, -0.3 if r % 2 == 0 else 0.05, s) if s else torch.zeros(0)
    w, _ = oops.render_weight_from_alpha(alphas, packed)
    rpi, ts, offs, idx, fg, bg, surf = oops.ray_resampling(packed, starts[:, None], ends[:, None], w, sdfs, spp)
    assert torch.equal(rpi[:, 1], (steps > 0).int() * spp)
    for r in range(n_rays):
        b, s = int(base[r]), int(steps[r])
        rb, rn = int(rpi[r, 0]), int(rpi[r, 1])
        if s == 0:
            assert rn == 0
            continue
        assert int(fg[b:b + s].sum()) + int(bg[r]) == spp
        is_fg = offs[rb:rb + rn, 0] < 1e4
        t_fg = ts[rb:rb + rn, 0][is_fg]
        assert (t_fg[1:] >= t_fg[:-1] - 1e-6).all()
        assert int(is_fg.sum()) == int(fg[b:b + s].sum())
        if int(surf[r]) >= 0 and is_fg.sum() > 1:
            # samples placed after the crossing interval all share the snapped t
            after = idx[rb:rb + rn][is_fg] > int(surf[r])
            if after.sum() > 1:
                assert float(t_fg[after].max() - t_fg[after].min()) == 0.0


def test_traverse_grid_lattice_and_miss():
    """nerfacc traverse_grids semantics: samples live on t = near + k*step, emitted iff their midpoint
    is in an occupied cell; a ray that misses the box yields nothing."""
    res = 8
    binaries = torch.zeros(res, res, res, dtype=torch.bool)
    binaries[:, 3:5, :] = True                                    # a slab in y
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1])
    o = torch.tensor([[0.05, -3.0, 0.1], [5.0, 5.0, 5.0]])
    d = torch.tensor([[0.0, 1.0, 0.0], [0.0, 1.0, 0.0]])
    step = 0.07
    tg = oops.traverse_grid(o, d, binaries, aabb, 0.0, 1e10, step)
    ts, te = tg["t_starts"], tg["t_ends"]
    assert tg["sample_packed_info"][1, 1] == 0 if "sample_packed_info" in tg else True
    assert len(ts) > 0 and torch.allclose(te - ts, torch.full_like(ts, step), atol=1e-6)
    k = ts / step
    assert torch.allclose(k, torch.round(k), atol=1e-3)           # global lattice from near = 0
    mid_y = -3.0 + (ts + te) / 2
    assert (mid_y > -0.25 - 1e-4).all() and (mid_y < 0.25 + 1e-4).all()
    assert abs(len(ts) - 0.5 / step) <= 1.5
    # edges: one run -> first edge is_left only, last is_right only, inner both
    il, ir = tg["is_left"], tg["is_right"]
    assert il[0] and not ir[0] and ir[-1] and not il[-1] and (il[1:-1] & ir[1:-1]).all()


@pytest.mark.parametrize("l", [2, 3, 4, 7, 256, 1000, 1024])
def test_light_permutation_is_a_permutation(l):
    for ray in (0, 1, 12345):
        key = pixel_key(7, np.full(l, ray, np.int64))
        p = kensler_permute(np.arange(l, dtype=np.uint64), l, key)
        assert sorted(p.tolist()) == list(range(l))
    a = kensler_permute(np.arange(l, dtype=np.uint64), l, pixel_key(7, np.full(l, 1, np.int64)))
    b = kensler_permute(np.arange(l, dtype=np.uint64), l, pixel_key(7, np.full(l, 2, np.int64)))
    if l >= 256:
        assert (a != b).mean() > 0.9                               # different pixels decorrelate


def test_oracle_tiny_frame(scene):
    """End-to-end oracle on a 12x12 / 2 spp frame: output contract of forward() (keys, shapes, background)."""
    fr = scene.frame(0)
    R = scene.oracle_renderer(spp=2, grid_res=16)
    R.set_pose(fr["tfs"], fr["w2s"])
    tabs = scene.syn.random_tables(2, 16, seed=0)
    R.build_occupancy(fr["deformed_bbox"], tabs["jitter"])
    R.set_light(scene.syn.load_envmap(), tabs["u1"], tabs["u2"])
    rays = torch.from_numpy(scene.syn.make_rays(12, 12, fr["transl"]))
    out = R.forward(rays, seed=0)
    for k, c in (("comp_rgb", 3), ("comp_normal", 3), ("opacity", 1), ("depth", 1), ("comp_rgb_phys", 3),
                 ("comp_demod_phys", 3), ("comp_albedo", 3), ("comp_metallic", 1), ("comp_roughness", 1),
                 ("comp_rgb_full", 3), ("comp_rgb_phys_full", 3)):
        assert out[k].shape == (144, c), k
    miss = out["opacity"][:, 0] == 0
    assert miss.any() and (~miss).any()
    assert (out["comp_rgb_phys"][miss] == 1).all()                # white background (systems/...:133-136)
    assert out["opacity"].max() <= 1 + 1e-5 and torch.isfinite(out["comp_rgb_phys"]).all()
    prim = R.forward(rays, seed=0, albedo_only=True)
    assert torch.equal(prim["comp_albedo"], out["comp_albedo"]) and (prim["comp_rgb_phys"] == 1).all()


def test_counter_rng_is_uniform_and_decorrelated():
    """oracle.pbr.rng_uniform (the stand-in for torch.rand in MultiLobe.sample / emitter.sample)."""
    from oracle.pbr import rng_uniform
    key = pixel_key(3, np.arange(4096, dtype=np.int64))
    j = np.arange(64, dtype=np.uint64)
    u = torch.stack([rng_uniform(key[:, None], j[None, :], d) for d in range(4)], -1)     # [4096, 64, 4]
    assert u.min() >= 0 and u.max() < 1
    assert abs(float(u.mean()) - 0.5) < 2e-3 and abs(float(u.var()) - 1 / 12) < 2e-3
    flat = u.reshape(-1, 4)
    c = torch.corrcoef(flat.t())
    assert float((c - torch.eye(4)).abs().max()) < 5e-3                                    # streams independent
    assert float(torch.corrcoef(torch.stack([u[:-1, :, 0].reshape(-1), u[1:, :, 0].reshape(-1)]))[0, 1].abs()) < 5e-3


@pytest.mark.parametrize("mode", ["uniform_light", "mats", "mis"])
def test_oracle_render_modes_tiny_frame(scene, mode):
    """The oracle's restatements of pbr_uniform_light_forward / pbr_mats_forward / pbr_mis_forward run end to end
    on a 12x12 / 8 spp frame: finite, non-negative radiance, white background on missed rays, visibility map."""
    from oracle.render import OracleRenderer
    fr = scene.frame(0)
    R = OracleRenderer(scene.fields, scene.snarf.lbs_voxel, scene.snarf.offset_kernel, scene.snarf.scale_kernel,
                       samples_per_pixel=8, grid_res=16, render_mode=mode)
    R.set_pose(fr["tfs"], fr["w2s"])
    tabs = scene.syn.random_tables(8, 16, seed=0)
    R.build_occupancy(fr["deformed_bbox"], tabs["jitter"])
    env = scene.syn.load_envmap()
    if mode == "uniform_light":
        R.set_light_uniform(env, 2, 4)
    else:
        R.set_light(env, tabs["u1"], tabs["u2"])
    rays = torch.from_numpy(scene.syn.make_rays(12, 12, fr["transl"]))
    out = R.forward(rays, seed=0)
    miss = out["opacity"][:, 0] == 0
    assert miss.any() and (~miss).any()
    assert (out["comp_rgb_phys"][miss] == 1).all()
    assert torch.isfinite(out["comp_rgb_phys"]).all() and out["comp_rgb_phys"].min() >= 0
    if mode == "uniform_light":
        assert out["visibility"].shape == (144, 1) and (out["visibility"][miss] == 0).all()
        assert 0 < float(out["visibility"].max()) <= 2 + 1e-5
    else:
        assert "visibility" not in out


def test_render_flag_constants_match_header():
    """capi.RENDER_* mirror the IA_RENDER_* macros of include/ia_b200.h."""
    import re
    from intrinsicavatar_b200 import capi
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ia_b200.h")).read()
    shift = int(re.search(r"#define IA_RENDER_MODE_SHIFT (\d+)", src).group(1))
    for name, macro in (("light", "LIGHT"), ("uniform_light", "UNIFORM_LIGHT"), ("mats", "MATS"), ("mis", "MIS")):
        v = int(re.search(rf"#define IA_RENDER_{macro} \((\d+) << IA_RENDER_MODE_SHIFT\)", src).group(1))
        assert capi.RENDER_MODES[name] == v << shift
    assert capi.RENDER_ADD_EMITTER == int(re.search(r"#define IA_RENDER_ADD_EMITTER (\d+)", src).group(1))
    assert capi.RENDER_GI == 2 and capi.RENDER_PRIMARY_ONLY == 1


def test_lightning_checkpoint_ingestion(tmp_path):
    """weights.load_lightning_checkpoint: ``model.``-prefixed reference keys in, unrelated keys dropped, shapes checked
    (launch.py:110-124 strict=False semantics)."""
    from intrinsicavatar_b200 import weights as W
    sd = W.random_state_dict(3)
    shapes = W.random_state_dict_shapes()
    assert set(shapes) == set(sd) and all(tuple(shapes[k].shape) == tuple(sd[k].shape) for k in sd)
    ck = {"state_dict": {**{"model." + k: v for k, v in sd.items()},
                         "model.pose_correction.delta": torch.zeros(3), "model.occupancy_grid.occs": torch.zeros(8),
                         "model.emitter.base": torch.zeros(4, 8, 3)},
          "epoch": 249, "global_step": 25000}
    # tcnn checkpoints hold the hash grid as fp16-sized flat params of the same element count: accept a reshape
    ck["state_dict"]["model.density.beta"] = sd["density.beta"].reshape(1)
    p = tmp_path / "last.ckpt"
    torch.save(ck, p)
    got = W.load_lightning_checkpoint(str(p))
    assert set(got) == set(sd)
    for k in sd:
        assert torch.equal(got[k], sd[k].float().reshape(got[k].shape)), k
    a, b = W.fold(got), W.fold(sd)
    assert all(torch.equal(a[k], b[k]) if torch.is_tensor(a[k]) else a[k] == b[k] for k in b)
    bad = dict(ck["state_dict"])
    del bad["model.radiance.network.layers.2.weight"]
    torch.save({"state_dict": bad}, p)
    with pytest.raises(KeyError, match="radiance.network.layers.2.weight"):
        W.load_lightning_checkpoint(str(p))
    bad = dict(ck["state_dict"])
    bad["model.geometry.network.layers.0.weight_v"] = torch.zeros(64, 27)
    torch.save({"state_dict": bad}, p)
    with pytest.raises(ValueError, match="geometry.network.layers.0.weight_v"):
        W.load_lightning_checkpoint(str(p))


def test_num_samples_is_reported_per_ray_chunk():
    """forward() reports num_samples like the reference's chunk_batch does: one entry per ray_chunk rays
    (models/utils.py:16-61)."""
    from intrinsicavatar_b200.model import IntrinsicAvatarModel as M
    assert M._chunk_sums(torch.ones(400, dtype=torch.int32), 4096).tolist() == [400]
    assert M._chunk_sums(torch.ones(8192, dtype=torch.int32), 4096).tolist() == [4096, 4096]
    assert M._chunk_sums(torch.arange(10, dtype=torch.int32), 4).tolist() == [6, 22, 17]
    assert M._chunk_sums(torch.zeros(0, dtype=torch.int32), 4096).tolist() == [0]
    assert M._chunk_sums(torch.ones(5, dtype=torch.int32), 4096).dtype == torch.int32


def test_nested_config_validation_against_the_reference_yaml():
    """model._check_nested: the reference's own config.model node (tests/golden/reference_model_config.json, made from
    configs/*.yaml by scripts/make_golden.py config) passes; contradicting what the kernels are compiled for is reported
    key by key (a mismatching Hydra config used to render wrong silently)."""
    import copy
    import json
    import os
    from intrinsicavatar_b200 import model as M
    cfg = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model_config.json")))
    errors = []
    M._check_nested(M._plain(cfg), M.EXPECTED_NESTED, "config", errors)
    assert errors == []
    bad = copy.deepcopy(cfg)
    bad["radiance"]["dir_encoding_config"]["degree"] = 3
    bad["geometry"]["xyz_encoding_config"]["per_level_scale"] = 1.5
    bad["deformer"]["rigid_deformer"]["deformer_config"]["use_j_inv"] = True
    bad["light"]["name"] = "envlight-SG"
    M._check_nested(M._plain(bad), M.EXPECTED_NESTED, "config", errors)
    assert len(errors) == 4 and any("degree" in e for e in errors) and any("use_j_inv" in e for e in errors)

    class Attr(dict):          # Hydra's DictConfig quacks like this
        __getattr__ = dict.__getitem__
    assert M._plain(Attr(a=Attr(b=[1, Attr(c=2)]))) == {"a": {"b": [1, {"c": 2}]}}


def test_png_and_exr_encoders_round_trip(tmp_path):
    """frames.png_bytes (cv2 and the zlib-only fallback) and frames.exr_bytes decode back to the pixels they were given."""
    import struct
    import zlib
    from intrinsicavatar_b200 import frames
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (9, 13, 3), dtype=np.uint8)
    rgba = rng.integers(0, 256, (9, 13, 4), dtype=np.uint8)

    def decode(b):        # minimal PNG reader: 8-bit RGB / RGBA, any filter type
        assert b[:8] == b"\x89PNG\r\n\x1a\n"
        pos, idat, hdr = 8, b"", None
        while pos < len(b):
            n, tag = struct.unpack(">I4s", b[pos:pos + 8])
            data = b[pos + 8:pos + 8 + n]
            assert struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xffffffff
            if tag == b"IHDR":
                hdr = struct.unpack(">IIBBBBB", data)
            elif tag == b"IDAT":
                idat += data
            pos += 12 + n
        W, H, depth, ctype = hdr[:4]
        C = {2: 3, 6: 4}[ctype]
        raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(H, 1 + W * C).astype(np.int64)
        out = np.zeros((H, W * C), np.int64)
        for y in range(H):
            f, line = raw[y, 0], raw[y, 1:]
            up = out[y - 1] if y else np.zeros(W * C, np.int64)
            for x in range(W * C):
                a = out[y, x - C] if x >= C else 0
                c = up[x - C] if x >= C else 0
                pred = {0: 0, 1: a, 2: up[x], 3: (a + up[x]) // 2}.get(int(f))
                if pred is None:
                    p = a + up[x] - c
                    pa, pb, pc = abs(p - a), abs(p - up[x]), abs(p - c)
                    pred = a if pa <= pb and pa <= pc else (up[x] if pb <= pc else c)
                out[y, x] = (line[x] + pred) & 255
        return out.reshape(H, W, C).astype(np.uint8)
    assert np.array_equal(decode(frames.png_bytes(img)), img)
    assert np.array_equal(decode(frames.png_bytes(rgba)), rgba)
    import builtins
    real_import = builtins.__import__

    def no_cv2(name, *a, **k):
        if name == "cv2":
            raise ImportError(name)
        return real_import(name, *a, **k)
    builtins.__import__ = no_cv2
    try:
        fallback = frames.png_bytes(img)
    finally:
        builtins.__import__ = real_import
    assert np.array_equal(decode(fallback), img)
    # EXR: header attributes, offset table, B / G / R float rows
    hdr = rng.random((5, 7, 3), dtype=np.float32) * 40
    b = frames.exr_bytes(hdr)
    assert struct.unpack("<II", b[:8]) == (20000630, 2)
    pos, attrs = 8, {}
    while b[pos] != 0:
        e = b.index(b"\0", pos); name = b[pos:e].decode(); pos = e + 1
        e = b.index(b"\0", pos); typ = b[pos:e].decode(); pos = e + 1
        n = struct.unpack("<i", b[pos:pos + 4])[0]
        attrs[name] = (typ, b[pos + 4:pos + 4 + n]); pos += 4 + n
    pos += 1
    assert attrs["compression"][1] == b"\0" and struct.unpack("<iiii", attrs["dataWindow"][1]) == (0, 0, 6, 4)
    offs = struct.unpack("<5Q", b[pos:pos + 40])
    for y, o in enumerate(offs):
        yy, n = struct.unpack("<ii", b[o:o + 8])
        assert yy == y and n == 3 * 7 * 4
        row = np.frombuffer(b[o + 8:o + 8 + n], np.float32).reshape(3, 7)
        assert np.array_equal(row[::-1].T, hdr[y])
    try:
        os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
        import cv2
        path = str(tmp_path / "e.exr")
        open(path, "wb").write(b)
        back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if back is not None:                     # OpenEXR support is optional in cv2 builds
            assert np.allclose(back[..., ::-1], hdr)
    except Exception:
        pass


def test_training_seam_routes_gradients_without_a_device():
    """intrinsicavatar_b200.train on a stand-in engine (no device here): every leaf of render_radiance / pbr_light receives the
    gradient the ops return, in its own shape, dtype and place -- the bone transforms get the query's and the canonical
    point's contribution in rows 0..2 only, roots are handed to the deformer backward in groups of 13, beta may be a tensor
    or a float.  (The ops themselves are held to autograd on the GPU, tests/test_gpu_ops.py.)"""
    from intrinsicavatar_b200.train import SHADE_PARAMS, pbr_light, render_radiance, volrend
    calls = {}

    class StandIn:
        dev = torch.device("cpu")

        def op_query_train(self, xd):
            n = xd.shape[0]
            return {"sdf": torch.randn(n) * 0.01, "x_c": torch.randn(n, 3), "valid": torch.rand(n) > 0.3, "grad": torch.randn(n, 3),
                    "feature": torch.randn(n, 13), "J_inv": torch.randn(n, 3, 3), "best": torch.zeros(n, dtype=torch.int32)}

        def op_query_backward(self, fwd, d_out):
            calls["query_d_out"] = d_out.clone()
            return {"hash": torch.full((100, 2), 2.0), "w1": torch.ones(64, 35), "b1": torch.ones(64), "w2": torch.ones(13, 64),
                    "b2": torch.ones(13), "tfs": torch.ones(24, 3, 4), "x": None}

        def op_deform_backward(self, xc, valid, J_inv, g_xc):
            assert xc.shape[1:] == (13, 3) and valid.shape[1:] == (13,) and J_inv.shape[1:] == (13, 3, 3) and g_xc.shape == xc.shape
            calls["deform_roots"] = xc.shape[0] * 13
            return torch.full((24, 3, 4), 0.5)

        def op_shade_fields(self, xc, f, v, n):
            return torch.rand(xc.shape[0], 3), torch.rand(xc.shape[0], 5)

        def op_shade_fields_backward(self, xc, f, v, n, d_rgb, d_mat):
            assert d_rgb.shape == (xc.shape[0], 3) and d_mat.shape == (xc.shape[0], 5)
            net = lambda i, o: {"w1": torch.ones(64, i), "b1": torch.ones(64), "w2": torch.ones(64, 64), "b2": torch.ones(64),
                                "w3": torch.ones(o, 64), "b3": torch.ones(o)}
            return {"hash": torch.ones(100, 2), "rad": net(67, 3), "mat": net(48, 5), "x": torch.ones_like(xc),
                    "feature": torch.ones_like(f), "normal": torch.ones_like(n)}

        def op_volrend(self, pi, sdf, dists, vals, beta):
            calls["beta"] = beta
            return torch.rand_like(sdf), torch.rand(pi.shape[0], vals.shape[1]), torch.rand(pi.shape[0])

        def op_volrend_backward(self, pi, sdf, dists, vals, beta, d_comp, d_op, d_w=None):
            assert d_comp.shape == (pi.shape[0], vals.shape[1])
            return torch.ones_like(sdf), torch.ones_like(vals), torch.full((1,), 3.0)

        def op_secondary(self, o, d, gi=False):
            return torch.full((o.shape[0],), 0.5), torch.zeros(o.shape[0], 3)

        def op_env(self, u=None, dirs_world=None):
            return None, None, torch.ones(dirs_world.shape[0], 3)

        def op_env_backward(self, dirs, d_em, shape):
            calls["env_d_em"] = d_em.clone()
            return torch.ones(shape[0], shape[1], 3)

        def op_pbr_shade(self, *a):
            m = a[0].shape[0]
            return torch.rand(m, 3), torch.rand(m, 3), torch.rand(m, 3)

        def op_pbr_shade_backward(self, wi, n, wo, rough, albedo, metal, Li, inv_pdf, d_Lo, d_Ld=None, d_Ls=None):
            m = wi.shape[0]
            return {"normal": torch.ones(m, 3), "rough": torch.ones(m), "albedo": torch.ones(m, 3), "metal": torch.ones(m),
                    "Li": torch.full((m, 3), 2.0)}

    shapes = {"geo_hash": (100, 2), "geo_w1": (64, 35), "geo_b1": (64,), "geo_w2": (13, 64), "geo_b2": (13,), "rad_hash": (200,),
              "rad_w1": (64, 67), "rad_b1": (64,), "rad_w2": (64, 64), "rad_b2": (64,), "rad_w3": (3, 64), "rad_b3": (3,),
              "mat_w1": (64, 48), "mat_b1": (64,), "mat_w2": (64, 64), "mat_b2": (64,), "mat_w3": (5, 64), "mat_b3": (5,)}
    assert set(SHADE_PARAMS) < set(shapes)
    P = {k: torch.zeros(v, requires_grad=True) for k, v in shapes.items()}
    tfs = torch.eye(4).repeat(24, 1, 1).requires_grad_(True)
    beta = torch.tensor(0.01, requires_grad=True)
    counts = torch.tensor([0, 3, 5, 1])
    pi = torch.stack([torch.cumsum(counts, 0) - counts, counts], 1).int()
    e = StandIn()
    out = render_radiance(e, P, tfs, torch.eye(4), torch.randn(4, 3), torch.randn(4, 3), pi, torch.rand(9), torch.rand(9) + 1, beta)
    assert out["comp_rgb"].shape == (4, 3) and out["comp_mats"].shape == (4, 5) and out["depth"].shape == (4,)
    assert out["weights"].requires_grad and not out["valid"].requires_grad and not out["normal_smpl"].requires_grad
    (out["comp_rgb"].sum() + out["depth"].sum() + out["opacity"].sum()).backward()
    for k, v in shapes.items():
        assert P[k].grad is not None and P[k].grad.shape == v, k
    assert float(P["geo_hash"].grad[0, 0]) == 2.0 and float(P["rad_hash"].grad[0]) == 1.0      # reshaped to the leaf's layout
    assert torch.equal(tfs.grad[:, :3, :], torch.full((24, 3, 4), 1.5)) and float(tfs.grad[:, 3, :].abs().max()) == 0.0
    assert calls["deform_roots"] == 13 and float(beta.grad) == 3.0 and calls["beta"] == pytest.approx(0.01)
    # d sdf from the compositing (ones) lands in channel 0 of the query's upstream gradient next to the shading's (ones)
    assert torch.equal(calls["query_d_out"][:, 0], torch.full((9,), 2.0)) and torch.equal(calls["query_d_out"][:, 1:], torch.ones(9, 12))
    # beta as a plain float: no gradient slot, same call
    w, comp, op = volrend(e, pi, torch.randn(9, requires_grad=True), torch.rand(9), torch.randn(9, 2), 0.02)
    assert calls["beta"] == 0.02 and comp.requires_grad
    # physically based branch
    m = 50
    leaves = {"n_raw": torch.randn(m, 3), "albedo": torch.rand(m, 3), "rough": torch.rand(m, 1), "metal": torch.rand(m), "env": torch.rand(4, 8, 3)}
    L = {k: v.requires_grad_(True) for k, v in leaves.items()}
    nrm = torch.nn.functional.normalize(L["n_raw"], dim=-1)
    light = torch.nn.functional.normalize(torch.randn(m, 3), dim=-1)
    Lo, Ld, Ls, vis = pbr_light(e, L["env"], torch.eye(4), nrm, L["albedo"], L["rough"], L["metal"], torch.randn(m, 3),
                                torch.randn(m, 3), light, torch.full((m,), 12.566))
    cm = (nrm.detach() * light).sum(-1) > 1e-6
    assert torch.equal(vis[:, 0], torch.where(cm, torch.ones(m), torch.zeros(m)))            # 2 x the stand-in's 0.5
    Lo.sum().backward()
    assert L["rough"].grad.shape == (m, 1) and L["metal"].grad.shape == (m,) and L["env"].grad.shape == (4, 8, 3)
    assert torch.equal(calls["env_d_em"], 2.0 * torch.where(cm, 0.5, 0.0)[:, None].expand(-1, 3))  # dLi x transmittance
    assert L["n_raw"].grad is not None and float(L["albedo"].grad.min()) == 1.0


def test_folded_leaves_keep_the_reference_parameters_in_the_graph():
    """train.folded_leaves: the effective weights equal weights.fold's and stay attached to the reference-keyed parameters --
    a gradient of ones on the effective first geometry layer arrives at weight_g as sum_i v_oi / |v_o| (weight normalisation,
    models/network_utils.py:201-244), beta = |b| + 1e-4 passes d|b|, and the hash tables pass through unchanged."""
    from intrinsicavatar_b200.train import folded_leaves
    from intrinsicavatar_b200.weights import fold, random_state_dict
    sd = random_state_dict(3)
    theta = {("model." + k): v.clone().float().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    f, f0 = folded_leaves(theta, "cpu"), fold(sd)
    for k, v in f0.items():
        if k == "beta":
            assert float(f[k].detach()) == pytest.approx(v)
        else:
            assert torch.equal(f[k].detach(), v), k
    (f["geo_w1"].sum() + 3.0 * f["beta"] + 2.0 * f["geo_hash"].sum() + f["mat_w2"].sum()).backward()
    v = theta["model.geometry.network.layers.0.weight_v"].detach()
    g = theta["model.geometry.network.layers.0.weight_g"]
    assert torch.allclose(g.grad, (v / v.norm(dim=1, keepdim=True)).sum(1, keepdim=True), atol=1e-5)
    assert torch.equal(theta["model.geometry.encoding.encoding.encoding.params"].grad,
                       torch.full_like(theta["model.geometry.encoding.encoding.encoding.params"], 2.0))
    b = theta["model.density.beta"]
    assert float(b.grad) == pytest.approx(3.0 * float(torch.sign(b.detach())))
    assert theta["model.material.network.layers.1.weight"].grad is not None
