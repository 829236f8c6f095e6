"""GPU tests of the model seam: IntrinsicAvatarModel built from the reference's own ``config.model`` node through a
models.register / models.make registry (models/__init__.py:1-14), driven with a ``preprocess_data``-shaped batch
(systems/intrinsic_avatar.py:84-158), as a sub-module of a system-like parent (checkpoint round trip), and through the
render_image / render_image_relight conveniences."""
import copy
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# ---- the reference's registry, restated (models/__init__.py:1-14)
_models = {}


def register(name):
    def decorator(cls):
        _models[name] = cls
        return cls
    return decorator


def make(name, config):
    return _models[name](config)


def _reference_config(tmp_path, spp=8):
    """configs/config.yaml's model node (tests/golden/reference_model_config.json, scripts/make_golden.py config) pointed at
    an SMPL-shaped model file: the random model of reference_vectors_smpl.npz (the licensed SMPL data is not shipped)."""
    cfg = json.load(open(os.path.join(GOLD, "reference_model_config.json")))
    z = np.load(os.path.join(GOLD, "reference_vectors_smpl.npz"))
    d = tmp_path / "smpl"
    d.mkdir(exist_ok=True)
    np.savez(d / "SMPL_MALE.npz", v_template=z["smpl_v_template"], shapedirs=z["smpl_shapedirs"], posedirs=z["smpl_posedirs"],
             J_regressor=z["smpl_J_regressor"], weights=z["smpl_weights"])
    cfg["deformer"]["rigid_deformer"]["model_path"] = str(d)
    cfg["samples_per_pixel"] = spp
    return cfg, z["smpl_betas"].astype(np.float32)


def _batch(scene, betas, H, frame=0, hdri=True):
    """What preprocess_data hands to the model (systems/intrinsic_avatar.py:84-116), on the GPU with leading batch dim 1."""
    bp, go, tr = scene.syn.load_pose(frame)
    b = {"rays": torch.from_numpy(scene.syn.make_rays(H, H, tr)).cuda(), "betas": torch.from_numpy(betas).cuda(),
         "body_pose": torch.from_numpy(bp[None]).cuda(), "global_orient": torch.from_numpy(go[None]).cuda(),
         "transl": torch.from_numpy(tr[None]).cuda(), "index": frame}
    if hdri:
        b["hdri"] = torch.from_numpy(scene.syn.load_envmap_full()).cuda()
    return b


@pytest.fixture(scope="module")
def built(scene, tmp_path_factory):
    from intrinsicavatar_b200.model import IntrinsicAvatarModel
    register("intrinsic-avatar")(IntrinsicAvatarModel)
    cfg, betas = _reference_config(tmp_path_factory.mktemp("seam"))
    assert cfg["name"] == "intrinsic-avatar"
    model = make(cfg["name"], cfg)                      # systems/base.py:22: models.make(config.model.name, config.model)
    model.train(False)
    model.update_step(250, 25000)
    return model, cfg, betas


def test_registry_builds_the_subject_from_the_config(scene, built):
    model, cfg, betas = built
    from intrinsicavatar_b200.body import SMPLBody
    assert model.setup_snarf is None                    # lazy: the subject needs batch["betas"] (snarf_deformer.py:89-91)
    H = 24
    b = _batch(scene, betas, H)
    model.background_color = torch.ones(3, device="cuda")          # systems/base.py:112-119 sets it on the module
    model.prepare(b)
    assert isinstance(model.setup_snarf.body, SMPLBody) and np.allclose(model.setup_snarf.body.betas, betas.reshape(-1))
    assert model.setup_snarf.lbs_voxel.shape == (24, 32, 128, 128)   # deformer_config.resolution = 128
    out = model(b["rays"])
    n = H * H
    for k in ("comp_rgb", "comp_normal", "comp_rgb_phys", "comp_albedo", "comp_rgb_phys_full"):
        assert out[k].shape == (n, 3) and not out[k].is_cuda and torch.isfinite(out[k]).all(), k
    assert out["opacity"].shape == (n, 1) and float(out["opacity"].max()) > 0.5       # the body is hit
    # add_emitter (README.md relight command): a ray that misses the body shows the envmap, not the background colour
    miss = out["opacity"][:, 0] == 0
    assert miss.any() and not torch.allclose(out["comp_rgb_phys"][miss], torch.ones(3))
    with pytest.raises(ValueError):                     # optimize_betas = false: the shape is fixed at initialisation
        model.prepare({**b, "betas": b["betas"] + 0.5})


def test_resample_light_false_keeps_the_first_light(scene, built):
    """resample_light = false (the README's animation command): envmap, pdf and light directions are drawn ONCE
    (models/intrinsic_avatar.py:291-301); a later frame's hdri is not looked at."""
    model, cfg, betas = built
    assert cfg["resample_light"] is False
    b = _batch(scene, betas, 16)
    # the same occupancy jitter for both frames: a voxel that flips between two random draws would change the image
    jitter = torch.rand(int(model.config["occ_resolution"]) ** 3, 3, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    model.prepare(b, jitter=jitter)
    a = model(b["rays"])["comp_rgb_phys"].clone()
    dark = {**b, "hdri": b["hdri"] * 0.0}
    model.prepare(dark, jitter=jitter)
    c = model(b["rays"])["comp_rgb_phys"]
    assert float((a - c).abs().max()) < 1e-4 * float(a.abs().max())      # (accumulation order into a pixel is not fixed)


def test_checkpoint_round_trip_as_a_submodule(scene, built):
    """The reference keeps the model as ``system.model`` and loads ``ckpt['state_dict']`` into the system with strict=False
    (launch.py:110-124): parent.state_dict() must carry the render-path parameters under ``model.``, parent.load_state_dict
    must reach the device, tcnn-style flat fp16 hash-grid parameters must load, a wrong width must raise."""
    model, cfg, betas = built

    class System(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.model = m
    system = System(model)
    sd = {k: v.clone() for k, v in system.state_dict().items()}     # (state_dict() aliases the live parameters)
    keys = [k for k in sd if k.startswith("model.")]
    from intrinsicavatar_b200.weights import random_state_dict_shapes
    assert sorted(k[len("model."):] for k in keys) == sorted(random_state_dict_shapes())
    b = _batch(scene, betas, 16)
    model.prepare(b)
    before = model(b["rays"])
    new = {k: v.clone() for k, v in sd.items()}
    new["model.density.beta"] = sd["model.density.beta"] * 2.0
    new["model.geometry.encoding.encoding.encoding.params"] = sd["model.geometry.encoding.encoding.encoding.params"].half()
    new["loss.some_buffer"] = torch.zeros(3)            # a checkpoint carries keys outside the model
    res = system.load_state_dict(new, strict=False)
    assert "loss.some_buffer" in res.unexpected_keys and not res.missing_keys
    after = model(b["rays"])
    assert abs(float(after["beta"]) - 2.0 * float(before["beta"])) < 1e-3 * float(before["beta"]) + 2e-4
    assert float((after["opacity"] - before["opacity"]).abs().max()) > 1e-4           # the new beta reached the kernels
    r2 = model.load_state_dict({k: v for k, v in sd.items()})                          # 'model.'-prefixed keys are accepted
    assert not r2.missing_keys and not r2.unexpected_keys
    again = model(b["rays"])
    assert torch.allclose(again["opacity"], before["opacity"], atol=1e-6)
    bad = {"model.radiance.network.layers.0.weight": torch.zeros(64, 70)}
    with pytest.raises(Exception):
        model.load_state_dict(bad)


def test_render_image_surfaces(scene, built):
    model, cfg, betas = built
    H = 20
    b = _batch(scene, betas, H)
    jitter = torch.rand(64 ** 3, 3, 3, device="cuda")      # the occupancy grid's jitter, shared by both renders
    prim = model.render_image(b, b["rays"], H, H, jitter=jitter)
    assert prim["comp_rgb"].shape == (H, H, 3) and prim["opacity"].shape == (H, H, 1)
    full = model.render_image_relight(b, b["rays"], H, H, jitter=jitter)
    assert full["comp_rgb_phys"].shape == (H, H, 3)
    # the primary buffers do not depend on the shading stage
    assert torch.allclose(prim["comp_albedo"], full["comp_albedo"], atol=1e-6)
    assert torch.allclose(prim["depth"], full["depth"], atol=1e-5)
    hit = full["opacity"][..., 0] > 0.9
    assert hit.any() and float((prim["comp_rgb_phys"][hit] - full["comp_rgb_phys"][hit]).abs().max()) > 1e-3


def test_mismatching_nested_config_raises(built):
    from intrinsicavatar_b200.model import IntrinsicAvatarModel
    _, cfg, _ = built
    bad = copy.deepcopy(cfg)
    bad["geometry"]["xyz_encoding_config"]["n_levels"] = 12
    bad["material"]["mlp_network_config"]["n_neurons"] = 128
    with pytest.raises(ValueError) as e:
        IntrinsicAvatarModel(bad)
    assert "n_levels" in str(e.value) and "n_neurons" in str(e.value)
    nosmpl = copy.deepcopy(cfg)
    nosmpl["deformer"]["rigid_deformer"]["model_path"] = "/nonexistent/smpl"
    m = IntrinsicAvatarModel(nosmpl)
    with pytest.raises(FileNotFoundError):
        m.prepare({"betas": torch.zeros(1, 10), "body_pose": torch.zeros(1, 69), "global_orient": torch.zeros(1, 3),
                   "transl": torch.zeros(1, 3)})
