"""Drop-in model for the reference's ``IntrinsicAvatarModel`` on the render path.

Keeps the reference surface (models/intrinsic_avatar.py:166-305, 1653-1674; models/base.py):
  ``__init__(config)`` + ``setup()``, ``prepare(batch)``, ``forward(rays[N,8]) -> dict`` (CPU tensors in eval, same keys incl.
  ``*_bg`` / ``*_full`` and ``beta``), ``update_step(epoch, global_step)``, ``train()/eval()``, the externally set attributes
  ``background_color``, ``albedo_only``, ``albedo_align_ratio``, ``t_idx``, and the reference's PARAMETER TREE: the render-path
  parameters are registered under the reference's names (``geometry.network.layers.0.weight_v`` ...), so ``state_dict`` /
  ``load_state_dict`` behave like any ``nn.Module``'s, also when the model is a sub-module (the reference keeps it as
  ``system.model``; launch.py:110-124 loads checkpoints with ``strict=False``).
It adds ``render_image`` / ``render_image_relight`` conveniences named by BASELINE.json.

``config`` is the reference's ``config.model`` node (Hydra DictConfig or plain dict, configs/config.yaml:43-80 with its nested
geometry / radiance / material / density / deformer / light / scatterer nodes).  The subject is built the way
``SNARFDeformer`` does (models/deformers/snarf_deformer.py:37-104): the SMPL model named by
``config.deformer.rigid_deformer.{model_path, gender}``, initialised lazily from ``batch["betas"]`` on the first ``prepare``,
voxelised at ``deformer_config.resolution``.  A config WITHOUT a ``deformer`` node (the tests and the bench: the licensed SMPL
file is not shipped) or with ``subject: synthetic`` uses the procedural 24-joint body of ``body.SyntheticBody``.
Nested nodes are checked against what the kernels are compiled for (hash-grid layout, MLP widths, ...): a mismatch raises.

Only the eval render path is implemented (render_mode = light | uniform_light | mats | mis, with or without
global_illumination / add_emitter); ``train(True)`` raises -- the training-mode forward / backward lives below this seam,
in ``intrinsicavatar_b200.train`` (autograd nodes over the CUDA ops).
Every numeric step runs in libia_b200.so; this file is glue (pose -> 24 matrices, pointer passing).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import weights as W
from .engine import RenderEngine
from .snarf import SnarfSetup

DEFAULT_CONFIG = {
    # configs/config.yaml:43-80 + configs/dataset/animation/male-3-casual.yaml:10
    "name": "intrinsic-avatar-b200",
    "global_illumination": False,
    "render_mode": "light",
    "scene_aabb": [-1.25, -1.55, -1.25, 1.25, 0.95, 1.25],
    "samples_per_pixel": 1024,
    "num_samples_per_ray": 128,
    "num_samples_per_secondary_ray": 64,
    "secondary_near_plane": 0.0,
    "secondary_far_plane": 1.5,
    "secondary_importance_sample": True,
    "zero_crossing_search": True,
    "resample_light": True,
    "add_emitter": False,
    "grid_prune_occ_thre": 0.001,
    "ray_chunk": 4096,                # the kernels are persistent over all rays; only shapes num_samples (one entry per chunk)
    "secondary_shader_chunk": 160000, # accepted, unused
    "material_feature": "hybrid",
    "phys_kick_in_step": 10000,
    "importance_sample_kick_in_step": 1000,
    "occ_resolution": 64,
}

# What the kernels are compiled for (csrc/ia_types.cuh, weights.py).  A nested config node may omit a key (the reference's
# default applies) but may not contradict one.
_HASH_ENCODING = {"otype": "ProgressiveBandHashGrid", "n_levels": W.N_LEVELS, "n_features_per_level": W.N_FEAT,
                  "log2_hashmap_size": W.LOG2_T, "base_resolution": W.BASE_RES, "per_level_scale": W.PER_LEVEL_SCALE,
                  "interpolation": "Linear", "include_xyz": True}
EXPECTED_NESTED = {
    "geometry": {"name": "volume-sdf", "feature_dim": 13, "grad_type": "analytic", "xyz_encoding_config": _HASH_ENCODING,
                 "mlp_network_config": {"otype": "VanillaMLP", "output_activation": "none", "n_neurons": 64,
                                        "n_hidden_layers": 1, "weight_norm": True}},
    "radiance": {"name": "volume-ref-dir-radiance", "input_feature_dim": 16, "xyz_encoding_config": _HASH_ENCODING,
                 "dir_encoding_config": {"otype": "SphericalHarmonics", "degree": 4},
                 "mlp_network_config": {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none",
                                        "n_neurons": 64, "n_hidden_layers": 2},
                 "color_activation": "sigmoid"},
    "material": {"name": "volume-material", "n_output_dim": 5,
                 "mlp_network_config": {"otype": "LipshitzMLP", "activation": "ReLU", "output_activation": "none",
                                        "n_neurons": 64, "n_hidden_layers": 2},
                 "material_activation": "sigmoid"},
    "density": {"name": "learned-laplace-density"},
    "scatterer": {"name": "brdf-multi-lobe"},
    "light": {"name": "envlight-tensor", "xyz2lonlat_mode": None},
    "deformer": {"name": "snarf_deformer",
                 "rigid_deformer": {"name": "fast-snarf",
                                    "deformer_config": {"cano_pose": "A_pose", "use_j_inv": False, "optimize_betas": False}},
                 "non_rigid_deformer": {"name": "dummy_non_rigid_deformer"}},
}
_MATERIAL_AFFINE = ("albedo_scale", "albedo_bias", "roughness_scale", "roughness_bias", "metallic_scale", "metallic_bias")


def _plain(node):
    """Hydra DictConfig / attribute dict / dict -> plain nested dict (lists stay lists)."""
    if hasattr(node, "items"):
        return {k: _plain(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [_plain(v) for v in node]
    return node


def _check_nested(cfg, expected, path, errors):
    for k, want in expected.items():
        if k not in cfg:
            continue
        got = cfg[k]
        if isinstance(want, dict):
            if isinstance(got, dict):
                _check_nested(got, want, f"{path}.{k}", errors)
            else:
                errors.append(f"{path}.{k}: expected a config node, got {got!r}")
        elif isinstance(want, float):
            if not (isinstance(got, (int, float)) and math.isclose(float(got), want, rel_tol=1e-9)):
                errors.append(f"{path}.{k} = {got!r}, the kernels are built for {want!r}")
        elif got != want:
            errors.append(f"{path}.{k} = {got!r}, the kernels are built for {want!r}")


class _Node(torch.nn.Module):
    """Empty container: the reference's module tree only as far as its parameter NAMES go."""


class IntrinsicAvatarModel(torch.nn.Module):
    def __init__(self, config=None, body=None, device: int | None = None, seed: int = 0):
        super().__init__()
        user = _plain(config) if config is not None else {}
        self.config = dict(DEFAULT_CONFIG)
        self.config.update(user)
        cfg = self.config
        if cfg["render_mode"] not in ("light", "uniform_light", "mats", "mis"):
            # same failure as the reference's dispatch (models/intrinsic_avatar.py:1435-1438)
            raise NotImplementedError(f"Render mode {cfg['render_mode']} not supported.")
        if cfg["render_mode"] == "uniform_light":
            assert cfg["samples_per_pixel"] == 512  # models/intrinsic_avatar.py:1391 (16 x 32 stratified sphere)
        if cfg["material_feature"] not in W.MATERIAL_IN:
            raise ValueError(f"material_feature {cfg['material_feature']!r}: geometry | radiance | hybrid (models/intrinsic_avatar.py:1102-1113)")
        errors = []
        _check_nested(cfg, EXPECTED_NESTED, "config", errors)
        want_in = W.MATERIAL_IN[cfg["material_feature"]]
        if isinstance(cfg.get("material"), dict) and cfg["material"].get("input_feature_dim", want_in) != want_in:
            errors.append(f"config.material.input_feature_dim = {cfg['material']['input_feature_dim']!r}, material_feature = "
                          f"{cfg['material_feature']!r} needs {want_in}")
        if errors:
            raise ValueError("IntrinsicAvatarModel: this build of libia_b200 cannot render the configured model:\n  "
                             + "\n  ".join(errors))
        self._body_arg = body
        self._device_arg = device
        self.seed = seed
        self.background_color = torch.ones(3)
        self.albedo_only = False
        self.t_idx = 0.0
        self.enable_phys = True
        self.importance_sample = True
        self._frame = None
        self._light_key = None
        self._light_hdri = None
        self._betas = None
        self.setup_snarf = None
        self.engine = None
        self.setup()

    # ------------------------------------------------------------------- set-up ----
    def setup(self):
        """models/base.py: BaseModel.__init__ calls setup().  Creates the device context and the parameter tree (random
        initial values as ``weights.random_state_dict``; a checkpoint replaces them through ``load_state_dict``).  The
        subject (skinning-weight voxels, canonical bbox) follows on the first ``prepare`` -- it needs ``batch["betas"]`` --
        unless a body object was handed to the constructor."""
        self.engine = RenderEngine(self._device_arg)
        self.layout = W.hashgrid_layout()
        for key, v in W.material_state_dict_for(W.random_state_dict(self.seed), self.config["material_feature"]).items():
            self._register(key, v)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._upload_fields())
        if self._body_arg is not None:
            self._init_subject(self._body_arg)

    def _register(self, key: str, value: torch.Tensor):
        parts = key.split(".")
        node = self
        for p in parts[:-1]:
            if not hasattr(node, p):
                node.add_module(p, _Node())
            node = getattr(node, p)
        node.register_parameter(parts[-1], torch.nn.Parameter(value.detach().clone().float(), requires_grad=False))

    def _resolve_body(self, betas):
        """SNARFDeformer.__init__ (snarf_deformer.py:39): ``SMPL(config.model_path, gender=config.gender)``."""
        from .body import SMPLBody, SyntheticBody
        dcfg = self.config.get("deformer")
        subject = self.config.get("subject", "smpl" if dcfg else "synthetic")
        if subject == "synthetic":
            return SyntheticBody()
        rigid = (dcfg or {}).get("rigid_deformer", {})
        model_path, gender = rigid.get("model_path"), str(rigid.get("gender", "neutral"))
        if not model_path:
            raise ValueError("config.deformer.rigid_deformer.model_path is not set (or pass subject='synthetic')")
        cands = [model_path] if os.path.isfile(model_path) else [
            os.path.join(model_path, f"SMPL_{gender.upper()}{ext}") for ext in (".pkl", ".npz")]
        for c in cands:
            if os.path.isfile(c):
                return SMPLBody.from_file(c, betas=betas)
        raise FileNotFoundError(f"SMPL model not found (tried {cands}); the reference reads it the same way "
                                "(models/deformers/smplx/body_models.py:113-129).  Use subject='synthetic' for the "
                                "procedural test body.")

    def _init_subject(self, body):
        res = int(self.config.get("deformer", {}).get("rigid_deformer", {}).get("deformer_config", {}).get("resolution", 128))
        self.setup_snarf = SnarfSetup(body, resolution=res, engine=self.engine)     # voxelisation on the device
        self.engine.set_lbs_voxels(self.setup_snarf.lbs_voxel, self.setup_snarf.offset_kernel, self.setup_snarf.scale_kernel)
        self._upload_fields()

    # ---------------------------------------------------------------- parameters ----
    def load_state_dict(self, state_dict, strict: bool = False, assign: bool = False):
        """Reference-keyed state dict, optionally under a ``model.`` prefix (a Lightning checkpoint's ``state_dict``).  Keys
        outside the render path are reported as unexpected, not loaded; every accepted tensor must have the element count
        of the parameter it replaces (tiny-cuda-nn stores the hash grids as flat fp16).  Returns torch's
        ``_IncompatibleKeys(missing_keys, unexpected_keys)``; ``strict=True`` raises on either, like nn.Module."""
        own = dict(self.named_parameters())
        sd, unexpected = {}, []
        for k, v in state_dict.items():
            k2 = k[len("model."):] if k.startswith("model.") else k
            if k2 not in own:
                unexpected.append(k)
                continue
            v = torch.as_tensor(v)
            if v.numel() != own[k2].numel():
                raise ValueError(f"load_state_dict: {k} has shape {tuple(v.shape)}, the render path is built for "
                                 f"{tuple(own[k2].shape)}")
            sd[k2] = v.detach().float().reshape(own[k2].shape)
        res = super().load_state_dict(sd, strict=False)
        missing = list(res.missing_keys)
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict(strict=True): missing {missing[:4]}, unexpected {unexpected[:4]}")
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def load_checkpoint(self, path: str):
        """Ingest a Lightning checkpoint of the reference (launch.py:110-124): the render-path parameters under
        ``model.`` are taken, everything else is ignored (strict=False semantics)."""
        return self.load_state_dict(W.load_lightning_checkpoint(path, material_feature=self.config["material_feature"]))

    def _upload_fields(self):
        if self.setup_snarf is None:
            return                                       # the canonical bbox is not known before the subject is
        sd = {k: v.detach() for k, v in self.named_parameters()}
        folded = W.fold(sd, self.config["material_feature"])
        mcfg = self.config.get("material") or {}
        if any(k in mcfg for k in _MATERIAL_AFFINE):     # configs/material/shallow_mlp.yaml:4-9
            g = lambda k, d: float(mcfg.get(k, d))
            folded["mat_scale"] = [g("albedo_scale", 0.77)] * 3 + [g("roughness_scale", 0.9), g("metallic_scale", 1.0)]
            folded["mat_bias"] = [g("albedo_bias", 0.03)] * 3 + [g("roughness_bias", 0.09), g("metallic_bias", 0.0)]
        self._beta = folded["beta"]
        self.engine.set_fields(folded, self.layout, self.setup_snarf.bbox)

    def update_step(self, epoch, global_step):
        # eval-time state after update_module_step(model, ...) at test: all hash levels / SH bands on
        self.enable_phys = global_step >= self.config["phys_kick_in_step"]
        self.importance_sample = global_step > self.config["importance_sample_kick_in_step"]
        if not (self.enable_phys and self.importance_sample):
            raise NotImplementedError("the accelerated path implements the fully warmed-up eval state only")

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("forward() implements the eval render path; the training-mode forward / backward is "
                                      "intrinsicavatar_b200.train (fused_query, shade_fields, volrend, pbr_light; SURVEY.md 8f.4)")
        return super().train(False)

    # -------------------------------------------------------------------- prepare ----
    def _apply_config(self):
        ratio = getattr(self, "albedo_align_ratio", None)
        cfg = self.config
        self.engine.set_render_config(
            cfg["scene_aabb"], cfg["num_samples_per_ray"], cfg["num_samples_per_secondary_ray"],
            cfg["secondary_near_plane"], cfg["secondary_far_plane"], cfg["grid_prune_occ_thre"],
            np.asarray(torch.as_tensor(self.background_color).cpu(), np.float32),
            None if ratio is None else np.asarray(torch.as_tensor(ratio).cpu(), np.float32))
        self.engine.set_secondary_sampling(cfg["secondary_importance_sample"], cfg["zero_crossing_search"])

    def prepare(self, batch: dict, jitter=None, light_uniforms=None):
        """models/intrinsic_avatar.py:281-305.  ``batch`` holds betas[1,10] (first call: the subject is built from them,
        snarf_deformer.py:89-91), body_pose[1,69], global_orient[1,3], transl[1,3] and optionally hdri[H,W,3].
        Randomness of the reference (occupancy jitter, light sample uniforms) is drawn from torch's generator unless
        given explicitly."""
        def _np(x):
            return np.asarray(torch.as_tensor(x).detach().cpu(), np.float32).reshape(-1)
        betas = _np(batch["betas"])[:10] if "betas" in batch else None
        if self.setup_snarf is None:
            self._betas = betas
            self._init_subject(self._resolve_body(betas))
        elif betas is not None and self._betas is not None and not np.allclose(betas, self._betas, atol=1e-6):
            # optimize_betas is false: the shape is fixed at initialisation (snarf_deformer.py:83-91)
            raise ValueError("prepare: batch['betas'] differ from the betas the subject was initialised with")
        fr = self.setup_snarf.frame(_np(batch["body_pose"]), _np(batch["global_orient"]), _np(batch["transl"]))
        self._frame = fr
        self._apply_config()
        self.engine.set_pose(fr["tfs"], fr["w2s"])
        res = self.config["occ_resolution"]
        if jitter is None:
            jitter = torch.rand(res ** 3, 3, 3, device=self.engine.dev)
        self.engine.build_occupancy(fr["deformed_bbox"], jitter, res)
        if self.enable_phys and "hdri" in batch:
            spp = self.config["samples_per_pixel"]
            resample = self.config["resample_light"] or self._light_key is None
            if resample:
                # emitter.base = hdri; update_pdf(); emitter.sample(spp) (models/intrinsic_avatar.py:291-301).  With
                # resample_light = false all of that happens once: later frames keep the first envmap, pdf and directions.
                if light_uniforms is None:
                    light_uniforms = (torch.rand(spp, device=self.engine.dev), torch.rand(spp, device=self.engine.dev))
                self._light_key = light_uniforms
                self._light_hdri = batch["hdri"]
            # the directions are used in the per-frame SMPL-root frame: the tables are rebuilt from the kept envmap / uniforms
            if self.config["render_mode"] == "uniform_light":
                self.engine.set_light_uniform(self._light_hdri, 16, 32)
            else:
                self.engine.set_light(self._light_hdri, self._light_key[0], self._light_key[1])

    # -------------------------------------------------------------------- forward ----
    def forward(self, rays: torch.Tensor, move_to_cpu: bool = True) -> dict:
        """models/intrinsic_avatar.py:1653-1666 + forward_ :950-1651 (eval).  rays [N,8] (any device)."""
        if self._frame is None:
            raise RuntimeError("call prepare(batch) before forward(rays)")
        self._apply_config()
        dev = self.engine.dev
        r = rays.to(dev, torch.float32, non_blocking=True)
        primary_only = self.albedo_only or not self.engine.spp
        o = self.engine.render(r, primary_only=primary_only, gi=bool(self.config["global_illumination"]), seed=self.seed,
                               render_mode=self.config["render_mode"], add_emitter=bool(self.config["add_emitter"]),
                               check_overflow=True)
        n = r.shape[0]
        if move_to_cpu:
            # eval outputs are CPU tensors (models/utils.py:48-55): one packed D2H copy into pinned memory
            o = self.engine.outputs_to_host(o)
            dev = torch.device("cpu")
        bg = torch.as_tensor(self.background_color, dtype=torch.float32).to(dev)
        valid = o["opacity"] > 0
        out = {
            "comp_rgb": o["comp_rgb"], "comp_normal": o["comp_normal"], "opacity": o["opacity"], "depth": o["depth"],
            "rays_valid": valid, "rays_valid_phys": valid,
            # one entry per ray_chunk rays, as chunk_batch concatenates forward_'s per-chunk count (models/utils.py:16-61;
            # SURVEY Appendix A.16): [ceil(N / ray_chunk)] int32
            "num_samples": self._chunk_sums(o["num_samples"].reshape(-1), int(self.config["ray_chunk"])),
            "comp_rgb_phys": o["comp_rgb_phys"], "comp_demod_phys": o["comp_demod_phys"],
            "comp_albedo": o["comp_albedo"], "comp_metallic": o["comp_metallic"], "comp_roughness": o["comp_roughness"],
        }
        if self.config["render_mode"] == "uniform_light":
            out["visibility"] = o["visibility"]      # models/intrinsic_avatar.py:1516-1517
        zeros_b = torch.zeros_like(valid)
        out_bg = {
            "comp_rgb": bg[None].expand(n, 3), "num_samples": torch.zeros_like(out["num_samples"]),
            "rays_valid": zeros_b, "rays_valid_phys": zeros_b,
            "comp_albedo": torch.zeros(1, 3, device=dev).expand(n, 3),
            "comp_metallic": bg.mean().reshape(1, 1).expand(n, 1), "comp_roughness": bg.mean().reshape(1, 1).expand(n, 1),
        }
        out_full = {
            "comp_rgb": o["comp_rgb_full"], "num_samples": out["num_samples"], "rays_valid": valid,
            "rays_valid_phys": valid, "comp_rgb_phys": o["comp_rgb_phys_full"],
            "comp_demod_phys": o["comp_demod_phys_full"], "comp_albedo": o["comp_albedo_full"],
            "comp_metallic": o["comp_metallic_full"], "comp_roughness": o["comp_roughness_full"],
        }
        res = {**out, **{k + "_bg": v for k, v in out_bg.items()}, **{k + "_full": v for k, v in out_full.items()}}
        res["beta"] = torch.tensor(self._beta)
        return res

    @staticmethod
    def _chunk_sums(ns: torch.Tensor, chunk: int) -> torch.Tensor:
        n = ns.shape[0]
        n_chunks = max(1, (n + chunk - 1) // chunk)
        pad = n_chunks * chunk - n
        if pad:
            ns = torch.cat([ns, ns.new_zeros(pad)])
        return ns.reshape(n_chunks, chunk).sum(1).to(torch.int32)

    # ---------------------------------------------------------------- conveniences ----
    def render_image(self, batch: dict, rays: torch.Tensor, H: int, W_: int, **kw) -> dict:
        """prepare + primary volume render (no secondary rays): images [H,W,C]."""
        old = self.albedo_only
        self.albedo_only = True
        try:
            self.prepare({k: v for k, v in batch.items() if k != "hdri"}, **kw)
            out = self.forward(rays)
        finally:
            self.albedo_only = old
        return {k: v.reshape(H, W_, -1) for k, v in out.items() if torch.is_tensor(v) and v.shape[:1] == (H * W_,)}

    def render_image_relight(self, batch: dict, rays: torch.Tensor, H: int, W_: int, **kw) -> dict:
        """prepare (with batch['hdri']) + full relighting forward: images [H,W,C]."""
        self.prepare(batch, **kw)
        out = self.forward(rays)
        return {k: v.reshape(H, W_, -1) for k, v in out.items() if torch.is_tensor(v) and v.shape[:1] == (H * W_,)}
