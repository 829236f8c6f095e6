"""Drop-in model for the reference's ``IntrinsicAvatarModel`` on the render path.

Keeps the reference surface (models/intrinsic_avatar.py:166-305, 1653-1674):
  ``prepare(batch)``, ``forward(rays[N,8]) -> dict`` (CPU tensors in eval, same keys incl. ``*_bg`` /
  ``*_full`` and ``beta``), ``update_step(epoch, global_step)``, ``train()/eval()``, externally set
  attributes ``background_color``, ``albedo_only``, ``albedo_align_ratio``, ``t_idx``; ``state_dict`` /
  ``load_state_dict`` with the reference's parameter keys (weights.py).
It adds ``render_image`` / ``render_image_relight`` conveniences named by BASELINE.json.

Only the eval render path is implemented (render_mode = light | uniform_light | mats | mis, with or without
global_illumination / add_emitter); training-mode calls raise.
Every numeric step runs in libia_b200.so; this file is glue (pose -> 24 matrices, pointer passing).
"""
from __future__ import annotations

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


class IntrinsicAvatarModel(torch.nn.Module):
    def __init__(self, config: dict | None = None, body=None, device: int | None = None, seed: int = 0):
        super().__init__()
        self.config = dict(DEFAULT_CONFIG)
        if config:
            self.config.update(config)
        cfg = self.config
        if cfg["render_mode"] not in ("light", "uniform_light", "mats", "mis"):
            # same failure as the reference's dispatch (models/intrinsic_avatar.py:1435-1438)
            raise NotImplementedError(f"Render mode {cfg['render_mode']} not supported.")
        if cfg["render_mode"] == "uniform_light":
            assert cfg["samples_per_pixel"] == 512  # models/intrinsic_avatar.py:1391 (16 x 32 stratified sphere)
        if not (cfg["secondary_importance_sample"] and cfg["zero_crossing_search"]) or cfg["material_feature"] != "hybrid":
            raise NotImplementedError("non-default secondary sampling / material_feature not supported")
        self.engine = RenderEngine(device)
        self.setup_snarf = SnarfSetup(body)
        self.layout = W.hashgrid_layout()
        self._params = torch.nn.ParameterDict()  # reference-keyed parameters ('.' -> '/')
        self.background_color = torch.ones(3)
        self.albedo_only = False
        self.t_idx = 0.0
        self.enable_phys = True
        self.importance_sample = True
        self.seed = seed
        self._frame = None
        self._light_key = None
        self.engine.set_lbs_voxels(self.setup_snarf.lbs_voxel, self.setup_snarf.offset_kernel,
                                   self.setup_snarf.scale_kernel)
        self.load_state_dict(W.random_state_dict(seed))

    # ---------------------------------------------------------------- parameters ----
    def state_dict(self, *a, **k):
        return {key.replace("/", "."): v.detach() for key, v in self._params.items()}

    def load_state_dict(self, sd, strict=False):
        for k, v in sd.items():
            if k.startswith("model."):
                k = k[len("model."):]
            if k.split(".")[0] in ("geometry", "radiance", "material", "density"):
                self._params[k.replace(".", "/")] = torch.nn.Parameter(torch.as_tensor(v).clone().float(),
                                                                      requires_grad=False)
        self._upload_fields()

    def load_checkpoint(self, path: str):
        """Ingest a Lightning checkpoint of the reference (launch.py:110-124): the render-path parameters under
        ``model.`` are taken, everything else is ignored (strict=False semantics)."""
        self.load_state_dict(W.load_lightning_checkpoint(path))

    def _upload_fields(self):
        folded = W.fold(self.state_dict())
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
            raise NotImplementedError("training is out of scope of the render-path drop-in (SURVEY.md 8f.4)")
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

    def prepare(self, batch: dict, jitter=None, light_uniforms=None):
        """models/intrinsic_avatar.py:281-305.  ``batch`` holds body_pose[1,69], global_orient[1,3],
        transl[1,3] and optionally hdri[H,W,3].  Randomness of the reference (occupancy jitter, light
        sample uniforms) is drawn from torch's generator unless given explicitly."""
        def _np(x):
            return np.asarray(torch.as_tensor(x).detach().cpu(), np.float32).reshape(-1)
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
                if light_uniforms is None:
                    light_uniforms = (torch.rand(spp, device=self.engine.dev), torch.rand(spp, device=self.engine.dev))
                self._light_key = light_uniforms
            # directions live in the per-frame SMPL-root frame: refresh the tables every frame
            if self.config["render_mode"] == "uniform_light":
                self.engine.set_light_uniform(batch["hdri"], 16, 32)
            else:
                self.engine.set_light(batch["hdri"], self._light_key[0], self._light_key[1])

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
