"""Thin host wrapper over the C ABI: owns a context, keeps referenced device tensors alive, turns
torch tensors into pointers.  All compute happens in libia_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import IaOutputs, check, fptr, ptr

_vp, _i32, _i64, _cf32, _u32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32

_ARGTYPES = {
    "ia_create": [C.POINTER(_vp), _i32],
    "ia_destroy": [_vp],
    "ia_set_fields": [_vp, _vp, _vp, _i64] + [_vp] * 4 + [_vp] * 18 + [_vp, _cf32, _vp],
    "ia_set_lbs_voxels": [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp],
    "ia_smpl_lbs": [_vp] * 7 + [_i32, _i32] + [_vp] * 7,
    "ia_voxelize_lbs": [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp],
    "ia_set_pose": [_vp, _vp, _vp, _vp],
    "ia_set_render_config": [_vp, _vp, _i32, _i32, _cf32, _cf32, _cf32, _vp, _vp],
    "ia_reserve_samples": [_vp, _i64],
    "ia_set_secondary_sampling": [_vp, _i32, _i32],
    "ia_build_occupancy": [_vp, _vp, _i32, _vp, _vp, _vp],
    "ia_set_occupancy": [_vp, _vp, _i32, _vp, _vp],
    "ia_set_light": [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp],
    "ia_set_light_uniform": [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp],
    "ia_render": [_vp, _vp, _i64, _i64, _i32, _u32, C.POINTER(IaOutputs), _vp],
    "ia_get_counters": [_vp, _vp, _vp],
    "ia_set_timing": [_vp, _i32],
    "ia_get_timings": [_vp, _vp, _vp, _vp],
    "ia_op_precompute": [_vp, _vp, _vp],
    "ia_op_broyden": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "ia_op_query": [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ia_op_shade_fields": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "ia_op_geometry": [_vp, _vp, _i64, _vp, _vp],
    "ia_op_geometry_backward": [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "ia_op_deform_backward": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "ia_op_query_train": [_vp, _vp, _i64] + [_vp] * 9,
    "ia_op_pbr_shade": [_vp] * 9 + [_i64] + [_vp] * 4,
    "ia_op_pbr_shade_backward": [_vp] * 12 + [_i64] + [_vp] * 6,
    "ia_op_env_backward": [_vp, _vp, _vp, _i64, _vp, _vp],
    "ia_op_volrend": [_vp, _vp, _vp, _vp, _vp, _i32, _cf32, _i64, _vp, _vp, _vp, _vp],
    "ia_op_volrend_backward": [_vp, _vp, _vp, _vp, _vp, _i32, _cf32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "ia_op_shade_fields_backward": [_vp] * 7 + [_i64] + [_vp] * 6,
    "ia_op_query_backward": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "ia_op_traverse": [_vp, _vp, _vp, _i64, _cf32, _cf32, _cf32] + [_vp] * 9 + [_vp],
    "ia_op_ray_resampling": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp] + [_vp] * 6 + [_vp],
    "ia_op_ray_resampling_merge": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp] + [_vp] * 6 + [_vp],
    "ia_op_ray_resampling_sdf_fine": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "ia_op_ray_resampling_fine": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "ia_op_unpack_info": [_vp, _vp, _i64, _vp, _vp],
    "ia_op_secondary": [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp],
    "ia_op_brdf": [_vp] * 7 + [_i64, _vp, _vp, _vp],
    "ia_op_bsdf_sample_pdf": [_vp] * 8 + [_i64, _vp, _vp, _vp],
    "ia_op_env": [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "ia_make_rays": [_vp, _vp, _vp, _vp, _i32, _i32, _cf32, _cf32, _vp, _vp],
    "ia_pack_rgb8": [_vp, _vp, _i64, _i32, _cf32, _cf32, _i32, _vp, _vp],
    "ia_update_occupancy_ema": [_vp, _vp, _i32, _vp, _vp, _cf32, _cf32, _vp, _vp],
    "ia_pack_grid8": [_vp, _vp, _i32, _i32, _i32, _i32, _cf32, _cf32, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
}

OUTPUT_SPECS = [  # name, channels, dtype
    ("comp_rgb", 3, torch.float32), ("comp_normal", 3, torch.float32), ("opacity", 1, torch.float32),
    ("depth", 1, torch.float32), ("comp_albedo", 3, torch.float32), ("comp_roughness", 1, torch.float32),
    ("comp_metallic", 1, torch.float32), ("comp_rgb_phys", 3, torch.float32),
    ("comp_demod_phys", 3, torch.float32), ("num_samples", 1, torch.int32),
    ("comp_rgb_full", 3, torch.float32), ("comp_rgb_phys_full", 3, torch.float32),
    ("comp_demod_phys_full", 3, torch.float32), ("comp_albedo_full", 3, torch.float32),
    ("comp_roughness_full", 1, torch.float32), ("comp_metallic_full", 1, torch.float32),
    ("visibility", 1, torch.float32),
]


def _lib():
    lib = capi.load()
    if not getattr(lib, "_ia_typed", False):
        for name, at in _ARGTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes = at
            fn.restype = C.c_int
        lib._ia_typed = True
    return lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class RenderEngine:
    """One libia_b200 context on one CUDA device."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("RenderEngine needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        h = C.c_void_p()
        check(self.lib.ia_create(C.byref(h), self.device), "ia_create")
        self.h = h
        self._keep = {}
        self.spp = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.ia_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------- state ----
    def set_fields(self, folded: dict, layout: dict, bbox):
        geo = folded["geo_hash"].to(self.dev, torch.float32).contiguous()
        rad = folded["rad_hash"].to(self.dev, torch.float32).contiguous()
        self._keep["geo"], self._keep["rad"] = geo, rad
        names = ["geo_w1", "geo_b1", "geo_w2", "geo_b2", "rad_w1", "rad_b1", "rad_w2", "rad_b2", "rad_w3", "rad_b3",
                 "mat_w1", "mat_b1", "mat_w2", "mat_b2", "mat_w3", "mat_b3"]
        host = [_f32(folded[n].cpu().numpy()) for n in names]
        shapes = {"geo_w1": (64, 35), "geo_b1": (64,), "geo_w2": (13, 64), "geo_b2": (13,), "rad_w1": (64, 67), "rad_b1": (64,),
                  "rad_w2": (64, 64), "rad_b2": (64,), "rad_w3": (3, 64), "rad_b3": (3,), "mat_w1": (64, 48), "mat_b1": (64,),
                  "mat_w2": (64, 64), "mat_b2": (64,), "mat_w3": (5, 64), "mat_b3": (5,)}
        for n, a in zip(names, host):            # ia_set_fields copies fixed extents out of these host arrays
            if tuple(a.shape) != shapes[n]:
                raise ValueError(f"set_fields: {n} has shape {tuple(a.shape)}, libia_b200 is built for {shapes[n]}")
        lv = [_f32(layout["scale"]), np.ascontiguousarray(layout["res"], np.int32),
              np.ascontiguousarray(layout["size"], np.int32), np.ascontiguousarray(layout["offset"], np.int32)]
        mat_scale = _f32(folded.get("mat_scale", [0.77, 0.77, 0.77, 0.9, 1.0]))
        mat_bias = _f32(folded.get("mat_bias", [0.03, 0.03, 0.03, 0.09, 0.0]))
        bb = _f32(bbox).reshape(6)
        # both hash tables are gathered with the same level layout: the smaller one bounds what may be indexed
        n_entries = min(geo.numel(), rad.numel()) // 2
        check(self.lib.ia_set_fields(self.h, ptr(geo), ptr(rad), n_entries, *[fptr(a) for a in lv],
                                     *[fptr(a) for a in host], fptr(mat_scale), fptr(mat_bias), fptr(bb),
                                     float(folded["beta"]), _stream()), "ia_set_fields")

    def set_lbs_voxels(self, lbs_voxel, offset_kernel, scale_kernel):
        v = torch.as_tensor(lbs_voxel, dtype=torch.float32).to(self.dev).contiguous()
        assert v.dim() == 4 and v.shape[0] == 24
        check(self.lib.ia_set_lbs_voxels(self.h, ptr(v), v.shape[1], v.shape[2], v.shape[3], fptr(_f32(offset_kernel)),
                                         fptr(_f32(scale_kernel)), _stream()), "ia_set_lbs_voxels")
        torch.cuda.current_stream().synchronize()  # v is repacked into the context; safe to drop

    # ------------------------------------------------------------- subject set-up ----
    def smpl_arrays(self, body):
        """Upload the arrays of a body.SMPLBody once (device fp32, the layouts ia_smpl_lbs takes)."""
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).to(self.dev)
        return {"v_template": t(body.v_template), "shapedirs": t(body.shapedirs), "posedirs": t(body.posedirs),
                "J_regressor": t(body.J_regressor), "lbs_weights": t(body.lbs_weights),
                "parents": np.ascontiguousarray(body.parents, np.int32), "V": int(body.v_template.shape[0]),
                "NB": int(body.shapedirs.shape[-1])}

    def smpl_lbs(self, arrays, betas, body_pose, global_orient, transl):
        """SMPL forward on the device: (vertices [V,3], joints [24,3], A [24,4,4]) CUDA tensors."""
        V, NB = arrays["V"], arrays["NB"]
        b = np.zeros(NB, np.float32)
        bb = _f32(betas).reshape(-1)[:NB]
        b[:len(bb)] = bb
        pose = _f32(np.concatenate([_f32(global_orient).reshape(3), _f32(body_pose).reshape(69)]))
        tr = _f32(transl).reshape(3)
        verts = torch.empty(V, 3, device=self.dev)
        joints = torch.empty(24, 3, device=self.dev)
        A = torch.empty(24, 4, 4, device=self.dev)
        check(self.lib.ia_smpl_lbs(self.h, ptr(arrays["v_template"]), ptr(arrays["shapedirs"]), ptr(arrays["posedirs"]),
                                   ptr(arrays["J_regressor"]), ptr(arrays["lbs_weights"]), fptr(arrays["parents"]), V, NB,
                                   fptr(b), fptr(pose), fptr(tr), ptr(verts), ptr(joints), ptr(A), _stream()), "ia_smpl_lbs")
        return verts, joints, A

    def voxelize_lbs(self, verts, weights, resolution=128):
        """Skinning-weight voxel grid on the device: (lbs_voxel [24, res/4, res, res] CUDA, offset_kernel [3], scale_kernel [3])."""
        v = torch.as_tensor(verts, dtype=torch.float32).to(self.dev).reshape(-1, 3).contiguous()
        w = torch.as_tensor(weights, dtype=torch.float32).to(self.dev).reshape(-1, 24).contiguous()
        assert v.shape[0] == w.shape[0]
        vox = torch.empty(24, resolution // 4, resolution, resolution, device=self.dev)
        off, scl = np.zeros(3, np.float32), np.zeros(3, np.float32)
        check(self.lib.ia_voxelize_lbs(self.h, ptr(v), ptr(w), v.shape[0], int(resolution), ptr(vox), fptr(off), fptr(scl),
                                       _stream()), "ia_voxelize_lbs")
        return vox, off, scl

    def set_pose(self, tfs, w2s):
        check(self.lib.ia_set_pose(self.h, fptr(_f32(tfs).reshape(24 * 16)), fptr(_f32(w2s).reshape(16)), _stream()),
              "ia_set_pose")

    def set_render_config(self, scene_aabb, num_samples_per_ray=128, num_samples_per_secondary_ray=64,
                          secondary_near=0.0, secondary_far=1.5, occ_thre=0.001, background=(1, 1, 1),
                          albedo_align_ratio=None):
        ratio = fptr(_f32(albedo_align_ratio)) if albedo_align_ratio is not None else C.c_void_p(0)
        check(self.lib.ia_set_render_config(self.h, fptr(_f32(scene_aabb)), num_samples_per_ray,
                                            num_samples_per_secondary_ray, secondary_near, secondary_far, occ_thre,
                                            fptr(_f32(background)), ratio), "ia_set_render_config")

    def set_secondary_sampling(self, importance_sample=True, zero_crossing_search=True):
        check(self.lib.ia_set_secondary_sampling(self.h, int(bool(importance_sample)), int(bool(zero_crossing_search))),
              "ia_set_secondary_sampling")

    def build_occupancy(self, aabb, jitter, res=64, return_grid=False):
        j = torch.as_tensor(jitter, dtype=torch.float32).to(self.dev).contiguous()
        assert j.numel() == res ** 3 * 9
        out = torch.empty(res ** 3, dtype=torch.uint8, device=self.dev) if return_grid else None
        check(self.lib.ia_build_occupancy(self.h, fptr(_f32(aabb).reshape(6)), res, ptr(j), ptr(out), _stream()),
              "ia_build_occupancy")
        self._keep["jitter"] = j
        return out.bool().reshape(res, res, res) if return_grid else None

    def update_occupancy_ema(self, aabb, jitter, occs, res=64, ema_decay=0.8, occ_thre=0.001, return_grid=False):
        """Training-time grid update (OccGridEstimator._update): ``occs`` [res^3] float CUDA tensor, the EMA state of the
        frame's level, updated IN PLACE; jitter [res^3, 3]."""
        j = torch.as_tensor(jitter, dtype=torch.float32).to(self.dev).contiguous()
        if j.numel() != res ** 3 * 3 or not (occs.is_cuda and occs.is_contiguous() and occs.dtype == torch.float32
                                            and occs.numel() == res ** 3):
            raise ValueError("update_occupancy_ema: jitter [res^3, 3] and a contiguous float32 CUDA occs [res^3] are required")
        out = torch.empty(res ** 3, dtype=torch.uint8, device=self.dev) if return_grid else None
        check(self.lib.ia_update_occupancy_ema(self.h, fptr(_f32(aabb).reshape(6)), res, ptr(j), ptr(occs), float(ema_decay),
                                               float(occ_thre), ptr(out), _stream()), "ia_update_occupancy_ema")
        self._keep["jitter"] = j
        return out.bool().reshape(res, res, res) if return_grid else None

    def set_occupancy(self, aabb, binaries):
        b = torch.as_tensor(binaries).to(self.dev).to(torch.uint8).contiguous()
        res = round(b.numel() ** (1 / 3))
        check(self.lib.ia_set_occupancy(self.h, fptr(_f32(aabb).reshape(6)), res, ptr(b), _stream()), "ia_set_occupancy")
        torch.cuda.current_stream().synchronize()

    def set_light(self, envmap, u1, u2, return_tables=False):
        env = torch.as_tensor(envmap, dtype=torch.float32).to(self.dev).contiguous()
        u1 = torch.as_tensor(u1, dtype=torch.float32).to(self.dev).contiguous()
        u2 = torch.as_tensor(u2, dtype=torch.float32).to(self.dev).contiguous()
        self._keep["env"], self._keep["u1"], self._keep["u2"] = env, u1, u2
        spp = u1.numel()
        H, W = env.shape[:2]
        d = e = p = None
        if return_tables:
            d = torch.empty(spp, 3, device=self.dev)
            e = torch.empty(spp, 3, device=self.dev)
            p = torch.empty(spp, device=self.dev)
        check(self.lib.ia_set_light(self.h, ptr(env), H, W, ptr(u1), ptr(u2), spp, ptr(d), ptr(e), ptr(p), _stream()),
              "ia_set_light")
        self.spp = spp
        return (d, e, p) if return_tables else None

    def set_light_uniform(self, envmap, n_rows=16, n_cols=32, return_tables=False):
        """render_mode = uniform_light: stratified-sphere light table (samples_per_pixel = n_rows * n_cols)."""
        env = torch.as_tensor(envmap, dtype=torch.float32).to(self.dev).contiguous()
        self._keep["env"] = env
        spp = n_rows * n_cols
        H, W = env.shape[:2]
        d = e = None
        if return_tables:
            d = torch.empty(spp, 3, device=self.dev)
            e = torch.empty(spp, 3, device=self.dev)
        check(self.lib.ia_set_light_uniform(self.h, ptr(env), H, W, n_rows, n_cols, ptr(d), ptr(e), _stream()),
              "ia_set_light_uniform")
        self.spp = spp
        return (d, e) if return_tables else None

    # ------------------------------------------------------- frame producer / consumer ----
    def make_rays(self, K, H, W, near, far, c2w=None, w2c=None, out=None):
        """AnimationDataset rays [H*W,8] generated on the device (datasets/animation.py:13-34, 163-189)."""
        def m34(a):
            return None if a is None else np.ascontiguousarray(np.asarray(a, np.float64)[:3, :4])
        Kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(K, np.float64)))
        c = m34(c2w)
        e = m34(np.linalg.inv(np.asarray(w2c, np.float32))) if w2c is not None else None   # c2w = inv(w2c), float32
        rays = out if out is not None else torch.empty(H * W, 8, device=self.dev)
        vp = lambda a: C.c_void_p(0) if a is None else a.ctypes.data_as(C.c_void_p)
        check(self.lib.ia_make_rays(self.h, vp(Kinv), vp(c), vp(e), H, W, float(near), float(far), ptr(rays), _stream()),
              "ia_make_rays")
        return rays

    def pack_rgb8(self, img, data_range=(0.0, 1.0), bgr=False):
        """SaverMixin.get_rgb_image_ on the device: float [n, C] (CUDA) -> uint8 [n, C]."""
        img = img.to(self.dev, torch.float32).contiguous()
        n, ch = img.reshape(-1, img.shape[-1]).shape
        out = torch.empty(n, ch, dtype=torch.uint8, device=self.dev)
        check(self.lib.ia_pack_rgb8(self.h, ptr(img), n, ch, float(data_range[0]), float(data_range[1]), int(bgr), ptr(out),
                                    _stream()), "ia_pack_rgb8")
        return out.reshape(img.shape)

    def pack_grid8(self, grid, x0, img, kind="rgb", data_range=(0.0, 1.0), lut=None, bgr=False):
        """One column of SaverMixin.get_image_grid_ written into the uint8 grid [H, grid_w, 3] (CUDA) at pixel column
        x0.  img [H, W, C] float (C <= 3; grayscale: [H, W]); data_range None = the image's own min / max after
        nan_to_num (utils/mixins.py:89-91), computed on the device; lut uint8 [256, 3] (CUDA) or None."""
        H, gw = int(grid.shape[0]), int(grid.shape[1])
        img = img.to(self.dev, torch.float32)
        if img.dim() == 2:
            img = img[..., None]
        img = img.contiguous()
        if tuple(img.shape[:1]) != (H,) or grid.dtype != torch.uint8 or not grid.is_contiguous():
            raise ValueError("pack_grid8: the image must have the grid's height and the grid must be contiguous uint8")
        W, ch = int(img.shape[1]), int(img.shape[2])
        rng, lo, hi = None, 0.0, 1.0
        if data_range is None:
            mn, mx = torch.aminmax(torch.nan_to_num(img))
            rng = torch.stack([mn, mx]).contiguous()
        else:
            lo, hi = float(data_range[0]), float(data_range[1])
        check(self.lib.ia_pack_grid8(self.h, ptr(img), H, W, ch, 0 if kind == "rgb" else 1, lo, hi,
                                     ptr(rng) if rng is not None else None, ptr(lut) if lut is not None else None,
                                     ptr(grid), gw, int(x0), int(bgr), _stream()), "ia_pack_grid8")
        return grid

    # ------------------------------------------------------------------------ render ----
    def alloc_outputs(self, n, device=None, pin_memory=False):
        """All output buffers of one forward pass as views of ONE block (key "_block", 4-byte words), so that a
        frame leaves the device with a single copy.  Each view is a contiguous [n, C] tensor as the C ABI wants."""
        total = sum(ch for _, ch, _ in OUTPUT_SPECS) * n
        block = torch.empty(total, dtype=torch.float32, device=self.dev if device is None else device,
                            pin_memory=pin_memory)
        out, off = {"_block": block}, 0
        for name, ch, dt in OUTPUT_SPECS:
            seg = block[off:off + n * ch]
            out[name] = (seg if dt == torch.float32 else seg.view(dt)).view(n, ch)
            off += n * ch
        return out

    def outputs_to_host(self, out):
        """One D2H copy of a packed output set into pinned host memory; returns the same views on the host."""
        n = out["opacity"].shape[0]
        host = self.alloc_outputs(n, device="cpu", pin_memory=True)
        host["_block"].copy_(out["_block"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host

    def render(self, rays: torch.Tensor, *, primary_only=False, gi=False, seed=0, ray_index_base=0, outputs=None,
               render_mode="light", add_emitter=False, check_overflow=False):
        """rays: CUDA float32 [n,8].  Returns dict of CUDA tensors (no sync unless ``check_overflow``).

        ``check_overflow``: read the overflow counter after the launch (syncs the stream); a frame whose primary samples
        did not fit the pool is rendered again with a pool grown to the per-ray maximum, and a ray with more edges than
        the per-ray capacity raises -- such a frame is never returned silently."""
        assert rays.is_cuda and rays.dtype == torch.float32 and rays.shape[-1] == 8
        rays = rays.contiguous()
        n = rays.shape[0]
        out = outputs if outputs is not None else self.alloc_outputs(n)
        st = IaOutputs(**{name: out[name].data_ptr() for name, _, _ in OUTPUT_SPECS})
        flags = (capi.RENDER_PRIMARY_ONLY if primary_only else 0) | (capi.RENDER_GI if gi else 0) | \
            capi.RENDER_MODES[render_mode] | (capi.RENDER_ADD_EMITTER if add_emitter else 0)
        check(self.lib.ia_render(self.h, ptr(rays), n, ray_index_base, flags, seed, C.byref(st), _stream()), "ia_render")
        if check_overflow and self.counters()["overflow"]:
            self.reserve_samples(n * capi.MAX_SAMPLES_PER_RAY)
            check(self.lib.ia_render(self.h, ptr(rays), n, ray_index_base, flags, seed, C.byref(st), _stream()), "ia_render")
            over = self.counters()["overflow"]
            if over:
                raise RuntimeError(f"ia_render: {over} rays exceed the per-ray edge capacity ({capi.MAX_SAMPLES_PER_RAY}); "
                                   "lower num_samples_per_ray")
        return out

    def reserve_samples(self, n_samples: int):
        check(self.lib.ia_reserve_samples(self.h, int(n_samples)), "ia_reserve_samples")

    def counters(self) -> dict:
        a = np.zeros(capi.N_COUNTERS, np.uint64)
        check(self.lib.ia_get_counters(self.h, fptr(a), _stream()), "ia_get_counters")
        out = {k: int(a[i]) for i, k in enumerate(capi.COUNTER_NAMES)}
        # counters as they stood after the primary stage (hit_rays is a launch-wide figure, not a stage counter)
        out["primary"] = {k: int(a[capi.CNT_PRIMARY_BASE + i]) for i, k in enumerate(capi.COUNTER_NAMES) if i > 0}
        return out

    def set_timing(self, enable=True):
        check(self.lib.ia_set_timing(self.h, int(enable)), "ia_set_timing")

    def timings(self):
        """(dict stage -> ms of its last execution, total kernel launches); syncs the stream."""
        ms = np.zeros(len(capi.STAGE_NAMES), np.float32)
        n = C.c_uint64(0)
        check(self.lib.ia_get_timings(self.h, fptr(ms), C.byref(n), _stream()), "ia_get_timings")
        return {k: float(ms[i]) for i, k in enumerate(capi.STAGE_NAMES)}, int(n.value)

    # ---------------------------------------------------------------------- op level ----
    def op_precompute(self, D=32, H=128, W=128):
        out = torch.empty(12, D, H, W, device=self.dev)
        check(self.lib.ia_op_precompute(self.h, ptr(out), _stream()), "ia_op_precompute")
        return out

    def op_broyden(self, xd, with_jinv=True):
        xd = xd.to(self.dev, torch.float32).contiguous()
        n = xd.shape[0]
        x = torch.zeros(n, 13, 3, device=self.dev)
        J = torch.zeros(n, 13, 3, 3, device=self.dev) if with_jinv else None
        vr = torch.zeros(n, 13, dtype=torch.uint8, device=self.dev)
        v = torch.zeros(n, 13, dtype=torch.uint8, device=self.dev)
        check(self.lib.ia_op_broyden(self.h, ptr(xd), n, ptr(x), ptr(J), ptr(vr), ptr(v), _stream()), "ia_op_broyden")
        return x, J, vr.bool(), v.bool()

    def op_query(self, xd, with_grad=False):
        xd = xd.to(self.dev, torch.float32).contiguous()
        n = xd.shape[0]
        r = {"sdf": torch.empty(n, device=self.dev), "x_c": torch.empty(n, 3, device=self.dev),
             "valid": torch.empty(n, dtype=torch.uint8, device=self.dev)}
        g = gc = f = None
        if with_grad:
            g, gc, f = (torch.empty(n, 3, device=self.dev), torch.empty(n, 3, device=self.dev),
                        torch.empty(n, 13, device=self.dev))
            r.update(grad=g, grad_cano=gc, feature=f)
        check(self.lib.ia_op_query(self.h, ptr(xd), n, int(with_grad), ptr(r["sdf"]), ptr(r["x_c"]), ptr(r["valid"]),
                                   ptr(g), ptr(gc), ptr(f), _stream()), "ia_op_query")
        r["valid"] = r["valid"].bool()
        return r

    def op_shade_fields(self, xc, feature, view_world, normal_world):
        a = [t.to(self.dev, torch.float32).contiguous() for t in (xc, feature, view_world, normal_world)]
        n = a[0].shape[0]
        rgb, mat = torch.empty(n, 3, device=self.dev), torch.empty(n, 5, device=self.dev)
        check(self.lib.ia_op_shade_fields(self.h, *[ptr(t) for t in a], n, ptr(rgb), ptr(mat), _stream()),
              "ia_op_shade_fields")
        return rgb, mat

    def op_shade_fields_backward(self, xc, feature, view_world, normal_world, d_rgb, d_mat):
        """Gradients of  sum <d_rgb, rgb> + <d_mat, mat>  for the outputs of ``op_shade_fields``: dict with ``hash`` (radiance
        table), ``rad`` / ``mat`` = {w1, b1, w2, b2, w3, b3} in the shapes of the folded weights, and the input gradients
        ``x`` [n,3], ``feature`` [n,13], ``normal`` [n,3]."""
        a = [t.to(self.dev, torch.float32).contiguous() for t in (xc, feature, view_world, normal_world, d_rgb, d_mat)]
        n = a[0].shape[0]
        g_hash = torch.zeros_like(self._keep["rad"])
        g_mlp = torch.zeros(16332, device=self.dev)
        g_x, g_f, g_n = (torch.empty(n, 3, device=self.dev), torch.empty(n, 13, device=self.dev),
                         torch.empty(n, 3, device=self.dev))
        check(self.lib.ia_op_shade_fields_backward(self.h, *[ptr(t) for t in a], n, ptr(g_hash), ptr(g_mlp), ptr(g_x), ptr(g_f),
                                                   ptr(g_n), _stream()), "ia_op_shade_fields_backward")

        def net(o, n_in, n_out, pad):
            w1 = g_mlp[o:o + n_in * 64].reshape(n_in, 64).t().contiguous(); o += n_in * 64
            b1 = g_mlp[o:o + 64]; o += 64
            w2 = g_mlp[o:o + 4096].reshape(64, 64).t().contiguous(); o += 4096
            b2 = g_mlp[o:o + 64]; o += 64
            w3 = g_mlp[o:o + n_out * 64].reshape(n_out, 64); o += n_out * 64
            b3 = g_mlp[o:o + n_out]; o += pad
            return {"w1": w1, "b1": b1, "w2": w2, "b2": b2, "w3": w3, "b3": b3}, o
        rad, o = net(0, 67, 3, 4)
        mat, o = net(o, 48, 5, 8)
        return {"hash": g_hash, "rad": rad, "mat": mat, "x": g_x, "feature": g_f, "normal": g_n}

    def op_volrend(self, packed_info, sdf, dists, values, beta):
        """Laplace-density alpha, nerfacc weights and accumulation along packed rays: (weights [m], comp [n_rays,C], opacity)."""
        pi = packed_info.to(self.dev, torch.int32).contiguous()
        sdf, dists, values = [t.to(self.dev, torch.float32).contiguous() for t in (sdf, dists, values)]
        n_rays, C = pi.shape[0], values.shape[1]
        w = torch.empty_like(sdf)
        comp, op = torch.empty(n_rays, C, device=self.dev), torch.empty(n_rays, device=self.dev)
        check(self.lib.ia_op_volrend(self.h, ptr(pi), ptr(sdf), ptr(dists), ptr(values), C, float(beta), n_rays, ptr(w),
                                     ptr(comp), ptr(op), _stream()), "ia_op_volrend")
        return w, comp, op

    def op_volrend_backward(self, packed_info, sdf, dists, values, beta, d_comp, d_opacity=None, d_weights=None):
        """Backward of ``op_volrend``: (g_sdf [m], g_values [m,C], g_beta [1]) for upstream gradients on comp, opacity and
        (optionally) the weights themselves."""
        pi = packed_info.to(self.dev, torch.int32).contiguous()
        sdf, dists, values, d_comp = [t.to(self.dev, torch.float32).contiguous() for t in (sdf, dists, values, d_comp)]
        d_op = d_opacity.to(self.dev, torch.float32).contiguous() if d_opacity is not None else None
        d_w = d_weights.to(self.dev, torch.float32).reshape(sdf.shape).contiguous() if d_weights is not None else None
        n_rays, C = pi.shape[0], values.shape[1]
        g_sdf, g_val, g_beta = torch.zeros_like(sdf), torch.zeros_like(values), torch.zeros(1, device=self.dev)
        check(self.lib.ia_op_volrend_backward(self.h, ptr(pi), ptr(sdf), ptr(dists), ptr(values), C, float(beta), ptr(d_comp),
                                              ptr(d_op), ptr(d_w), n_rays, ptr(g_sdf), ptr(g_val), ptr(g_beta), _stream()),
              "ia_op_volrend_backward")
        return g_sdf, g_val, g_beta

    def _pbr_args(self, wi, n, wo, rough, albedo, metal, Li, inv_pdf):
        v3 = [t.to(self.dev, torch.float32).reshape(-1, 3).contiguous() for t in (wi, n, wo)]
        m = v3[0].shape[0]
        s1 = [t.to(self.dev, torch.float32).reshape(-1).contiguous() for t in (rough, metal, inv_pdf)]
        al, Li = [t.to(self.dev, torch.float32).reshape(-1, 3).contiguous() for t in (albedo, Li)]
        assert all(t.shape[0] == m for t in v3 + s1 + [al, Li])
        return m, [*v3, s1[0], al, s1[1], Li, s1[2]]

    def op_pbr_shade(self, wi, n, wo, rough, albedo, metal, Li, inv_pdf):
        """The integrators' differentiable combine (training path): (Lo, Lo_diff, Lo_spec) [n,3] of one light direction per
        shading sample -- MultiLobe.eval under the cosine mask, times Li and the inverse pdf."""
        m, a = self._pbr_args(wi, n, wo, rough, albedo, metal, Li, inv_pdf)
        Lo, Ld, Ls = (torch.empty(m, 3, device=self.dev) for _ in range(3))
        check(self.lib.ia_op_pbr_shade(self.h, *[ptr(t) for t in a], m, ptr(Lo), ptr(Ld), ptr(Ls), _stream()), "ia_op_pbr_shade")
        return Lo, Ld, Ls

    def op_pbr_shade_backward(self, wi, n, wo, rough, albedo, metal, Li, inv_pdf, d_Lo, d_Lo_diff=None, d_Lo_spec=None):
        """Backward of ``op_pbr_shade``: dict with ``normal`` [n,3], ``rough`` [n], ``albedo`` [n,3], ``metal`` [n], ``Li`` [n,3]."""
        m, a = self._pbr_args(wi, n, wo, rough, albedo, metal, Li, inv_pdf)
        up = [None if t is None else t.to(self.dev, torch.float32).reshape(m, 3).contiguous() for t in (d_Lo, d_Lo_diff, d_Lo_spec)]
        g_n, g_al, g_Li = (torch.empty(m, 3, device=self.dev) for _ in range(3))
        g_r, g_m = torch.empty(m, device=self.dev), torch.empty(m, device=self.dev)
        check(self.lib.ia_op_pbr_shade_backward(self.h, *[ptr(t) for t in a], *[ptr(t) for t in up], m, ptr(g_n), ptr(g_r),
                                                ptr(g_al), ptr(g_m), ptr(g_Li), _stream()), "ia_op_pbr_shade_backward")
        return {"normal": g_n, "rough": g_r, "albedo": g_al, "metal": g_m, "Li": g_Li}

    def op_env_backward(self, dirs_world, d_em, env_shape):
        """Gradient of  sum <d_em, emitter.eval(dirs_world)>  with respect to the texels of the map set by ``set_light*``
        (``env_shape`` = (H, W) of that map)."""
        d = dirs_world.to(self.dev, torch.float32).reshape(-1, 3).contiguous()
        g = d_em.to(self.dev, torch.float32).reshape(-1, 3).contiguous()
        assert d.shape == g.shape
        if "env" not in self._keep or tuple(self._keep["env"].shape[:2]) != (int(env_shape[0]), int(env_shape[1])):
            raise ValueError("op_env_backward: env_shape does not match the map of the last set_light* call "
                             "(the kernel scatters into that map's texel grid)")
        g_env = torch.zeros(int(env_shape[0]), int(env_shape[1]), 3, device=self.dev)
        check(self.lib.ia_op_env_backward(self.h, ptr(d), ptr(g), d.shape[0], ptr(g_env), _stream()), "ia_op_env_backward")
        return g_env

    def op_geometry(self, xc):
        """Canonical SDF of points [n,3] on the tensor-core path of the wavefront integrator's geometry phase."""
        xc = xc.to(self.dev, torch.float32).contiguous()
        sdf = torch.empty(xc.shape[0], device=self.dev)
        if xc.shape[0] == 0:
            return sdf
        check(self.lib.ia_op_geometry(self.h, ptr(xc), xc.shape[0], ptr(sdf), _stream()), "ia_op_geometry")
        return sdf

    def op_geometry_backward(self, xc, d_out):
        """Gradients of sum_i <d_out[i], net(xc[i])> for the geometry network: dict with ``hash`` (the table's layout),
        ``w1`` [64,35], ``b1`` [64], ``w2`` [13,64], ``b2`` [13] (effective weights) and ``x`` [n,3]."""
        xc = xc.to(self.dev, torch.float32).contiguous()
        d_out = d_out.to(self.dev, torch.float32).reshape(-1, 13).contiguous()
        n = xc.shape[0]
        g_hash = torch.zeros_like(self._keep["geo"])
        g_mlp = torch.zeros(3152, device=self.dev)
        g_x = torch.empty(n, 3, device=self.dev)
        check(self.lib.ia_op_geometry_backward(self.h, ptr(xc), ptr(d_out), n, ptr(g_hash), ptr(g_mlp), ptr(g_x), _stream()),
              "ia_op_geometry_backward")
        w1t, b1 = g_mlp[:35 * 64].reshape(35, 64), g_mlp[35 * 64:36 * 64]
        w2, b2 = g_mlp[36 * 64:49 * 64].reshape(13, 64), g_mlp[49 * 64:49 * 64 + 13]
        return {"hash": g_hash, "w1": w1t.t().contiguous(), "b1": b1, "w2": w2, "b2": b2, "x": g_x}

    def op_deform_backward(self, xc, valid, J_inv, g_xc):
        """Gradient with respect to the bone transforms (rows 0..2, [24,3,4]) of  sum <g_xc, x_c>  through the implicit-
        differentiation correction of the roots (ForwardDeformer.forward, training mode): inputs as ``op_broyden`` returns them."""
        xc = xc.to(self.dev, torch.float32).reshape(-1, 13, 3).contiguous()
        valid = valid.to(self.dev, torch.uint8).reshape(-1, 13).contiguous()
        J_inv = J_inv.to(self.dev, torch.float32).reshape(-1, 13, 3, 3).contiguous()
        g_xc = g_xc.to(self.dev, torch.float32).reshape(-1, 13, 3).contiguous()
        n = xc.shape[0]
        assert valid.shape[0] == n and J_inv.shape[0] == n and g_xc.shape[0] == n
        g_tfs = torch.zeros(24, 3, 4, device=self.dev)
        check(self.lib.ia_op_deform_backward(self.h, ptr(xc), ptr(valid), ptr(J_inv), ptr(g_xc), n, ptr(g_tfs), _stream()),
              "ia_op_deform_backward")
        return g_tfs

    def op_query_train(self, xd):
        """Training-mode forward of the fused query: ``op_query(with_grad=True)`` plus ``J_inv`` [n,3,3] and ``best`` [n] of the
        arg-min root (what ``op_query_backward`` reads)."""
        xd = xd.to(self.dev, torch.float32).contiguous()
        n = xd.shape[0]
        r = {"sdf": torch.empty(n, device=self.dev), "x_c": torch.empty(n, 3, device=self.dev),
             "valid": torch.empty(n, dtype=torch.uint8, device=self.dev), "grad": torch.empty(n, 3, device=self.dev),
             "grad_cano": torch.empty(n, 3, device=self.dev), "feature": torch.empty(n, 13, device=self.dev),
             "J_inv": torch.empty(n, 3, 3, device=self.dev), "best": torch.empty(n, dtype=torch.int32, device=self.dev)}
        check(self.lib.ia_op_query_train(self.h, ptr(xd), n, ptr(r["sdf"]), ptr(r["x_c"]), ptr(r["valid"]), ptr(r["grad"]),
                                         ptr(r["grad_cano"]), ptr(r["feature"]), ptr(r["J_inv"]), ptr(r["best"]), _stream()),
              "ia_op_query_train")
        r["valid"] = r["valid"].bool()
        return r

    def op_query_backward(self, fwd, d_out):
        """Backward of the fused query for an upstream gradient ``d_out`` [n,13] on the network outputs at the arg-min root
        (channel 0 = sdf): dict with ``hash``, ``w1``, ``b1``, ``w2``, ``b2`` (as ``op_geometry_backward``), ``tfs`` [24,3,4]
        and ``x`` [n,3].  ``fwd``: what ``op_query_train`` returned."""
        xc = fwd["x_c"].contiguous()
        n = xc.shape[0]
        valid = fwd["valid"].to(torch.uint8).contiguous()
        d_out = d_out.to(self.dev, torch.float32).reshape(n, 13).contiguous()
        g_hash = torch.zeros_like(self._keep["geo"])
        g_mlp = torch.zeros(3152, device=self.dev)
        g_tfs = torch.zeros(24, 3, 4, device=self.dev)
        g_x = torch.empty(n, 3, device=self.dev)
        check(self.lib.ia_op_query_backward(self.h, ptr(xc), ptr(valid), ptr(fwd["J_inv"].contiguous()), ptr(d_out), n,
                                            ptr(g_hash), ptr(g_mlp), ptr(g_tfs), ptr(g_x), _stream()), "ia_op_query_backward")
        w1t, b1 = g_mlp[:35 * 64].reshape(35, 64), g_mlp[35 * 64:36 * 64]
        w2, b2 = g_mlp[36 * 64:49 * 64].reshape(13, 64), g_mlp[49 * 64:49 * 64 + 13]
        return {"hash": g_hash, "w1": w1t.t().contiguous(), "b1": b1, "w2": w2, "b2": b2, "tfs": g_tfs, "x": g_x}

    def op_traverse(self, rays_o, rays_d, near, far, step):
        o = rays_o.to(self.dev, torch.float32).contiguous()
        d = rays_d.to(self.dev, torch.float32).contiguous()
        n = o.shape[0]
        ne = torch.zeros(n, dtype=torch.int32, device=self.dev)
        ns = torch.zeros(n, dtype=torch.int32, device=self.dev)
        z = C.c_void_p(0)
        check(self.lib.ia_op_traverse(self.h, ptr(o), ptr(d), n, near, far, step, ptr(ne), ptr(ns), z, z, z, z, z, z, z,
                                      _stream()), "ia_op_traverse")
        eb = (torch.cumsum(ne, 0) - ne).int()
        sb = (torch.cumsum(ns, 0) - ns).int()
        E, S = int(ne.sum()), int(ns.sum())
        vals = torch.zeros(E, device=self.dev)
        il = torch.zeros(E, dtype=torch.uint8, device=self.dev)
        ir = torch.zeros(E, dtype=torch.uint8, device=self.dev)
        ts, te = torch.zeros(S, device=self.dev), torch.zeros(S, device=self.dev)
        check(self.lib.ia_op_traverse(self.h, ptr(o), ptr(d), n, near, far, step, ptr(ne), ptr(ns), ptr(eb), ptr(sb),
                                      ptr(vals), ptr(il), ptr(ir), ptr(ts), ptr(te), _stream()), "ia_op_traverse")
        return {"vals": vals, "is_left": il.bool(), "is_right": ir.bool(), "packed_info": torch.stack([eb, ne], 1),
                "t_starts": ts, "t_ends": te, "sample_packed_info": torch.stack([sb, ns], 1)}

    @staticmethod
    def _rpacked(num_steps, add):
        rs = add.int()
        cum = torch.cumsum(rs, 0).int()
        return torch.stack([cum - rs, rs], 1).contiguous(), int(cum[-1]) if len(cum) else 0

    def op_ray_resampling(self, packed_info, starts, ends, weights, sdfs, n_samples):
        pi = packed_info.to(self.dev).int().contiguous()
        a = [t.to(self.dev, torch.float32).reshape(-1).contiguous() for t in (starts, ends, weights, sdfs)]
        n = pi.shape[0]
        rpi, total = self._rpacked(pi[:, 1], (pi[:, 1] > 0) * n_samples)
        ts = torch.zeros(total, 1, device=self.dev)
        offs = torch.zeros(total, 1, device=self.dev)
        idx = torch.zeros(total, dtype=torch.int64, device=self.dev)
        fg = torch.zeros(a[2].numel(), dtype=torch.int32, device=self.dev)
        bg = torch.zeros(n, dtype=torch.int32, device=self.dev)
        surf = torch.zeros(n, dtype=torch.int64, device=self.dev)
        check(self.lib.ia_op_ray_resampling(self.h, ptr(pi), *[ptr(t) for t in a], n, n_samples, ptr(rpi), ptr(ts),
                                            ptr(offs), ptr(idx), ptr(fg), ptr(bg), ptr(surf), _stream()),
              "ia_op_ray_resampling")
        return rpi, ts, offs, idx, fg, bg, surf

    def op_ray_resampling_merge(self, packed_info, vals, is_left, is_right, weights, n_samples):
        pi = packed_info.to(self.dev).int().contiguous()
        v = vals.to(self.dev, torch.float32).contiguous()
        w = weights.to(self.dev, torch.float32).contiguous()
        il = is_left.to(self.dev).to(torch.uint8).contiguous()
        ir = is_right.to(self.dev).to(torch.uint8).contiguous()
        n = pi.shape[0]
        rpi, total = self._rpacked(pi[:, 1], (pi[:, 1] > 0) * n_samples + pi[:, 1])
        rv, rd = torch.zeros(total, device=self.dev), torch.zeros(total, device=self.dev)
        o = [torch.zeros(total, dtype=torch.uint8, device=self.dev) for _ in range(4)]
        check(self.lib.ia_op_ray_resampling_merge(self.h, ptr(pi), ptr(v), ptr(il), ptr(ir), ptr(w), n, ptr(rpi), ptr(rv),
                                                  ptr(rd), *[ptr(t) for t in o], _stream()), "ia_op_ray_resampling_merge")
        return (rpi, rv, rd) + tuple(t.bool() for t in o)

    def op_ray_resampling_sdf_fine(self, packed_info, starts, ends, alphas, sdfs, n_samples):
        pi = packed_info.to(self.dev).int().contiguous()
        a = [t.to(self.dev, torch.float32).reshape(-1).contiguous() for t in (starts, ends, alphas, sdfs)]
        n = pi.shape[0]
        rpi, total = self._rpacked(pi[:, 1], (pi[:, 1] > 0) * n_samples)
        rs, re = torch.zeros(total, 1, device=self.dev), torch.zeros(total, 1, device=self.dev)
        fg = torch.zeros(total, dtype=torch.uint8, device=self.dev)
        check(self.lib.ia_op_ray_resampling_sdf_fine(self.h, ptr(pi), *[ptr(t) for t in a], n, ptr(rpi), ptr(rs), ptr(re),
                                                     ptr(fg), _stream()), "ia_op_ray_resampling_sdf_fine")
        return rpi, rs, re, fg.bool()

    def op_ray_resampling_fine(self, packed_info, starts, ends, weights, n_samples):
        pi = packed_info.to(self.dev).int().contiguous()
        a = [t.to(self.dev, torch.float32).reshape(-1).contiguous() for t in (starts, ends, weights)]
        n = pi.shape[0]
        rpi, total = self._rpacked(pi[:, 1], (pi[:, 1] > 0) * n_samples)
        rs, re = torch.zeros(total, 1, device=self.dev), torch.zeros(total, 1, device=self.dev)
        fg = torch.zeros(total, dtype=torch.uint8, device=self.dev)
        check(self.lib.ia_op_ray_resampling_fine(self.h, ptr(pi), *[ptr(t) for t in a], n, ptr(rpi), ptr(rs), ptr(re), ptr(fg),
                                                 _stream()), "ia_op_ray_resampling_fine")
        return rpi, rs, re, fg.bool()

    def op_unpack_info(self, packed_info, n_samples):
        pi = packed_info.to(self.dev).int().contiguous()
        out = torch.zeros(n_samples, dtype=torch.int64, device=self.dev)
        check(self.lib.ia_op_unpack_info(self.h, ptr(pi), pi.shape[0], ptr(out), _stream()), "ia_op_unpack_info")
        return out

    def op_secondary(self, o, d, gi=False):
        o = o.to(self.dev, torch.float32).contiguous()
        d = d.to(self.dev, torch.float32).contiguous()
        n = o.shape[0]
        T, rgb = torch.empty(n, device=self.dev), torch.empty(n, 3, device=self.dev)
        check(self.lib.ia_op_secondary(self.h, ptr(o), ptr(d), n, int(gi), ptr(T), ptr(rgb), _stream()), "ia_op_secondary")
        return T, rgb

    def op_brdf(self, wi, n, wo, rough, albedo, metal):
        a = [t.to(self.dev, torch.float32).contiguous() for t in (wi, n, wo, rough, albedo, metal)]
        m = a[0].shape[0]
        diff, spec = torch.empty(m, device=self.dev), torch.empty(m, 3, device=self.dev)
        check(self.lib.ia_op_brdf(self.h, *[ptr(t) for t in a], m, ptr(diff), ptr(spec), _stream()), "ia_op_brdf")
        return diff, spec

    def op_bsdf_sample_pdf(self, wi, n, rough, albedo, metal, sample=None, wo_query=None):
        """MultiLobe.sample (explicit uniforms) and MultiLobe.pdf: returns (wo or None, pdf or None)."""
        a = [t.to(self.dev, torch.float32).contiguous() for t in (wi, n, rough, albedo, metal)]
        m = a[0].shape[0]
        smp = sample.to(self.dev, torch.float32).contiguous() if sample is not None else None
        woq = wo_query.to(self.dev, torch.float32).contiguous() if wo_query is not None else None
        wo = torch.empty(m, 3, device=self.dev) if smp is not None else None
        pdf = torch.empty(m, device=self.dev) if (woq is not None or wo is not None) else None
        check(self.lib.ia_op_bsdf_sample_pdf(self.h, *[ptr(t) for t in a], ptr(smp), ptr(woq), m, ptr(wo), ptr(pdf),
                                             _stream()), "ia_op_bsdf_sample_pdf")
        return wo, pdf

    def op_env(self, u=None, dirs_world=None):
        """EnvironmentLightTensor.sample / pdf / eval per direction on the tables of the last set_light call."""
        uu = u.to(self.dev, torch.float32).contiguous() if u is not None else None
        dd = dirs_world.to(self.dev, torch.float32).contiguous() if dirs_world is not None else None
        m = (uu if uu is not None else dd).shape[0]
        dout = torch.empty(m, 3, device=self.dev) if uu is not None else None
        pdf, em = torch.empty(m, device=self.dev), torch.empty(m, 3, device=self.dev)
        check(self.lib.ia_op_env(self.h, ptr(uu), ptr(dd), m, ptr(dout), ptr(pdf), ptr(em), _stream()), "ia_op_env")
        return dout, pdf, em
