"""ORACLE (test infrastructure): Python faces of the serial C restatements, with the reference's
own op signatures (lib/nerfacc/cdf.py:12-244, lib/nerfacc/pack.py:46-190) on CPU torch tensors,
plus nerfacc's ``traverse_grids`` / ``render_weight_from_alpha`` / ``accumulate_along_rays``
(third-party, parity unpinned -- see oracle/serial_ops.c header).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import serial_lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _np(t, dtype):
    if torch.is_tensor(t):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _resample_packed_info(num_steps: np.ndarray, add: np.ndarray):
    rs = add.astype(np.int32)
    cum = np.cumsum(rs, dtype=np.int32)
    return np.stack([cum - rs, rs], axis=1).astype(np.int32), int(cum[-1]) if len(cum) else 0


def pack_info(ray_indices: torch.Tensor, n_rays: int) -> torch.Tensor:
    """lib/nerfacc/pack.py:46-76."""
    num = torch.zeros(n_rays, dtype=torch.int32)
    num.scatter_add_(0, ray_indices.long(), torch.ones_like(ray_indices, dtype=torch.int32))
    cum = num.cumsum(0, dtype=torch.int32)
    return torch.stack([cum - num, num], dim=-1)


def unpack_info(packed_info: torch.Tensor, n_samples: int) -> torch.Tensor:
    pi = _np(packed_info, np.int32)
    out = np.zeros(n_samples, dtype=np.int64)
    serial_lib().unpack_info(ctypes.c_int(pi.shape[0]), _p(pi), _p(out))
    return torch.from_numpy(out)


def unpack_data(packed_info: torch.Tensor, data: torch.Tensor, n_per_ray: int) -> torch.Tensor:
    pi = _np(packed_info, np.int32)
    d = _np(data, np.float32)
    out = np.zeros((pi.shape[0], n_per_ray, d.shape[1]), dtype=np.float32)
    serial_lib().unpack_data_f32(ctypes.c_int(pi.shape[0]), _p(pi), ctypes.c_int(d.shape[1]), _p(d),
                                 ctypes.c_int(n_per_ray), _p(out))
    return torch.from_numpy(out)


def ray_resampling(packed_info, starts, ends, weights, sdfs, n_samples: int):
    """lib/nerfacc/cdf.py:12-76 / cdf.cu:151-215. starts/ends [n,1]."""
    assert n_samples > 1
    pi = _np(packed_info, np.int32)
    st, en = _np(starts, np.float32).reshape(-1), _np(ends, np.float32).reshape(-1)
    w, s = _np(weights, np.float32), _np(sdfs, np.float32)
    n_rays = pi.shape[0]
    rpi, total = _resample_packed_info(pi[:, 1], (pi[:, 1] > 0) * n_samples)
    ts = np.zeros((total, 1), np.float32)
    offs = np.zeros((total, 1), np.float32)
    idx = np.zeros(total, np.int64)
    surf = -np.ones(n_rays, np.int64)
    fg = np.zeros(len(w), np.int32)
    bg = np.zeros(n_rays, np.int32)
    serial_lib().cdf_resampling(ctypes.c_int(n_rays), _p(pi), _p(st), _p(en), _p(w), _p(s), _p(rpi),
                                _p(ts), _p(offs), _p(surf), _p(idx), _p(fg), _p(bg))
    return tuple(torch.from_numpy(a) for a in (rpi, ts, offs, idx, fg, bg, surf))


def ray_resampling_merge(packed_info, vals, is_left, is_right, weights, n_samples: int):
    """lib/nerfacc/cdf.py:79-141 / cdf.cu:336-401."""
    pi = _np(packed_info, np.int32)
    v, w = _np(vals, np.float32), _np(weights, np.float32)
    il, ir = _np(is_left, np.uint8), _np(is_right, np.uint8)
    n_rays = pi.shape[0]
    rpi, total = _resample_packed_info(pi[:, 1], (pi[:, 1] > 0) * n_samples + pi[:, 1])
    rv = np.zeros(total, np.float32)
    rd = np.zeros(total, np.float32)
    ril = np.zeros(total, np.uint8)
    rir = np.zeros(total, np.uint8)
    isr = np.zeros(total, np.uint8)
    isfg = np.zeros(total, np.uint8)
    serial_lib().cdf_resampling_merge(ctypes.c_int(n_rays), _p(pi), _p(v), _p(il), _p(ir), _p(w), _p(rpi),
                                      _p(rv), _p(rd), _p(ril), _p(rir), _p(isr), _p(isfg))
    return (torch.from_numpy(rpi), torch.from_numpy(rv), torch.from_numpy(rd),
            torch.from_numpy(ril).bool(), torch.from_numpy(rir).bool(),
            torch.from_numpy(isr).bool(), torch.from_numpy(isfg).bool())


def _fine(fn_name, packed_info, starts, ends, a, sdfs, n_samples):
    pi = _np(packed_info, np.int32)
    st, en = _np(starts, np.float32).reshape(-1), _np(ends, np.float32).reshape(-1)
    a = _np(a, np.float32)
    n_rays = pi.shape[0]
    rpi, total = _resample_packed_info(pi[:, 1], (pi[:, 1] > 0) * n_samples)
    rs = np.zeros((total, 1), np.float32)
    re = np.zeros((total, 1), np.float32)
    isfg = np.zeros(total, np.uint8)
    if sdfs is None:
        serial_lib().cdf_resampling_fine(ctypes.c_int(n_rays), _p(pi), _p(st), _p(en), _p(a), _p(rpi),
                                         _p(rs), _p(re), _p(isfg))
    else:
        s = _np(sdfs, np.float32)
        serial_lib().cdf_resampling_sdf_fine(ctypes.c_int(n_rays), _p(pi), _p(st), _p(en), _p(a), _p(s),
                                             _p(rpi), _p(rs), _p(re), _p(isfg))
    return torch.from_numpy(rpi), torch.from_numpy(rs), torch.from_numpy(re), torch.from_numpy(isfg).bool()


def ray_resampling_fine(packed_info, starts, ends, weights, n_samples: int):
    """lib/nerfacc/cdf.py:199-244 / cdf.cu:480-534."""
    return _fine("fine", packed_info, starts, ends, weights, None, n_samples)


def ray_resampling_sdf_fine(packed_info, starts, ends, alphas, sdfs, n_samples: int):
    """lib/nerfacc/cdf.py:144-196 / cdf.cu:640-696."""
    return _fine("sdf_fine", packed_info, starts, ends, alphas, sdfs, n_samples)


def render_weight_from_alpha(alphas: torch.Tensor, packed_info: torch.Tensor):
    pi = _np(packed_info, np.int32)
    a = _np(alphas, np.float32)
    w = np.zeros_like(a)
    T = np.zeros_like(a)
    serial_lib().render_weight_from_alpha(ctypes.c_int(pi.shape[0]), _p(pi), _p(a), _p(w), _p(T))
    return torch.from_numpy(w), torch.from_numpy(T)


def accumulate_along_rays(weights, values, ray_indices, n_rays):
    if values is None:
        src = weights[:, None]
    else:
        src = weights[:, None] * values
    out = torch.zeros(n_rays, src.shape[-1], dtype=torch.float32)
    out.index_add_(0, ray_indices.long(), src)
    return out


def traverse_grid(rays_o, rays_d, binaries, aabb, near_plane: float, far_plane: float, step_size: float):
    """Single-level nerfacc ``traverse_grids`` as used by ``sampling_override``
    (models/intrinsic_avatar.py:49-141).  Returns dict with the RayIntervals fields and samples."""
    o, d = _np(rays_o, np.float32), _np(rays_d, np.float32)
    b = _np(binaries, np.uint8).reshape(-1)
    res = round(len(b) ** (1 / 3))
    bb = _np(aabb, np.float32).reshape(6)
    n = o.shape[0]
    ne = np.zeros(n, np.int32)
    ns = np.zeros(n, np.int32)
    lib = serial_lib()
    args = (ctypes.c_int(n), _p(o), _p(d), _p(b), ctypes.c_int(res), _p(bb), ctypes.c_float(near_plane),
            ctypes.c_float(far_plane), ctypes.c_float(step_size), _p(ne), _p(ns))
    lib.traverse_grid(*args, None, None, None, None, None, None, None)
    eb = (np.cumsum(ne) - ne).astype(np.int32)
    sb = (np.cumsum(ns) - ns).astype(np.int32)
    E, S = int(ne.sum()), int(ns.sum())
    vals = np.zeros(E, np.float32)
    il = np.zeros(E, np.uint8)
    ir = np.zeros(E, np.uint8)
    ts = np.zeros(S, np.float32)
    te = np.zeros(S, np.float32)
    ne2, ns2 = np.zeros_like(ne), np.zeros_like(ns)
    args2 = args[:9] + (_p(ne2), _p(ns2))
    lib.traverse_grid(*args2, _p(eb), _p(sb), _p(vals), _p(il), _p(ir), _p(ts), _p(te))
    ray_idx_e = np.repeat(np.arange(n, dtype=np.int64), ne)
    ray_idx_s = np.repeat(np.arange(n, dtype=np.int64), ns)
    return {
        "vals": torch.from_numpy(vals), "is_left": torch.from_numpy(il).bool(),
        "is_right": torch.from_numpy(ir).bool(), "ray_indices": torch.from_numpy(ray_idx_e),
        "packed_info": torch.from_numpy(np.stack([eb, ne], 1)),
        "t_starts": torch.from_numpy(ts), "t_ends": torch.from_numpy(te),
        "sample_ray_indices": torch.from_numpy(ray_idx_s),
        "sample_packed_info": torch.from_numpy(np.stack([sb, ns], 1)),
    }
