"""ctypes binding of ``libia_b200.so`` (include/ia_b200.h).  No torch types cross the ABI:
tensors are passed as ``data_ptr()`` integers and the current CUDA stream handle.

The library is the product; there is no fallback.  ``load()`` raises if it is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# IA_B200_LIB selects an alternative build of the same library (tuning variants made by build(defines=...))
LIB_PATH = os.environ.get("IA_B200_LIB") or os.path.join(_HERE, "libia_b200.so")
SRC_DIR = os.path.join(_HERE, "csrc")
N_COUNTERS = 32
MAX_SAMPLES_PER_RAY = 256   # IA_CAP (csrc/ia_types.cuh): edges / samples one primary ray can hold
CNT_PRIMARY_BASE = 16
COUNTER_NAMES = ["hit_rays", "samples", "queries", "queries_grad", "broyden_fetch", "geo_eval", "rad_eval",
                 "secondary_rays", "overflow", "skin_fetch", "chains_skipped"]

STAGE_NAMES = ["precompute", "occupancy", "light", "setup", "primary", "resample", "shade", "composite"]

RENDER_PRIMARY_ONLY = 1
RENDER_GI = 2
# config.model.render_mode -> bits 2-3 of the ia_render flags (IA_RENDER_* in include/ia_b200.h)
RENDER_ADD_EMITTER = 16
RENDER_MODES = {"light": 0 << 2, "uniform_light": 1 << 2, "mats": 2 << 2, "mis": 3 << 2}


class IaOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "comp_rgb", "comp_normal", "opacity", "depth", "comp_albedo", "comp_roughness", "comp_metallic",
        "comp_rgb_phys", "comp_demod_phys", "num_samples", "comp_rgb_full", "comp_rgb_phys_full",
        "comp_demod_phys_full", "comp_albedo_full", "comp_roughness_full", "comp_metallic_full", "visibility")]


EXPORTS = [
    "ia_last_error", "ia_version", "ia_voxel_format", "ia_create", "ia_destroy", "ia_set_fields", "ia_set_lbs_voxels", "ia_smpl_lbs", "ia_voxelize_lbs", "ia_set_pose",
    "ia_set_render_config", "ia_set_secondary_sampling", "ia_reserve_samples", "ia_build_occupancy", "ia_set_occupancy", "ia_set_light", "ia_set_light_uniform", "ia_render", "ia_get_counters", "ia_set_timing", "ia_get_timings",
    "ia_op_precompute", "ia_op_broyden", "ia_op_query", "ia_op_shade_fields", "ia_op_geometry", "ia_op_geometry_backward", "ia_op_deform_backward", "ia_op_shade_fields_backward", "ia_op_volrend", "ia_op_volrend_backward", "ia_op_query_train", "ia_op_query_backward", "ia_op_traverse",
    "ia_op_ray_resampling", "ia_op_ray_resampling_merge", "ia_op_ray_resampling_sdf_fine", "ia_op_ray_resampling_fine", "ia_op_unpack_info",
    "ia_op_secondary", "ia_op_brdf", "ia_op_bsdf_sample_pdf", "ia_op_env", "ia_op_pbr_shade", "ia_op_pbr_shade_backward", "ia_op_env_backward",
    "ia_make_rays", "ia_pack_rgb8", "ia_pack_grid8", "ia_update_occupancy_ema",
]


def build(force: bool = False, verbose: bool = False, defines: dict | None = None, out: str | None = None) -> str:
    """nvcc -> libia_b200.so for sm_100a (cross-compiles without a GPU).  ``defines`` / ``out`` build a
    tuning variant next to the default library."""
    if defines or out:
        out = out or LIB_PATH
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
               "-Xcompiler", "-fPIC", "-o", out, os.path.join(SRC_DIR, "ia_kernels.cu")]
        cmd += [f"-D{k}={v}" for k, v in (defines or {}).items()]
        subprocess.check_call(cmd)
        return out
    srcs = [os.path.join(SRC_DIR, f) for f in sorted(os.listdir(SRC_DIR))]
    hdr = os.path.join(_HERE, "..", "include", "ia_b200.h")
    newest = max(os.path.getmtime(p) for p in srcs + [hdr])
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-o", LIB_PATH, os.path.join(SRC_DIR, "ia_kernels.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU or PyTorch fallback for the render path.")
    lib = C.CDLL(LIB_PATH)
    lib.ia_last_error.restype = C.c_char_p
    for name in EXPORTS:
        getattr(lib, name)  # raises AttributeError if a declared symbol is not exported
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ia_last_error().decode()
        raise RuntimeError(f"libia_b200 {what} failed (code {rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_contiguous(), "libia_b200 takes contiguous buffers"
    return C.c_void_p(t.data_ptr())


def fptr(a):
    """Host float32/int32 numpy array -> pointer."""
    return a.ctypes.data_as(C.c_void_p)
