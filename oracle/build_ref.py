"""ORACLE (test infrastructure): compile the reference's OWN first-party CUDA extensions, from the
sources where they lie under /root/reference, into ``oracle/_ref/`` (git-ignored, travels to the GPU
box with the snapshot).  They are the strongest checker available for the op-level A/B tests
(tests/test_gpu_ref_ab.py): the product's kernels against the reference's kernels on one GPU.

No reference source is copied; the four extensions are built with ``torch.utils.cpp_extension.load``
exactly as the reference does (deformer_torch.py:8-18; lib/nerfacc/cuda/_backend.py:40-76), only with
an explicit sm_100a gencode and an in-tree build directory.

  fuse_cuda   <- models/deformers/fast_snarf/cuda/fuse_kernel/{fuse_cuda.cpp,fuse_cuda_kernel_fast.cu}
  filter      <- models/deformers/fast_snarf/cuda/filter/{filter.cpp,filter.cu}
  precompute  <- models/deformers/fast_snarf/cuda/precompute/{precompute.cpp,precompute.cu}
  nerfacc_cuda<- lib/nerfacc/cuda/csrc/{cdf.cu,pack.cu,pybind.cu}
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

_SNARF = os.path.join(REF, "models/deformers/fast_snarf/cuda")
EXTENSIONS = {
    "fuse_cuda": [f"{_SNARF}/fuse_kernel/fuse_cuda.cpp", f"{_SNARF}/fuse_kernel/fuse_cuda_kernel_fast.cu"],
    "filter": [f"{_SNARF}/filter/filter.cpp", f"{_SNARF}/filter/filter.cu"],
    "precompute": [f"{_SNARF}/precompute/precompute.cpp", f"{_SNARF}/precompute/precompute.cu"],
    "nerfacc_cuda": sorted(glob.glob(os.path.join(REF, "lib/nerfacc/cuda/csrc/*.cu"))),
}


def so_path(name: str) -> str:
    return os.path.join(OUT, name, name + ".so")


def build_all(verbose: bool = False) -> None:
    """Build whatever is missing.  Needs /root/reference (this container only)."""
    from torch.utils.cpp_extension import load
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    for name, srcs in EXTENSIONS.items():
        if os.path.exists(so_path(name)):
            continue
        if not srcs or not all(os.path.exists(s) for s in srcs):
            raise FileNotFoundError(f"reference sources for {name} not found under {REF}")
        bd = os.path.join(OUT, name)
        os.makedirs(bd, exist_ok=True)
        stale = os.path.join(bd, "lock")  # left behind by an interrupted build; load() would wait on it forever
        if os.path.exists(stale):
            os.remove(stale)
        load(name=name, sources=srcs, build_directory=bd, extra_cflags=["-O3"], extra_cuda_cflags=["-O3"],
             verbose=verbose, is_python_module=False)  # compile + link only; nothing is imported here


def load_ref(name: str):
    """Import a prebuilt reference extension (GPU box or here); None if it was never built."""
    p = so_path(name)
    if not os.path.exists(p):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules.setdefault(name, mod)
    return mod


if __name__ == "__main__":
    build_all(verbose=True)
    print({n: os.path.exists(so_path(n)) for n in EXTENSIONS})
