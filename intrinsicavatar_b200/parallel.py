"""Multi-GPU: frames are independent, so the path shards by frame with no data-path exchange
(SURVEY.md section 8e).  One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in
the CPU tests); the only collective is the final gather of finished frame buffers to rank 0.

The reference has no image gather (each Lightning rank writes its own PNGs,
systems/intrinsic_avatar.py:846-864) and only all_gathers per-frame metric scalars (:566, :881).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

# buffers gathered per frame (config 5): 3+3+3+1+1+1+1+3 floats per pixel = 16 floats
FRAME_KEYS = ("comp_rgb_phys_full", "comp_albedo_full", "comp_normal", "opacity", "depth", "comp_roughness_full",
              "comp_metallic_full", "comp_rgb_full")


def frames_for_rank(n_frames: int, rank: int, world_size: int) -> list[int]:
    """Frame i -> rank i mod world_size (DistributedSampler-style round robin)."""
    return list(range(rank, n_frames, world_size))


def pack_frame(out: dict) -> torch.Tensor:
    """[n_pix, 16] float32 image-buffer block of one frame."""
    return torch.cat([out[k].float() for k in FRAME_KEYS], dim=-1).contiguous()


def unpack_frame(block: torch.Tensor) -> dict:
    sizes = (3, 3, 3, 1, 1, 1, 1, 3)
    parts = torch.split(block, sizes, dim=-1)
    return dict(zip(FRAME_KEYS, parts))


def gather_frames(block: torch.Tensor, dst: int = 0):
    """Gather one packed frame per rank to ``dst``.  Returns the list of blocks on dst, None elsewhere.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    bufs, work = gather_frames_async(block, dst)
    if work is not None:
        work.wait()
    return bufs


def gather_frames_async(block: torch.Tensor, dst: int = 0):
    """Post the gather without waiting for it: returns (list of receive blocks on dst / None elsewhere, work handle or
    None).  Frames take different times (0.8-1.1 s here), so a rank -- rank ``dst`` above all -- should go on rendering
    its next frame instead of waiting for the slowest rank of this step; call ``work.wait()`` when the blocks are
    needed.  ``block`` (and the returned list) must stay alive until then."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [block], None
    ws, rank = dist.get_world_size(), dist.get_rank()
    bufs = [torch.empty_like(block) for _ in range(ws)] if rank == dst else None
    work = dist.gather(block, gather_list=bufs, dst=dst, async_op=True)
    return bufs, work
