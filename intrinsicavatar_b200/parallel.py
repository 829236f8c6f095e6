"""Multi-GPU: frames are independent, so the path shards by frame with no data-path exchange
(SURVEY.md section 8e).  One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in
the CPU tests); the only collective is the final gather of finished frame buffers to rank 0.

Two transports for that gather (``FrameCollector``):
  * ``nccl``  -- one ``dist.gather`` per step, posted asynchronously.  Its send / receive kernels need SMs and rendezvous:
    while the persistent shading kernel owns every SM of a rank, a peer's send kernel sits resident (holding SMs of ITS
    GPU) until the receive can start -- the 1 -> 8 GPU loss of round 1.  NCCL is therefore
    initialised with one CTA per peer-to-peer operation (NCCL_MAX_CTAS=1) when this transport is used.
  * ``p2p``   -- rank 0 exports its receive buffer through CUDA IPC, every rank maps it and writes its finished frame into
    its slot with a stream-ordered peer copy (copy engines over NVLink: no SM, no rendezvous, nothing for rank 0 to
    do).  Completion = each rank's stream sync + the closing barrier.  Falls back to ``nccl`` if the mapping fails.

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


def frame_of_step(step: int, rank: int, world_size: int, n_frames: int) -> int:
    """Frame rendered by ``rank`` at ``step`` of a pass over a sequence of ``n_frames``: step s covers the world_size
    consecutive frames [s * world_size, (s + 1) * world_size) (mod n_frames) -- every frame of the sequence exactly once per
    pass, frame f -> some rank, BASELINE configs[4] -- and the assignment inside a step rotates by one rank per step, so that
    over a pass every rank renders every residue class of frames (frames differ in cost by +-15 %; a fixed f -> f mod N
    assignment would measure that spread instead of the implementation)."""
    return (step * world_size + (rank + step) % world_size) % n_frames


def assign_frames(frames: list[int], costs: dict, world_size: int) -> list[list[int]]:
    """Static longest-first assignment of a batch of frames to ``world_size`` ranks, the same number of frames per rank
    (len(frames) must be a multiple of world_size): frames in order of decreasing estimated cost, each to the rank with the
    smallest sum so far that still has room (LPT with a cardinality bound), so that the per-rank sums of the estimates are
    nearly equal.  ``costs[f]`` is any estimate proportional to the render time of frame f (here: measured ms of an earlier
    pass over the sequence; in a deployment the hit-pixel count of the primary stage serves).  Deterministic: every rank
    computes the same table, no communication.  Returns per rank the list of its frames, most expensive first."""
    assert len(frames) % world_size == 0
    per = len(frames) // world_size
    order = sorted(range(len(frames)), key=lambda i: (-float(costs[frames[i]]), i))
    out = [[] for _ in range(world_size)]
    sums = [0.0] * world_size
    for i in order:
        r = min((r for r in range(world_size) if len(out[r]) < per), key=lambda r: (sums[r], r))
        out[r].append(frames[i])
        sums[r] += float(costs[frames[i]])
    return out


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


class FrameCollector:
    """Collects the packed frame of every rank and step on rank ``dst``.

    ``slots`` frames per rank are kept (slot = step mod slots); ``collect(step, block)`` never blocks the host,
    ``finish()`` makes every posted frame visible on ``dst`` (call it before reading ``frames(step)``)."""

    def __init__(self, n_pix: int, slots: int, device, transport: str = "p2p", dst: int = 0):
        assert transport in ("p2p", "nccl")
        self.ws = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst, self.slots, self.n_pix, self.device = dst, slots, n_pix, device
        self.transport = transport if self.ws > 1 else "local"
        self.pending = []
        self.recv = None        # dst: [slots, ws, n_pix, 16]
        self.peer = None        # p2p: view of dst's buffer in this process
        if self.transport == "p2p":
            ok = self._setup_p2p()
            flag = torch.tensor([1 if ok else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.transport, self.peer = "nccl", None
        if self.transport in ("nccl", "local") and self.rank == dst:
            if self.recv is None:
                self.recv = torch.empty(slots, self.ws, n_pix, 16, device=device)

    def _setup_p2p(self) -> bool:
        """rank dst allocates [slots, ws, n_pix, 16] and shares it through CUDA IPC (the mechanism torch.multiprocessing
        uses to pass CUDA tensors between processes); the others open the handle and get a device pointer into dst's
        memory that peer copies can target over NVLink."""
        try:
            shape = (self.slots, self.ws, self.n_pix, 16)
            if self.rank == self.dst:
                self.recv = torch.empty(shape, device=self.device)
                meta = [self.recv.untyped_storage()._share_cuda_()]
            else:
                meta = [None]
            dist.broadcast_object_list(meta, src=self.dst)
            if self.rank == self.dst:
                self.peer = self.recv
            else:
                (dev, handle, size, offset, ref_handle, ref_offset, ev_handle, ev_sync) = meta[0]
                st = torch.UntypedStorage._new_shared_cuda(torch.cuda.current_device(), handle, size, offset, ref_handle,
                                                           ref_offset, ev_handle, ev_sync)
                self._peer_storage = st
                self.peer = torch.empty(0, dtype=torch.float32, device=self.device).set_(st, 0, shape)
                # one small write proves the mapping before the timed region relies on it
                self.peer[0, self.rank, 0, :1].copy_(torch.zeros(1, device=self.device))
                torch.cuda.current_stream().synchronize()
            return True
        except Exception as e:  # no peer access / IPC unavailable: use the NCCL gather
            print(f"[parallel] rank {self.rank}: p2p frame transport unavailable ({type(e).__name__}: {e}); using nccl",
                  flush=True)
            return False

    def collect(self, step: int, block: torch.Tensor):
        slot = step % self.slots
        if self.transport == "local":
            self.recv[slot, 0].copy_(block, non_blocking=True)
        elif self.transport == "p2p":
            self.peer[slot, self.rank].copy_(block, non_blocking=True)     # stream-ordered peer copy, no SM
        else:
            # a slot is reused only after the gather that filled it last has completed (NCCL orders collectives on its
            # stream, gloo runs asynchronous work on a thread pool and may finish two gathers into one slot in either order)
            for s_, _, w_ in self.pending:
                if s_ == slot:
                    w_.wait()
            self.pending = [e for e in self.pending if e[0] != slot]
            bufs = [self.recv[slot, r] for r in range(self.ws)] if self.rank == self.dst else None
            work = dist.gather(block, gather_list=bufs, dst=self.dst, async_op=True)
            self.pending.append((slot, block, work))

    def finish(self):
        for _, _, work in self.pending:
            work.wait()
        self.pending.clear()
        if torch.device(self.device).type == "cuda":
            torch.cuda.current_stream().synchronize()
        if self.ws > 1:
            dist.barrier()

    def close(self):
        """Drop the mapping of dst's buffer everywhere BEFORE dst frees it (CUDA IPC: the producer must outlive its consumers)."""
        if self.transport == "p2p" and self.rank != self.dst:
            self.peer = None
            self._peer_storage = None
            import gc
            gc.collect()
            torch.cuda.ipc_collect()      # hand the consumer-side references back now, not at interpreter exit
        if self.ws > 1:
            if torch.device(self.device).type == "cuda":
                torch.cuda.synchronize()
            dist.barrier()
        self.recv = None

    def frames(self, step: int):
        """dst only: [ws, n_pix, 16] blocks of ``step`` (after ``finish``)."""
        return self.recv[step % self.slots] if self.rank == self.dst else None
