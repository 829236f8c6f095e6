"""CPU (-m "not gpu"): the N>1 path with world_size 2 over gloo -- frame partition, pack / gather / unpack."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from intrinsicavatar_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.frames_for_rank(5, rank, world)
        n_pix = 6
        out = {}
        for k, c in zip(parallel.FRAME_KEYS, (3, 3, 3, 1, 1, 1, 1, 3)):
            out[k] = torch.full((n_pix, c), float(rank * 100 + c)) + torch.arange(n_pix)[:, None]
        block = parallel.pack_frame(out)
        got = parallel.gather_frames(block, dst=0)
        ok = True
        if rank == 0:
            ok = got is not None and len(got) == world
            for r, b in enumerate(got):
                un = parallel.unpack_frame(b)
                for k, c in zip(parallel.FRAME_KEYS, (3, 3, 3, 1, 1, 1, 1, 3)):
                    ok = ok and un[k].shape == (n_pix, c) and float(un[k][0, 0]) == r * 100 + c
        else:
            ok = got is None
        # asynchronous form (bench.py): two steps posted back to back, waited for afterwards, order preserved
        posted = []
        for step in range(2):
            blk = block + 1000.0 * (step + 1)
            posted.append((blk,) + parallel.gather_frames_async(blk, dst=0))
        for step, (blk, bufs, work) in enumerate(posted):
            ok = ok and work is not None
            work.wait()
            if rank == 0:
                ok = ok and all(float(bufs[r][0, 0]) == r * 100 + 3 + 1000.0 * (step + 1) for r in range(world))
            else:
                ok = ok and bufs is None
        # FrameCollector, gather transport (the p2p transport needs CUDA IPC: covered on the GPU box by bench.py --gpus 2)
        col = parallel.FrameCollector(n_pix, slots=2, device="cpu", transport="nccl", dst=0)
        for step in range(3):
            col.collect(step, block + 10.0 * step)
        col.finish()
        if rank == 0:
            for step in (1, 2):        # slot of step 0 was reused by step 2
                fr = col.frames(step)
                ok = ok and fr.shape == (world, n_pix, 16)
                ok = ok and all(float(fr[r, 0, 0]) == r * 100 + 3 + 10.0 * step for r in range(world))
        else:
            ok = ok and col.frames(1) is None
        t = torch.tensor([float(len(mine))])
        dist.all_reduce(t)
        q.put((rank, mine, bool(ok), float(t)))
    finally:
        dist.destroy_process_group()


def test_frames_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]       # frame i -> rank i mod N, nothing dropped
    assert all(r[2] for r in res) and res[0][3] == 5.0


def test_single_process_gather_is_identity():
    b = torch.arange(32, dtype=torch.float32).reshape(2, 16)
    assert parallel.gather_frames(b)[0] is b
    bufs, work = parallel.gather_frames_async(b)
    assert bufs[0] is b and work is None
    un = parallel.unpack_frame(b)
    assert torch.equal(parallel.pack_frame(un), b)


def test_frame_of_step_covers_every_frame_once_per_pass_and_rotates():
    for world, n_frames in ((1, 16), (2, 16), (8, 16), (4, 32)):
        steps = n_frames // world
        seen = sorted(parallel.frame_of_step(s, r, world, n_frames) for s in range(steps) for r in range(world))
        assert seen == list(range(n_frames))                      # one pass = every frame exactly once
        # inside a step the ranks render distinct frames
        for s in range(steps):
            assert len({parallel.frame_of_step(s, r, world, n_frames) for r in range(world)}) == world
        if world > 1:
            # over world passes' worth of steps a rank sees every residue class
            res = {parallel.frame_of_step(s, 0, world, n_frames) % world for s in range(world)}
            assert res == set(range(world))


def test_assign_frames_longest_first_balances_the_ranks():
    costs = {f: c for f, c in enumerate([1022, 1059, 1091, 1098, 1053, 954, 822, 694, 658, 642, 633, 608, 553, 588, 723, 737])}
    frames = [f % 16 for f in range(40)]                       # 5 steps x 8 ranks
    per_rank = parallel.assign_frames(frames, costs, 8)
    assert sorted(f for r in per_rank for f in r) == sorted(frames)          # every frame exactly as often as asked
    assert all(len(r) == 5 for r in per_rank)
    sums = [sum(costs[f] for f in r) for r in per_rank]
    assert max(sums) / (sum(sums) / 8) < 1.02                   # (round robin f -> f mod N: 1.07 on this sequence)
    rr = [sum(costs[f] for f in frames[r::8]) for r in range(8)]
    assert max(rr) / (sum(rr) / 8) > 1.05
    assert parallel.assign_frames([3, 1, 2], costs, 1) == [[3, 2, 1]]
