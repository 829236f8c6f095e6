#!/usr/bin/env python
"""bench.py -- headline benchmark of the render hot path (BASELINE.json metric).

A "step" = one frame of the reference's test loop: ``model.prepare(batch)`` (pose -> voxel precompute,
test occupancy grid, envmap pdf/CDF + light sample tables) followed by ``model.forward(rays)`` for all
H*W primary rays at ``spp`` shading samples per pixel, render_mode="light".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--res 512] [--spp 1024] [--gi 0|1]
  python bench.py --impl reference ...     # the CPU restatement (oracle port) on the host cores

Default workload = BASELINE.json configs[3], the single-GPU configuration its metric is quoted on: 512x512
relight, 1024 spp, render_mode=light, global_illumination=true (one indirect bounce).  ``--gi 0`` is the
reference README's relight command (README.md:84-95: same size, global_illumination=false).

Metric: shaded samples/s = (primary rays x spp) / time, whole job over all ranks ("weak" scaling: one
frame per rank per step -- every rank renders frame (step mod 8) of the sequence so that the per-GPU work is the
same for every N; --distinct-frames shards the sequence frame f -> rank f mod N instead --, no data-path collective; the finished frame
buffers are gathered to rank 0 with one asynchronous NCCL gather per step, all of them completed
inside the timed region).
  value : inputs resident in HBM when the timed region starts (rays, envmap on device)
  e2e   : through IntrinsicAvatarModel.prepare/forward with HOST rays + HOST hdri, H2D and the D2H of
          every output buffer inside the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENE_AABB = [-1.25, -1.55, -1.25, 1.25, 0.95, 1.25]
METRIC = "shaded_samples_per_sec"
UNIT = "samples/s"

# Algorithmic bytes per unit of work in the formats the kernels actually read (DESIGN.md "Roofline"):
#   voxel_J trilinear fetch   8 corners x 12 x fp32 = 384 B   (per Broyden fetch)
#   skinning-weight fetch     8 corners x 24 x fp32 = 768 B   (per with-gradient query that is valid)
#   hash grid (geo / rad)     16 levels x 8 corners x 2 x fp32 = 1024 B   (per canonical evaluation)
B_BROYDEN_FETCH, B_SKIN_FETCH, B_HASH_EVAL = 384, 768, 1024


def algorithmic_bytes(cnt: dict, n_rays: int, spp: int) -> int:
    return (B_BROYDEN_FETCH * cnt["broyden_fetch"] + B_SKIN_FETCH * cnt["skin_fetch"]
            + B_HASH_EVAL * (cnt["geo_eval"] + cnt["rad_eval"]) + 52 * spp + 96 * n_rays)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in o.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def scene_inputs(res: int, spp: int, frame: int):
    from intrinsicavatar_b200 import synthetic as syn
    bp, go, tr = syn.load_pose(frame)
    rays = syn.make_rays(res, res, tr)
    tabs = syn.random_tables(spp, 64, seed=0)
    return bp, go, tr, rays, tabs


def run_reference(args):
    """--impl reference: the reference has no CPU path and cannot run here, so this arm
    times the oracle port of the same path on the host cores, each step a bounded sample of the
    workload (a sub-frame at reduced resolution / spp, same camera, pose, weights, light)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    v, sample, ms, _ = cpu_baseline(args, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args, steps=1, warmup=0):
    """Oracle port on the host cores over a bounded sample: cpu_res^2 primary rays at cpu_spp."""
    import torch
    from intrinsicavatar_b200 import synthetic as syn
    from intrinsicavatar_b200.snarf import SnarfSetup
    from intrinsicavatar_b200.weights import fold, hashgrid_layout, random_state_dict
    from oracle.fields import Fields
    from oracle.render import OracleRenderer

    res, spp = args.cpu_res, args.cpu_spp
    snarf = SnarfSetup()
    folded, layout = fold(random_state_dict(0)), hashgrid_layout()
    bp, go, tr = syn.load_pose(0)
    fr = snarf.frame(bp, go, tr)
    tabs = syn.random_tables(spp, args.cpu_grid, seed=0)
    env = syn.load_envmap_full()
    rays = torch.from_numpy(syn.make_rays(res, res, tr))
    R = OracleRenderer(Fields(folded, layout, snarf.bbox), snarf.lbs_voxel, snarf.offset_kernel, snarf.scale_kernel,
                       samples_per_pixel=spp, global_illumination=bool(args.gi), grid_res=args.cpu_grid,
                       render_mode=args.render_mode)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        R.set_pose(fr["tfs"], fr["w2s"])
        R.build_occupancy(fr["deformed_bbox"], tabs["jitter"])
        if args.render_mode == "uniform_light":
            R.set_light_uniform(env, 2, spp // 2)
        else:
            R.set_light(env, tabs["u1"], tabs["u2"])
        R.forward(rays, seed=0)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    sample = (f"{res}x{res} rays x {spp} spp frame (prepare incl. {args.cpu_grid}^3 occupancy grid + forward), same "
              f"camera/pose/weights/light, gi={int(bool(args.gi))}; {len(times)} timed run(s)")
    return res * res * spp / t, sample, t * 1e3, os.cpu_count() or 1


def workload_config(args):
    return {
        "workload": f"{args.res}x{args.res} relight frame, {args.spp} spp, render_mode={args.render_mode}, "
                    f"global_illumination={'true' if args.gi else 'false'}, prepare+forward per step",
        "frame_source": "AIST pose frames 0..7 (frame = " + ("(step + rank)" if args.distinct_frames else "step") + " mod 8 on every rank), synthetic 24-joint body, random-init "
                        "hash grids + MLPs (seed 0), city.hdr envmap at 1024x2048 (the reference's file as AnimationDataset loads it)",
        "rays_per_frame": args.res * args.res, "spp": args.spp, "gi": bool(args.gi),
        "parallelism": f"frame-per-gpu x{args.gpus}",
        "l2": "flushed between steps (256 MiB write) and per-step sample streams (3.2 GB at 512^2 x 1024) exceed L2",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--gi", type=int, default=1, help="config.model.global_illumination (default 1 = BASELINE configs[3])")
    ap.add_argument("--render-mode", default="light", choices=["light", "uniform_light", "mats", "mis"],
                    help="config.model.render_mode (uniform_light needs --spp 512); the headline workload is light")
    ap.add_argument("--cpu-res", type=int, default=64)
    ap.add_argument("--cpu-spp", type=int, default=8)
    ap.add_argument("--cpu-grid", type=int, default=32)
    ap.add_argument("--distinct-frames", action="store_true",
                    help="N>1: rank r renders frame (step + r) mod 8 instead of every rank rendering frame step mod 8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback "
                         "(use --impl reference for the CPU oracle port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from intrinsicavatar_b200 import parallel, synthetic as syn
    from intrinsicavatar_b200.model import IntrinsicAvatarModel

    cfg = {"samples_per_pixel": args.spp, "global_illumination": bool(args.gi), "scene_aabb": SCENE_AABB,
           "render_mode": args.render_mode}
    model = IntrinsicAvatarModel(cfg, device=local_rank, seed=0)
    model.train(False)
    model.update_step(250, 25000)
    eng = model.engine
    eng.set_timing(True)

    n_rays = args.res * args.res
    env_h = torch.from_numpy(syn.load_envmap_full()).pin_memory()
    env_d = env_h.to(dev)
    tabs = syn.random_tables(args.spp, 64, seed=0)
    jitter = torch.from_numpy(tabs["jitter"]).to(dev)
    lu = (torch.from_numpy(tabs["u1"]).to(dev), torch.from_numpy(tabs["u2"]).to(dev))
    frames = []
    for f in range(8):
        bp, go, tr = syn.load_pose(f)
        rays_h = torch.from_numpy(syn.make_rays(args.res, args.res, tr)).pin_memory()
        frames.append({"batch": {"body_pose": bp[None], "global_orient": go[None], "transl": tr[None]},
                       "rays_h": rays_h, "rays_d": rays_h.to(dev)})
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def frame_of(step):
        # Weak scaling = the same work on every GPU for every N: each rank renders its own copy of the same
        # animation sequence.  (--distinct-frames: frame = (step + rank) mod 8, the deployment sharding of BASELINE
        # configs[4]; the frames of this sequence cost 0.85-1.2 s each, so that variant also measures their spread.)
        return frames[(step + (rank if args.distinct_frames else 0)) % 8]

    pending = []   # (send block, receive blocks, work) of the frame gathers posted and not yet waited for

    def post_gather(block):
        # the gather is posted asynchronously: frames take different times, and rank 0 renders its next frame instead
        # of waiting for the slowest rank of this step; all gathers are waited for inside the timed region (drain)
        bufs, work = parallel.gather_frames_async(block, dst=0)
        pending.append((block, bufs, work))

    def drain():
        for _, _, work in pending:
            if work is not None:
                work.wait()
        pending.clear()

    def step_device(step):
        fr = frame_of(step)
        model.prepare({**fr["batch"], "hdri": env_d}, jitter=jitter, light_uniforms=lu)
        out = model.forward(fr["rays_d"], move_to_cpu=False)
        if world > 1:
            post_gather(parallel.pack_frame(out))
        return out

    def step_e2e(step):
        fr = frame_of(step)
        model.prepare({**fr["batch"], "hdri": env_h.to(dev, non_blocking=True)}, jitter=jitter, light_uniforms=lu)
        out = model.forward(fr["rays_h"], move_to_cpu=True)
        if world > 1:
            post_gather(parallel.pack_frame({k: out[k].to(dev) for k in parallel.FRAME_KEYS}))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for s in range(warmup):
            fn(s)
            flush.fill_(s & 0xFF)
        drain()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_ms, cnts = [], []
        l0 = eng.timings()[1]
        t_wall = time.perf_counter()
        ev0.record()
        for s in range(steps):
            fn(warmup + s)
            flush.fill_(s & 0xFF)              # L2 flush between timed iterations (inside the timed region)
            if fn is step_device:
                stage_ms.append(eng.timings()[0])  # syncs the stream: the stages of this step
                cnts.append(eng.counters())
        drain()                                # every frame of the timed steps has arrived on rank 0
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        wall = (time.perf_counter() - t_wall) * 1e3
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launches = eng.timings()[1] - l0
        return float(t.item()), wall, stage_ms, cnts, launches

    with ClockSampler(local_rank) as clk:
        ms_total, wall_ms, stage_ms, cnts, launches = timed(step_device, args.steps, args.warmup)
    clocks = clk.summary()
    samples_per_step = n_rays * args.spp * world
    value = samples_per_step * args.steps / (ms_total * 1e-3)

    e2e = None
    if not args.no_e2e:
        # same warm-up count as the device arm, so that both arms time the same frames (step mod 8)
        ms_e2e, _, _, _, _ = timed(step_e2e, args.steps, args.warmup)
        # forward() brings every output buffer of the frame to the host in one packed copy (engine.outputs_to_host)
        d2h = eng.alloc_outputs(1)["_block"].numel() * 4 * n_rays
        h2d = frames[0]["rays_h"].numel() * 4 + env_h.numel() * 4 + (24 * 16 + 16) * 4
        e2e = {"value": samples_per_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
               "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    if rank == 0:
        peak, peak_src = peaks()
        # dominant kernel: k_shade (secondary-ray integrator).  Launch duration = CUDA events recorded by the
        # library on the launching stream around that launch (ia_get_timings), averaged over the timed steps.
        def avg(key):
            v = [s[key] for s in stage_ms if s[key] >= 0]
            return sum(v) / len(v) if v else 0.0
        shade_ms = avg("shade")
        keys = [k for k in cnts[0] if k != "primary"] if cnts else []
        c = {k: sum(cc[k] for cc in cnts) / max(1, len(cnts)) for k in keys}
        cp = {k: sum(cc["primary"][k] for cc in cnts) / max(1, len(cnts)) for k in keys if k != "hit_rays"} if cnts else {}
        # work of the shading kernel alone = totals - snapshot taken when the primary stage had finished
        cs = {k: c[k] - cp.get(k, 0) for k in keys if k != "hit_rays"}
        n_samples = c.get("hit_rays", 0) * args.spp
        # + the sample streams the kernel reads (rs_src, rs_w 4 B each per shading sample; rs_t 4 B, the 48-B
        #   IaSample and 6 fp32 accumulations per traced ray)
        #   with global illumination also the radiance hash grid and the 24-channel skinning-weight fetch of every
        #   fine sample's root (its with-gradient geometry evaluation is counted in geo_eval)
        alg = (B_BROYDEN_FETCH * cs["broyden_fetch"] + B_HASH_EVAL * (cs["geo_eval"] + cs["rad_eval"])
               + B_SKIN_FETCH * cs["skin_fetch"] + 8 * n_samples + (4 + 48 + 24) * cs["secondary_rays"]) if cs else 0
        achieved = alg / (shade_ms * 1e-3) / 1e9 if shade_ms > 0 else 0.0
        render_ms = sum(avg(k) for k in ("setup", "primary", "resample", "shade", "composite"))
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": peak_src, "kernel": "k_shade_wf<%d,%s> (wavefront secondary-ray integrator)" % (int(bool(args.gi)), args.render_mode),
                    "algorithmic_bytes_per_launch": alg, "launch_ms": shade_ms, "share_of_step": shade_ms / (ms_total / args.steps),
                    "units_per_launch": {"broyden_voxel_fetches": cs.get("broyden_fetch"), "geometry_evals": cs.get("geo_eval"),
                                         "radiance_evals": cs.get("rad_eval"), "skinning_fetches": cs.get("skin_fetch"),
                                         "secondary_rays": cs.get("secondary_rays"), "shading_samples": n_samples},
                    "bytes_per_unit": {"broyden_voxel_fetch": B_BROYDEN_FETCH, "geometry_eval": B_HASH_EVAL,
                                       "radiance_eval": B_HASH_EVAL, "skinning_fetch": B_SKIN_FETCH,
                                       "shading_sample": 8, "secondary_ray": 76},
                    "note": "the gathers' working set (voxel_J 25 MB, geometry hash grid 50 MB) is L2-resident by design, "
                            "so the bytes the kernel requests are served by L2/L1, not HBM: DRAM traffic (`traffic`) is far "
                            "below the algorithmic bytes and frac can exceed what HBM could deliver (DESIGN.md, Roofline)"}
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            with open(tfile) as f:
                roofline["traffic"] = json.load(f).get("k_shade_wf_gi%d_dram_bytes_per_launch" % int(bool(args.gi)))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "stages_ms": {k: avg(k) for k in (stage_ms[0] if stage_ms else {})},
            "counters_per_frame": c, "counters_primary_stage": cp, "wall_ms_total": wall_ms,
            "ms_per_frame": ms_total / args.steps,
        }
        if not args.no_cpu_baseline and world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            v, sample, _, cores = cpu_baseline(args)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
