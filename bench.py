#!/usr/bin/env python
"""bench.py -- headline benchmark of the render hot path (BASELINE.json metric).

A "step" = one frame of the reference's test loop: ``model.prepare(batch)`` (pose -> voxel precompute,
test occupancy grid, envmap pdf/CDF + light sample tables) followed by ``model.forward(rays)`` for all
H*W primary rays at ``spp`` shading samples per pixel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3] [--res 512] [--spp 1024] [--gi 0|1]
  python bench.py --impl reference ...     # the CPU restatement (oracle port) on the host cores

Workloads (BASELINE.json ``configs``; ``--config``):
  3 (default)  configs[3]: 512x512 relight, 1024 spp, render_mode=light, global_illumination=true -- the single-GPU
               configuration the metric is quoted on
  2            configs[2]: the same at 256 spp, global_illumination=false
  1            configs[1]: 512x512 primary-only volume rendering (no secondary rays; metric counts 1 sample per ray)
The default run also times configs[1] and configs[2] for three steps each and reports them under ``other_configs``.

Metric: shaded samples/s = (primary rays x spp) / time, whole job over all ranks.  "weak" scaling: one frame per rank
per step, no data-path collective.  N > 1 is BASELINE configs[4]: the frames of the animation sequence are sharded over
the ranks -- step s covers N consecutive entries of a 16-frame AIST sequence -- and every finished frame is delivered to
rank 0 inside the timed region
(parallel.FrameCollector: stream-ordered peer copies over NVLink, or one asynchronous NCCL gather per step).
Frame costs differ by 2x (FRAME_COST_MS): the sequence is walked in an order that alternates expensive and cheap frames and
the frames of a run are dealt to the ranks longest-first (parallel.assign_frames).
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
N_SEQ_FRAMES = 16      # frames of the AIST sequence one pass of the multi-GPU job covers
# Measured cost of every frame of the sequence (ms per frame, configs[3], 1 x B200, profiles/r2_bench_frames16.json): the
# body turns and its limbs occlude each other differently from frame to frame -- 480 .. 947 ms.  Used (a) to put the
# sequence in an order whose every window is representative (expensive and cheap frames alternate), so that a short run
# does not time an unrepresentative stretch, and (b) as the estimate for the longest-first assignment of a step batch's
# frames to the ranks (parallel.assign_frames).
FRAME_COST_MS = [891.9, 922.6, 944.7, 947.1, 898.8, 816.4, 703.1, 594.3, 565.6, 555.7, 549.7, 527.8, 479.7, 508.9, 625.5, 635.9]   # profiles/r2_bench_frames16.json


def sequence_order():
    by_cost = sorted(range(N_SEQ_FRAMES), key=lambda f: -FRAME_COST_MS[f])
    order = []
    for i in range(N_SEQ_FRAMES // 2):
        order += [by_cost[i], by_cost[N_SEQ_FRAMES - 1 - i]]
    return order

# Algorithmic bytes per unit of work.  SURVEY.md 8(d) fixes them in the B200 DESIGN formats (fp16 channels-last):
#   voxel_J trilinear fetch 8 corners x 12 x 2 B = 192 B   per Broyden fetch
#   skinning-weight fetch   8 corners x 24 x 2 B = 384 B   per with-gradient query that is valid
#   hash grid (geo / rad)   16 levels x 8 corners x 2 x 2 B = 512 B   per canonical evaluation
# `roofline.achieved` / `frac` use these.  The product stores fp32 (parity: DESIGN.md section 3), i.e. it requests twice
# these bytes from the memory system; that figure is reported next to it as `achieved_as_stored`.
B_SURVEY = {"broyden_fetch": 192, "skin_fetch": 384, "hash_eval": 512}
B_STORED = {"broyden_fetch": 384, "skin_fetch": 768, "hash_eval": 1024}
CONFIGS = {1: dict(spp=1024, gi=0, primary_only=True), 2: dict(spp=256, gi=0, primary_only=False),
           3: dict(spp=1024, gi=1, primary_only=False)}


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

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self.stop, self.th, self.enabled = index, [], threading.Event(), None, enabled

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
        if self.enabled:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.th is not None:
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
# CPU arm: the reference has no CPU path and cannot run on the box, so the oracle port is what is timed
# on the host cores.  Recipe of SURVEY.md 8(d): 64 x 64 frames at 4 / 16 / 64 spp, a per-sample cost fitted to them, and
# the cost of the benched frame extrapolated from the fit.
CPU_RES, CPU_SPPS, CPU_GRID = 64, (4, 16, 64), 32


class CpuOracle:
    def __init__(self, args):
        import torch
        from intrinsicavatar_b200 import synthetic as syn
        from intrinsicavatar_b200.snarf import SnarfSetup
        from intrinsicavatar_b200.weights import fold, hashgrid_layout, random_state_dict
        from oracle.fields import Fields
        self.torch, self.syn, self.args = torch, syn, args
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.snarf = SnarfSetup()
        self.fields = Fields(fold(random_state_dict(0)), hashgrid_layout(), self.snarf.bbox)
        bp, go, tr = syn.load_pose(0)
        self.fr = self.snarf.frame(bp, go, tr)
        self.env = syn.load_envmap_full()
        self.rays = torch.from_numpy(syn.make_rays(CPU_RES, CPU_RES, tr))

    def frame(self, spp):
        """(seconds of prepare, seconds of forward) of one 64 x 64 frame at ``spp``."""
        from oracle.render import OracleRenderer
        a = self.args
        tabs = self.syn.random_tables(max(spp, 2), CPU_GRID, seed=0)
        R = OracleRenderer(self.fields, self.snarf.lbs_voxel, self.snarf.offset_kernel, self.snarf.scale_kernel,
                           samples_per_pixel=max(spp, 2), global_illumination=bool(a.gi), grid_res=CPU_GRID,
                           render_mode=a.render_mode)
        t0 = time.perf_counter()
        R.set_pose(self.fr["tfs"], self.fr["w2s"])
        R.build_occupancy(self.fr["deformed_bbox"], tabs["jitter"])
        if a.render_mode == "uniform_light":
            R.set_light_uniform(self.env, 2, spp // 2)
        else:
            R.set_light(self.env, tabs["u1"], tabs["u2"])
        t1 = time.perf_counter()
        R.forward(self.rays, seed=0, albedo_only=bool(a.primary_only))
        return t1 - t0, time.perf_counter() - t1

    @staticmethod
    def fit(points):
        """least squares t_forward = a + b * spp over [(spp, seconds)] -> (a = per-frame cost of the primary stage of the
        64 x 64 rays, b = seconds per spp of those rays)"""
        n = len(points)
        sx = sum(p[0] for p in points); sy = sum(p[1] for p in points)
        sxx = sum(p[0] * p[0] for p in points); sxy = sum(p[0] * p[1] for p in points)
        den = n * sxx - sx * sx
        if den == 0:
            return points[0][1], 0.0
        b = (n * sxy - sx * sy) / den
        return (sy - b * sx) / n, b

    def extrapolate(self, t_prep, a, b):
        """seconds per frame of the benched workload: the grid build does not scale with the image (the bench builds
        a 64^3 grid: 8x the cells of the 32^3 one timed here), the primary stage scales with the rays, the shading stage
        with rays x spp."""
        args = self.args
        scale = (args.res * args.res) / float(CPU_RES * CPU_RES)
        spp = 0 if args.primary_only else args.spp
        return t_prep * (64 ** 3) / float(CPU_GRID ** 3) + a * scale + b * spp * scale


def cpu_baseline(args, timed_spp16_runs=1, oracle=None):
    """-> dict for the JSON line.  One timed 64 x 64 frame at each of 4 / 16 / 64 spp (16 spp: ``timed_spp16_runs``
    runs), fit, extrapolation to the benched frame."""
    O = oracle or CpuOracle(args)
    O.frame(2)                                           # warm-up (page-in, thread pools)
    pts, preps = [], []
    spps = (1,) if args.primary_only else CPU_SPPS
    for spp in spps:
        for _ in range(timed_spp16_runs if spp == 16 else 1):
            tp, tf = O.frame(spp)
            preps.append(tp)
            pts.append((spp, tf))
    a, b = (pts[0][1], 0.0) if args.primary_only else O.fit(pts)
    t_prep = sum(preps) / len(preps)
    frame_s = O.extrapolate(t_prep, a, b)
    n_samples = args.res * args.res * (1 if args.primary_only else args.spp)
    return {
        "value": n_samples / frame_s, "unit": UNIT, "cores": O.cores, "kind": "port",
        "sample": (f"oracle port (pure PyTorch fp32, {O.cores} threads), {CPU_RES}x{CPU_RES} rays of the same camera / pose / "
                   f"weights / city.hdr at spp {list(spps)} with a {CPU_GRID}^3 occupancy grid, one timed frame each after a "
                   f"warm-up frame; forward seconds {[round(t, 2) for _, t in pts]}; fit t = a + b*spp: a={a:.3f} s, b={b:.4f} s; "
                   f"prepare {t_prep:.2f} s; extrapolated to the benched frame"),
        "measured_points": [{"spp": s, "forward_s": t} for s, t in pts], "prepare_s": t_prep,
        "fit": {"a_s": a, "b_s_per_spp": b}, "extrapolated_s_per_frame": frame_s,
        "per_sample_rate": (CPU_RES * CPU_RES / b) if b > 0 else None,
    }


def run_reference(args):
    """--impl reference: rank 0 alone; a step = one 64 x 64 x 16 spp frame of the oracle port (the bounded sample), the
    4 and 64 spp frames of the fit are taken once before the timed steps.  ``value`` is the benched workload's
    throughput extrapolated from the fit, so that it is comparable with the GPU arm's line."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    O = CpuOracle(args)
    for _ in range(min(max(args.warmup, 1), 2)):
        O.frame(2)
    spps = (1,) if args.primary_only else CPU_SPPS
    pts, preps = [], []
    for spp in spps:
        if spp == 16:
            continue
        tp, tf = O.frame(spp)
        preps.append(tp); pts.append((spp, tf))
    step_s = []
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        tp, tf = O.frame(1 if args.primary_only else 16)
        step_s.append(time.perf_counter() - t0)
        preps.append(tp)
        if not args.primary_only:
            pts.append((16, tf))
        elif not pts:
            pts.append((1, tf))
    a, b = (sum(t for _, t in pts) / len(pts), 0.0) if args.primary_only else O.fit(pts)
    t_prep = sum(preps) / len(preps)
    frame_s = O.extrapolate(t_prep, a, b)
    n_samples = args.res * args.res * (1 if args.primary_only else args.spp)
    v = n_samples / frame_s
    cfg = workload_config(args, 1)
    cfg["cpu_sample"] = (f"each timed step = one {CPU_RES}x{CPU_RES} frame at {1 if args.primary_only else 16} spp "
                         f"(prepare with a {CPU_GRID}^3 grid + forward) of the oracle port on {O.cores} host threads; value = the "
                         f"workload above extrapolated from t = a + b*spp fitted to spp {list(spps)} (a={a:.3f} s, b={b:.4f} s per "
                         f"spp for {CPU_RES * CPU_RES} rays, prepare {t_prep:.2f} s): {frame_s:.0f} s per frame")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(step_s) / len(step_s), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": O.cores, "kind": "port", "sample": cfg["cpu_sample"],
                         "extrapolated_s_per_frame": frame_s, "fit": {"a_s": a, "b_s_per_spp": b}, "prepare_s": t_prep,
                         "timed_step_s": step_s},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    what = ("primary-only volume rendering (no secondary rays)" if args.primary_only else
            f"relight frame, {args.spp} spp, render_mode={args.render_mode}, "
            f"global_illumination={'true' if args.gi else 'false'}")
    return {
        "workload": f"BASELINE configs[{args.config}]: {args.res}x{args.res} {what}, prepare+forward per step",
        "frame_source": (f"AIST animation sequence, frames 0..{N_SEQ_FRAMES - 1} in an order that alternates expensive and cheap "
                         f"frames ({sequence_order()}; per-frame cost 480..947 ms): step s covers entries [s*N, s*N+N) of it, the "
                         "frames of the timed steps are dealt to the N ranks longest-first (parallel.assign_frames); synthetic "
                         "24-joint body, random-init hash grids + MLPs (seed 0), the reference's city.hdr at 1024x2048"),
        "rays_per_frame": args.res * args.res, "spp": 1 if args.primary_only else args.spp, "gi": bool(args.gi),
        "parallelism": f"frame-per-gpu x{world}",
        "l2": "flushed between steps (256 MiB write) and per-step sample streams (3.2 GB at 512^2 x 1024) exceed L2",
    }


def roofline_of(args, stage_ms, cnts, ms_per_step):
    """Roofline object of the dominant kernel from device counters and the library's CUDA-event stage times."""
    peak, peak_src = peaks()

    def avg(key):
        v = [s[key] for s in stage_ms if s[key] >= 0]
        return sum(v) / len(v) if v else 0.0
    keys = [k for k in cnts[0] if k != "primary"] if cnts else []
    c = {k: sum(cc[k] for cc in cnts) / max(1, len(cnts)) for k in keys}
    cp = {k: sum(cc["primary"][k] for cc in cnts) / max(1, len(cnts)) for k in keys if k != "hit_rays"} if cnts else {}
    if args.primary_only:
        kernel, ms, u = "k_prim_edges + k_prim_shade + k_prim_accum (primary stage)", avg("primary"), cp
    else:
        # work of the shading kernel alone = totals - snapshot taken when the primary stage had finished
        kernel = "k_shade_wf<%d,%s> (wavefront secondary-ray integrator)" % (int(bool(args.gi)), args.render_mode)
        ms, u = avg("shade"), {k: c[k] - cp.get(k, 0) for k in keys if k != "hit_rays"}
    n_samples = c.get("hit_rays", 0) * args.spp
    n_rays = args.res * args.res

    def alg(B):
        if not u:
            return 0
        gathers = (B["broyden_fetch"] * u["broyden_fetch"] + B["hash_eval"] * (u["geo_eval"] + u["rad_eval"])
                   + B["skin_fetch"] * u["skin_fetch"])
        if args.primary_only:
            return gathers + 96 * n_rays
        # + the sample streams the kernel reads (rs_src, rs_w per shading sample; rs_t, the 48-B IaSample and 6 fp32
        #   accumulations per traced ray)
        return gathers + 8 * n_samples + (4 + 48 + 24) * u["secondary_rays"]
    a_survey, a_stored = alg(B_SURVEY), alg(B_STORED)
    achieved = a_survey / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    # bytes that MUST cross the HBM pins per launch: the per-sample streams once (12 B per shading sample) + one pass over
    # the tables the gathers hit (voxel_J 25 MB, geometry hash 50 MB, with GI radiance hash 50 MB + skinning weights 50 MB)
    compulsory = 12 * n_samples + (75 << 20) + ((100 << 20) if args.gi else 0)
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
         "peak_source": peak_src, "kernel": kernel, "launch_ms": ms, "share_of_step": ms / ms_per_step if ms_per_step else None,
         "algorithmic_bytes_per_launch": a_survey,
         "bytes_per_unit": {"broyden_voxel_fetch": B_SURVEY["broyden_fetch"], "geometry_eval": B_SURVEY["hash_eval"],
                            "radiance_eval": B_SURVEY["hash_eval"], "skinning_fetch": B_SURVEY["skin_fetch"],
                            "shading_sample": 8, "secondary_ray": 76, "source": "SURVEY.md 8(d) (fp16 design formats)"},
         "achieved_as_stored": a_stored / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
         "frac_as_stored": (a_stored / (ms * 1e-3) / 1e9 / peak) if ms > 0 else 0.0,
         "bytes_per_unit_as_stored": {"broyden_voxel_fetch": B_STORED["broyden_fetch"], "geometry_eval": B_STORED["hash_eval"],
                                      "skinning_fetch": B_STORED["skin_fetch"], "note": "fp32 storage, what the loads request"},
         "units_per_launch": {"broyden_voxel_fetches": u.get("broyden_fetch"), "geometry_evals": u.get("geo_eval"),
                              "radiance_evals": u.get("rad_eval"), "skinning_fetches": u.get("skin_fetch"),
                              "secondary_rays": u.get("secondary_rays"), "shading_samples": n_samples},
         "compulsory_hbm_bytes": compulsory, "l1_sector_bytes": None,
         "note": "the gathers' working set (voxel_J 25 MB, geometry hash grid 50 MB) is L2-resident by design: the bytes the "
                 "kernel requests are served by L1/L2, so DRAM traffic (`traffic`, ncu) is far below the algorithmic bytes; "
                 "the kernel is bound by L1 gather throughput (DESIGN.md section 5), frac is algorithmic bytes over the HBM "
                 "copy peak"}
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tfile) and not args.primary_only:
        with open(tfile) as f:
            t = json.load(f)
        r["traffic"] = t.get("k_shade_wf_gi%d_dram_bytes_per_launch" % int(bool(args.gi)))
        r["l1_sector_bytes"] = t.get("k_shade_wf_gi%d_l1_sector_bytes_per_launch" % int(bool(args.gi)))
        r["traffic_source"] = t.get("source")
    return r, c, cp, {k: avg(k) for k in (stage_ms[0] if stage_ms else {})}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3], help="BASELINE.json configs[i] (see module doc)")
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--spp", type=int, default=None)
    ap.add_argument("--gi", type=int, default=None, help="config.model.global_illumination")
    ap.add_argument("--render-mode", default="light", choices=["light", "uniform_light", "mats", "mis"],
                    help="config.model.render_mode (uniform_light needs --spp 512); the headline workload is light")
    ap.add_argument("--same-frame", action="store_true",
                    help="N>1: every rank renders frame (step mod 16) -- replicas instead of the sharded sequence")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="frame delivery to rank 0 (parallel.FrameCollector)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    preset = CONFIGS[args.config]
    args.spp = preset["spp"] if args.spp is None else args.spp
    args.gi = preset["gi"] if args.gi is None else args.gi
    args.primary_only = preset["primary_only"]

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and args.transport == "nccl":
        os.environ.setdefault("NCCL_MAX_CTAS", "1")    # see parallel.py: a resident send kernel holds SMs until its receive starts
    import torch
    import torch.distributed as dist

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

    n_rays = args.res * args.res
    env_h = torch.from_numpy(syn.load_envmap_full()).pin_memory()
    env_d = env_h.to(dev)
    frames = []
    for f in range(N_SEQ_FRAMES):
        bp, go, tr = syn.load_pose(f)
        rays_h = torch.from_numpy(syn.make_rays(args.res, args.res, tr)).pin_memory()
        frames.append({"batch": {"body_pose": bp[None], "global_orient": go[None], "transl": tr[None]},
                       "rays_h": rays_h, "rays_d": rays_h.to(dev)})
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    collector = parallel.FrameCollector(n_rays, 2, dev, transport=args.transport) if world > 1 else None

    # Frame schedule.  The job renders the sequence in `sequence_order()`: step s covers its entries [s N, s N + N).  The
    # frames of the timed steps (and, separately, of the warm-up steps) are dealt to the ranks longest-first, the same
    # number to every rank; each rank renders its share most expensive first.  N = 1: the sequence in order.
    order = sequence_order()
    costs = {f: FRAME_COST_MS[f] for f in range(N_SEQ_FRAMES)}

    def share(first_step, n_steps):
        seq = [order[i % N_SEQ_FRAMES] for i in range(first_step * world, (first_step + n_steps) * world)]
        if world == 1 or args.same_frame:
            return seq[::world] if args.same_frame else seq
        return parallel.assign_frames(seq, costs, world)[rank]
    schedule = {}

    def frame_of(step):
        return schedule[step]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_model(a):
        cfg = {"samples_per_pixel": a.spp, "global_illumination": bool(a.gi), "scene_aabb": SCENE_AABB,
               "render_mode": a.render_mode}
        m = IntrinsicAvatarModel(cfg, device=local_rank, seed=0)
        m.train(False)
        m.update_step(250, 25000)
        m.albedo_only = bool(a.primary_only)
        m.engine.set_timing(True)
        tabs = syn.random_tables(a.spp, 64, seed=0)
        jitter = torch.from_numpy(tabs["jitter"]).to(dev)
        lu = (torch.from_numpy(tabs["u1"]).to(dev), torch.from_numpy(tabs["u2"]).to(dev))
        return m, jitter, lu

    def run_arm(a, e2e_arm, steps, warmup, per_step_stats):
        """-> (ms total (max over ranks), per-step stage times, per-step counters, launches)"""
        model, jitter, lu = make_model(a)
        eng = model.engine

        def step_fn(step):
            fr = frames[frame_of(step)]
            hdri = env_h.to(dev, non_blocking=True) if e2e_arm else env_d
            batch = {**fr["batch"]} if a.primary_only else {**fr["batch"], "hdri": hdri}
            model.prepare(batch, jitter=jitter, light_uniforms=lu)
            out = model.forward(fr["rays_h"] if e2e_arm else fr["rays_d"], move_to_cpu=e2e_arm)
            if collector is not None:
                collector.collect(step, parallel.pack_frame({k: out[k].to(dev, non_blocking=True) for k in parallel.FRAME_KEYS}
                                                            if e2e_arm else out))
        schedule.clear()
        schedule.update({s: f for s, f in enumerate(share(0, warmup))})
        schedule.update({warmup + s: f for s, f in enumerate(share(warmup, steps))})
        for s in range(warmup):
            step_fn(s)
            flush.fill_(s & 0xFF)
        if collector is not None:
            collector.finish()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_ms, cnts = [], []
        l0 = eng.timings()[1]
        ev0.record()
        for s in range(steps):
            step_fn(warmup + s)
            flush.fill_(s & 0xFF)              # L2 flush between timed iterations (inside the timed region)
            if per_step_stats:                 # (N = 1 only: reading the stage events syncs the stream)
                stage_ms.append(eng.timings()[0])
                cnts.append(eng.counters())
        if collector is not None:
            collector.finish()                 # every frame of the timed steps has arrived on rank 0
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if not per_step_stats:                 # N > 1: the last step's figures, read after the timed region
            stage_ms.append(eng.timings()[0])
            cnts.append(eng.counters())
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launches = eng.timings()[1] - l0
        del model
        return float(t.item()), stage_ms, cnts, launches

    spp_metric = 1 if args.primary_only else args.spp
    samples_per_step = n_rays * spp_metric * world
    with ClockSampler(local_rank, enabled=(rank == 0)) as clk:
        ms_total, stage_ms, cnts, launches = run_arm(args, False, args.steps, args.warmup, world == 1)
    clocks = clk.summary()
    main_frames = [schedule[args.warmup + s] for s in range(args.steps)]     # (the later arms re-plan the schedule)
    value = samples_per_step * args.steps / (ms_total * 1e-3)

    e2e = None
    if not args.no_e2e:
        # same warm-up count as the device arm, so that both arms time the same frames
        ms_e2e, _, _, _ = run_arm(args, True, args.steps, args.warmup, False)
        from intrinsicavatar_b200.engine import OUTPUT_SPECS
        # forward() brings every output buffer of the frame to the host in one packed copy (engine.outputs_to_host)
        d2h = sum(ch for _, ch, _ in OUTPUT_SPECS) * 4 * n_rays
        h2d = frames[0]["rays_h"].numel() * 4 + (0 if args.primary_only else env_h.numel() * 4) + (24 * 16 + 16) * 4
        e2e = {"value": samples_per_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
               "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    other = {}
    if world == 1 and args.config == 3 and not args.no_other_configs:
        for ci in (1, 2):
            a = argparse.Namespace(**vars(args))
            a.config, a.spp, a.gi, a.primary_only = ci, CONFIGS[ci]["spp"], CONFIGS[ci]["gi"], CONFIGS[ci]["primary_only"]
            ms_o, st_o, cn_o, _ = run_arm(a, False, 3, 2, True)
            roof_o, _, _, stages_o = roofline_of(a, st_o, cn_o, ms_o / 3)
            n_s = n_rays * (1 if a.primary_only else a.spp)
            other[f"configs[{ci}]"] = {"workload": workload_config(a, 1)["workload"], "value": n_s * 3 / (ms_o * 1e-3),
                                       "unit": "rays/s" if a.primary_only else UNIT, "ms_per_step": ms_o / 3, "steps": 3, "warmup": 2,
                                       "stages_ms": stages_o,
                                       "roofline": {k: roof_o[k] for k in ("kernel", "launch_ms", "achieved", "frac", "achieved_as_stored",
                                                                          "algorithmic_bytes_per_launch", "units_per_launch")}}

    if rank == 0:
        roofline, c, cp, stages = roofline_of(args, stage_ms, cnts, ms_total / args.steps)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "stages_ms": stages, "counters_per_frame": c, "counters_primary_stage": cp,
            "ms_per_frame": ms_total / args.steps,
            "metric_note": "shaded samples = primary rays x spp as BASELINE.json defines the metric; of them "
                           "hit_rays x spp are shading samples of pixels that hit the body and `secondary_rays` trace a ray "
                           "(counters_per_frame)",
        }
        if world == 1 and stage_ms:
            per_frame = [sum(s[k] for k in ("setup", "primary", "resample", "shade", "composite") if s[k] >= 0) for s in stage_ms]
            line["frame_cost_spread_ms"] = {"min": min(per_frame), "max": max(per_frame),
                                            "frames": main_frames,
                                            "ms": [round(v, 1) for v in per_frame]}
        if world > 1:
            line["frame_transport"] = collector.transport
        if other:
            line["other_configs"] = other
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        collector.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
