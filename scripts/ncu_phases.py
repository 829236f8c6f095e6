#!/usr/bin/env python
"""Split the source page of an .ncu-rep of k_shade_wf at its CTA barriers (BAR.SYNC) and report, per segment: SASS
instructions, warp instructions executed, stall samples and their top reasons, global L1 tag requests, shared-memory
wavefronts, a few marker opcodes (HMMA = tensor-core geometry phase, LDG.E.128 = voxel gathers, ...).  A warp waiting at a
barrier is sampled at the first instruction AFTER it, so a barrier wait is listed with the segment that follows it
(column `first`: samples on the segment's first instruction).

usage: python scripts/ncu_phases.py report.ncu-rep [kernel-name substring] [--warpsync]
--warpsync also cuts at WARPSYNC instructions (the stages of the tensor-core phases inside one barrier interval)."""
import csv
import io
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
cuts = ("BAR.SYNC", "WARPSYNC") if "--warpsync" in sys.argv else ("BAR.SYNC",)
rep = args[0]
want = args[1] if len(args) > 1 else "k_shade_wf"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": line.split(",", 1)[1].strip('",'), "lines": []}
        blocks.append(cur)
    elif cur is not None:
        cur["lines"].append(line)
for b in blocks:
    if want not in b["name"]:
        continue
    rows = list(csv.reader(io.StringIO("\n".join(b["lines"]))))
    hdr, rows = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(r, h):
        try:
            return float(r[col[h]])
        except Exception:
            return 0.0
    segs, seg = [], {"rows": []}
    for r in rows:
        seg["rows"].append(r)
        if any(c in r[col["Source"]] for c in cuts):
            segs.append(seg)
            seg = {"rows": []}
    segs.append(seg)
    total = sum(num(r, "# Samples") for r in rows)
    print("==", b["name"][:90], "total samples", int(total))
    print(f"{'seg':>3} {'sass':>6} {'warp-instr':>11} {'samples':>9} {'share':>6} {'first':>8} {'L1 tags':>10} {'smem wf':>10}  markers / top stalls")
    for i, s in enumerate(segs):
        rs = s["rows"]
        if not rs:
            continue
        smp = sum(num(r, "# Samples") for r in rs)
        ins = sum(num(r, "Instructions Executed") for r in rs)
        tags = sum(num(r, "L1 Tag Requests Global") for r in rs)
        swf = sum(num(r, "L1 Wavefronts Shared") for r in rs)
        first = num(rs[0], "# Samples")
        ops = {}
        for r in rs:
            op = r[col["Source"]].split()[0] if r[col["Source"]].split() else ""
            if op.startswith("@"):
                op = r[col["Source"]].split()[1]
            for m in ("HMMA", "LDG.E.128", "LDG.E.64", "LDG.E.ENL2.256", "ATOMS", "SHFL", "LDS", "MUFU", "RED", "ATOMG"):
                if op.startswith(m):
                    ops[m] = ops.get(m, 0) + 1
        st = sorted(((sum(num(r, h) for r in rs), h) for h in stall_cols), reverse=True)[:3]
        sts = ", ".join(f"{h[6:]} {100 * v / max(smp, 1):.0f}%" for v, h in st)
        mk = " ".join(f"{k}x{v}" for k, v in sorted(ops.items()))
        if smp < 0.002 * total and ins < 1e6:
            continue
        print(f"{i:>3} {len(rs):>6} {ins:>11.3e} {int(smp):>9} {100 * smp / total:>5.1f}% {int(first):>8} {tags:>10.3e} {swf:>10.3e}  {mk} | {sts}")
