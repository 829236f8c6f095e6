#!/usr/bin/env python
"""Static instruction mix of a kernel between its CTA barriers (cuobjdump -sass): spills (LDL/STL), HMMA, LDG, SHFL, LDS/STS.
usage: python scripts/sass_phases.py lib.so mangled-name-prefix"""
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for pz in re.split(r"\n\s*Function : ", txt):
    name = pz.split("\n", 1)[0]
    if not name.startswith(sys.argv[2]):
        continue
    ins = [l for l in pz.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
    print(name, len(ins), "instructions")
    bars = [i for i, l in enumerate(ins) if "BAR.SYNC" in l]
    segs = [0] + bars + [len(ins)]
    for a, b in zip(segs[:-1], segs[1:]):
        seg = ins[a:b]
        n = lambda k: sum(k in l for l in seg)
        if b - a > 150:
            print(f"  {a:6d}-{b:6d}: n={b - a:5d} LDL={n('LDL'):3d} STL={n('STL'):3d} HMMA={n('HMMA'):4d} LDG={n('LDG'):3d} "
                  f"SHFL={n('SHFL'):3d} LDS={n('LDS'):3d} STS={n('STS'):3d} MUFU={n('MUFU'):3d}")
