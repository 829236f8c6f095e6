#!/usr/bin/env python
"""Instruction mix of a kernel of libia_b200.so from cuobjdump -sass (static SASS, no GPU needed): opcode histogram
and the Blackwell / tensor-core / wide-load mnemonics that prove what the build contains.
usage: python scripts/sass_mix.py [lib] [kernel substring]"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "intrinsicavatar_b200", "libia_b200.so")
want = sys.argv[2] if len(sys.argv) > 2 else "k_shade_wfILb1ELi0"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, mix = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        mix[cur][m.group(1)] += 1
for name, c in mix.items():
    if want not in name:
        continue
    n = sum(c.values())
    print(f"## {name}: {n} SASS instructions")
    base = collections.Counter()
    for op, k in c.items():
        base[op.split(".")[0]] += k
    print("by opcode:", ", ".join(f"{op} {k}" for op, k in base.most_common(24)))
    for title, pat in (("tensor core (mma.sync -> HMMA)", r"^HMMA"), ("global loads", r"^LDG"), ("global stores / atomics", r"^(STG|RED|ATOMG)"),
                       ("shared memory", r"^(LDS|STS|ATOMS)"), ("local memory (spills)", r"^(LDL|STL)"), ("shuffles", r"^SHFL"),
                       ("MUFU", r"^MUFU"), ("barriers", r"^BAR")):
        sel = {op: k for op, k in c.items() if re.match(pat, op)}
        print(f"{title}: " + (", ".join(f"{op} x{k}" for op, k in sorted(sel.items())) or "none"))
