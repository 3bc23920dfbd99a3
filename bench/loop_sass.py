#!/usr/bin/env python3
"""Static listing of the headline kernel's march loop (no GPU needed): cuobjdump -sass of vokselis_b200/libvokselis_rt.so,
the loop of raycast_kernel<M1, QUAD, u8, SKIP, !DBG, !CLIP> cut into head / leap fast path / rejoin stub / sample path / tail
with the issue slots of each part. Writes profiles/loop_sass_<round>.md (usage: loop_sass.py [r02]). Run after `make`."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r02"
KERNEL = "raycast_kernelILi1ELi4ELi0ELb1ELb0ELb0E"

sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "vokselis_b200" / "libvokselis_rt.so")], capture_output=True, text=True).stdout
ins, on = [], False
for line in sass.splitlines():
    if "Function :" in line:
        on = KERNEL in line
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?)\s*;", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2)))
target = lambda t: int(t.split("0x")[-1], 16)
itex = [i for i, (a, t) in enumerate(ins) if t.startswith("TEX")][-1]
iback = next(i for i in range(itex, len(ins)) if "BRA" in ins[i][1] and target(ins[i][1]) < ins[i][0])
istart = next(i for i, (a, t) in enumerate(ins) if a == target(ins[iback][1]))
ihead = next(i for i in range(istart, itex) if "BRA" in ins[i][1])          # branch to the sample path (d == 0)
isample = next(i for i, (a, t) in enumerate(ins) if a == target(ins[ihead][1]))
ifast = next(i for i in range(ihead + 1, isample) if ins[i][1].startswith("@") and "BRA" in ins[i][1])  # fast exit of leap_steps
istub = next(i for i, (a, t) in enumerate(ins) if a == target(ins[ifast][1]))
jstub = next(i for i in range(istub, len(ins)) if "BRA" in ins[i][1] or "BSYNC" in ins[i][1])
ijoin = next(i for i in range(itex, iback) if "BSYNC" in ins[i][1])
parts = [("head: position, voxel coordinate, brick coordinate, distance lookup", istart, ihead),
         ("leap, fast path: leap length (leap_count) + closed-form t (leap_steps) while t stays in the cached binade", ihead + 1, ifast),
         ("leap, rejoin stub", istub, jstub),
         ("sample: two point fetches, trilinear interpolation, transfer function, palette, compositing, termination flag", isample, ijoin),
         ("tail: loop condition", ijoin + 1, iback)]
out = [f"# March loop of the headline kernel, SASS of the {ROUND} final build (sm_100a)\n",
       "`raycast_kernel<M1, QUAD, u8, SKIP, !DBG, !CLIP>` from `cuobjdump -sass vokselis_b200/libvokselis_rt.so` (`bench/loop_sass.py`, no GPU needed).",
       "One warp-level iteration issues the head, then the leap path if any lane sits in an empty brick, then the sample path if any lane",
       "does not, then the tail (DESIGN.md section 7). The slow path of `leap_steps` (a leap that leaves the binade of `t`) lies between the",
       "fast path and the stub and is not listed.\n",
       "| part | issue slots |", "|---|---|"]
for name, a, b in parts:
    out.append(f"| {name} | {b - a + 1} |")
out.append("")
for name, a, b in parts:
    out.append(f"## {name} ({b - a + 1})\n\n```")
    out += [f"/*{ad:04x}*/  {t}" for ad, t in ins[a:b + 1]]
    out.append("```\n")
(ROOT / "profiles" / f"loop_sass_{ROUND}.md").write_text("\n".join(out))
print({name.split(":")[0]: b - a + 1 for name, a, b in parts})
