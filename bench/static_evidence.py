#!/usr/bin/env python3
"""Static evidence of the built library (no GPU needed): ptxas -v register/spill table from build/obj/*.ptxas.log and
the memory/texture/special-function mnemonics per kernel from cuobjdump -sass. Writes profiles/ptxas_<round>.md and
profiles/sass_evidence_<round>.md (usage: static_evidence.py [r02]). Run after `make`."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r02"  # file tag: profiles/ptxas_<ROUND>.md, profiles/sass_evidence_<ROUND>.md


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"vkrt::\(anonymous namespace\)::|vkrt::|\(vkrt::RenderArgs\)|\(vkrt::PartialArgs\)|\(int\)|\(bool\)|void ", "", x) for x in out]


def ptxas_table():
    rows = []
    for log in sorted((ROOT / "build" / "obj").glob("*.ptxas.log")):
        txt = log.read_text()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt, re.S):
            rows.append((m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
    names = demangle([r[0] for r in rows])
    lines = [f"# ptxas -v summary (sm_100a), {ROUND} final build\n", "| kernel | registers | stack B | spill st/ld B |", "|---|---|---|---|"]
    for n, r in sorted(zip(names, rows)):
        lines.append(f"| `{n}` | {r[1]} | {r[2]} | {r[3]}/{r[4]} |")
    (ROOT / "profiles" / f"ptxas_{ROUND}.md").write_text("\n".join(lines) + "\n")
    return len(rows)


def sass_table():
    sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "vokselis_b200" / "libvokselis_rt.so")], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    want = re.compile(r"\b(FFMA2|FMUL2|FADD2|TLD4|TEX|TLD|LDG|STG|ATOMG|REDG|RED|REDUX|VOTEU?|MUFU|F2I|I2FP|I2F|MEMBAR|UTMA\w*|HMMA|IMMA|UTC\w*|LDS|STS|BAR)(\.[A-Z0-9_.]+)?")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            mm = want.search(line)
            if mm:
                op = mm.group(1) + (mm.group(2) or "")
                op = re.sub(r"\.(CONSTANT|STRONG|SYS|GPU|E)\b", lambda k: "." + k.group(1) if k.group(1) in ("E", "CONSTANT") else "", op)
                op = re.sub(r"^(MUFU|F2I|I2F|I2FP|VOTEU?|REDUX|ATOMG|REDG|MEMBAR)\..*", r"\1", op)
                per[cur][op] += 1
    names = demangle(list(per))
    lines = [f"# SASS evidence (cuobjdump -sass vokselis_b200/libvokselis_rt.so, sm_100a), {ROUND} final build\n",
             "Static instruction counts of the memory/texture/special-function mnemonics per kernel. `TLD4.R` = the `tld4.r.a2d` gathers of the GATHER layout,",
             "`TEX` = tex3D fetches of the TEXTURE layout and the two point fetches per sample of the QUAD layout (`raycast_kernel<1, 4, ...>`, the bench headline), `FFMA2` / `FMUL2` / `FADD2` = Blackwell packed fp32 (two IEEE operations per issue slot), `LDG.E.128` = the interleaved 16-B texel of the BRICKED layout. No tensor-core (HMMA/UTC*MMA) or TMA (UTMA*)",
             "mnemonics and no shared memory or barriers: the path is not a contraction and its loads are per-ray gathers, not tiles.\n",
             "| kernel | mnemonics |", "|---|---|"]
    for n, k in sorted(zip(names, per)):
        lines.append(f"| `{n}` | " + ", ".join(f"{op} x{c}" for op, c in sorted(per[k].items())) + " |")
    (ROOT / "profiles" / f"sass_evidence_{ROUND}.md").write_text("\n".join(lines) + "\n")
    return len(per)


if __name__ == "__main__":
    print("ptxas entries:", ptxas_table(), "sass kernels:", sass_table())
