"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) as markdown: per-kernel totals and
the collapsed launch sequence. usage: launch_list.py launches.csv "command that was profiled" > launches.md"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, ig, iv, iu = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    name = re.sub(r"^void (vkrt::)?(<unnamed>::|\(anonymous namespace\)::)?", "", r[ik])
    name = re.sub(r"\(vkrt::\w+\)$|\(.*\)$", "", name).replace("(int)", "").replace("(bool)", "")
    v = float(r[iv].replace(",", ""))
    us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
    launches.append((name, r[ig].replace(" ", ""), us))
tot = sum(u for _, _, u in launches)
print(f"# ncu launch list of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` (B200; --metrics gpu__time_duration.sum --clock-control none)\n")
print(f"Per-launch times under the profiler are cold-cache and serialised: compare SHARES, not absolutes. {len(launches)} launches, {tot / 1e3:.1f} ms in total.\n")
agg = OrderedDict()
for n, g, u in launches:
    k = f"{n} grid {g}" if n.startswith("raycast_kernel") else n
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += u
print("## per-kernel totals\n\n| kernel | launches | total | share |\n|---|---|---|---|")
for k, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {u:.1f} us | {100 * u / tot:.1f} % |")
print("\n## sequence (consecutive identical kernels collapsed)\n\n| kernel | grid | count | mean duration |\n|---|---|---|---|")
i = 0
while i < len(launches):
    j = i
    while j < len(launches) and launches[j][:2] == launches[i][:2]:
        j += 1
    print(f"| `{launches[i][0]}` | {launches[i][1]} | {j - i} | {sum(u for _, _, u in launches[i:j]) / (j - i):.1f} us |")
    i = j
