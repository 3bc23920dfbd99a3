#!/usr/bin/env python3
"""Summarise an .ncu-rep (run where ncu exists; no GPU needed): headline metrics + the source lines
that execute the most instructions / collect the most stall samples. Writes markdown to stdout.
usage: ncu_summary.py report.ncu-rep [top_n_lines=25]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__cycles_elapsed.avg.per_second",
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
    for k, row in enumerate(raw[2:]):
        name = row[hdr.index("Kernel Name")]
        print(f"## launch {k}: `{name}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"| {w} | {row[i]} | {units[i]} |")
        stalls = [(float(row[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and row[i]]
        if not stalls:
            stalls = [(float(row[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("warp_active.pct") and row[i]]
        print("\nTop stall reasons: " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in sorted(stalls, reverse=True)[:6]) + "\n")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    # the cuda view: rows start with line number + source text
    h = None
    lines = []
    first_file, seen_files = None, 0
    for r in src:
        if r and r[0] == "Line No":
            h = r
            continue
        if h and len(r) == len(h) and r[0].isdigit():
            try:
                lines.append((int(r[h.index("Instructions Executed")]), int(r[h.index("# Samples")]), int(r[0]), r[1].strip()[:120]))
            except ValueError:
                pass
        if r and r[0] == "File Path" and r[1].endswith(".cu") and lines and r[1] == first_file and seen_files > 1:
            break  # first launch only
        if r and r[0] == "File Path":
            if first_file is None:
                first_file = r[1]
            seen_files += 1
    tot_i = sum(x[0] for x in lines) or 1
    tot_s = sum(x[1] for x in lines) or 1
    print(f"## hottest source lines (first launch; {tot_i} warp instructions, {tot_s} stall samples)\n")
    print("| line | warp inst | % inst | % samples | source |\n|---|---|---|---|---|")
    for n, s, ln, text in sorted(lines, reverse=True)[:top]:
        print(f"| {ln} | {n} | {100 * n / tot_i:.1f} | {100 * s / tot_s:.1f} | `{text}` |")


if __name__ == "__main__":
    main()
