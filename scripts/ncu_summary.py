"""Turns an ncu report (gpurun_out/*.ncu-rep) into the committed summary under profiles/: per-kernel duration, DRAM
bytes, issue/stall picture and the hottest SASS lines.  Usage: python scripts/ncu_summary.py <rep> <out.md> [title]"""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
with open(out, "w") as f:
    f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` (one pass of bench.py's workload; "
            "cold-cache, serialised replays: compare shares, not absolutes).\n\n")
    for r in data:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        f.write(f"## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {r[i]} | {units[i]} |\n")
        f.write("\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + name.split("::")[-1]],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        if len(srows) > 2:
            sh = srows[1]
            sd = [x for x in srows[2:] if len(x) == len(sh) and x[0].startswith('0x')]
            ia, isrc, ismp = sh.index("Instructions Executed"), sh.index("Source"), sh.index("# Samples")
            tot = sum(int(x[ismp]) for x in sd) or 1
            f.write("Hottest SASS (warp-stall samples):\n\n| samples | share | executed | SASS |\n|---|---|---|---|\n")
            for x in sorted(sd, key=lambda x: -int(x[ismp]))[:12]:
                f.write(f"| {x[ismp]} | {100 * int(x[ismp]) / tot:.1f}% | {x[ia]} | `{x[isrc].strip()[:80]}` |\n")
            f.write(f"\nTotal warp instructions {sum(int(x[ia]) for x in sd)}, SASS lines {len(sd)}.\n\n")
print("wrote", out)
