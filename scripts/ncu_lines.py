"""Per-CUDA-source-line view of an ncu report: warp-stall samples and executed warp instructions, summed over the SASS of
each line.  Usage: python scripts/ncu_lines.py <rep> [kernel regex] [top N]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kern:
    cmd += ["--kernel-name", "regex:" + kern]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, hdr = None, None
agg = {}
tot_s = tot_i = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":  # a CUDA source line row carries the line's totals
        try:
            s, e = int(r[i_s]), int(r[i_e])
        except ValueError:
            continue
        key = (cur_file, int(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0])
        a[0] += s
        a[1] += e
        tot_s += s
        tot_i += e
print(f"total samples {tot_s}, warp instructions {tot_i}")
for (f, ln, src), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/max(1,tot_s):5.1f}%  inst {100*e/max(1,tot_i):5.1f}%  {f}:{ln}  {src}")
