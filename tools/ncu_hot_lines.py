"""Source lines of one kernel in an `ncu --set full --import-source on` report, ranked by warp-stall samples.
Usage: python tools/ncu_hot_lines.py <report.ncu-rep> <kernel regex> [top N] > profiles/<tag>_<kernel>_hot_lines.csv"""
import csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{pat}"],
                     check=True, capture_output=True, text=True).stdout
cur, hdr, agg = None, None, {}
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and cur and len(r) == len(hdr) and r[2] == "-":  # a source-level row (no SASS address)
        try:
            inst, samp = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        if inst or samp:
            k = (cur, int(r[0]))
            a = agg.get(k, [0, 0, r[1].strip()])
            a[0] += inst; a[1] += samp
            agg[k] = a
ti, ts = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
w = csv.writer(sys.stdout)
w.writerow(["file", "line", "pct_of_instructions", "pct_of_stall_samples", "source"])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    w.writerow([k[0], k[1], round(100 * v[0] / ti, 2), round(100 * v[1] / ts, 2), v[2][:140]])
