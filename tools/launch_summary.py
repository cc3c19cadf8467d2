"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, mean, min, max (us).
Usage: python tools/launch_summary.py gpurun_out/<tag>_launches.csv [skip_first_n_frames_of_kernel_regex]"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            agg.setdefault(d["Kernel Name"].split("(")[0][:80], []).append(float(d["Metric Value"].replace(",", "")) / 1000.0)
print(f"{'kernel':80s} {'n':>4s} {'mean':>8s} {'min':>8s} {'max':>8s}")
for k, v in agg.items():
    print(f"{k:80s} {len(v):4d} {sum(v) / len(v):8.1f} {min(v):8.1f} {max(v):8.1f}")
