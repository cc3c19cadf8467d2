"""Top stall lines of one captured kernel: python tools/ncu_src.py REP SKIP [TOP]
Reads `ncu -i REP --page source --csv --print-source cuda,sass` for launch number SKIP and prints the CUDA source
lines sorted by warp-stall samples, with instruction counts and the dominant stall reasons."""
import csv, subprocess, sys, io
rep, skip = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1", "--print-source", "cuda,sass"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, ix, lines = "", None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
    elif hdr and r[0].strip().isdigit() and len(r) == len(hdr):
        lines.append((fname, r))
print(func[:120])
S = ix["# Samples"]
I = ix["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
f = lambda x: float(x) if x not in ("", "-") else 0.0
tot = sum(f(r[S]) for _, r in lines)
toti = sum(f(r[I]) for _, r in lines)
print("total samples", tot, "warp instructions", toti)
agg = {}
for _, r in lines:
    for h in stalls:
        agg[h] = agg.get(h, 0) + f(r[ix[h]])
print("stall mix:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
lines.sort(key=lambda fr: -f(fr[1][S]))
for fn, r in lines[:top]:
    st = sorted(((f(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{fn[:16]:>16}:{r[0]:>4} {100*f(r[S])/max(tot,1):5.1f}% inst {100*f(r[I])/max(toti,1):5.1f}%  {r[1].strip()[:100]:100s} " + " ".join(f"{n}:{v:.0f}" for v, n in st if v > 0))
