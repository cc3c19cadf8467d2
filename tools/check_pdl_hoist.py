"""Every kernel of libgsm_b200.so: global loads / atomics the compiler scheduled BEFORE the programmatic-dependent-launch wait
(SASS `ACQBULK`). A `const T* __restrict__` load is `invariant` for the compiler and may legally be hoisted above the inline-asm
wait -- reading the predecessor kernel's output before it exists (found as a sort that did nothing: bucketsort.cu r2).
Usage: python tools/check_pdl_hoist.py [lib]; exit code 1 if any kernel has one."""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gsm_renderer_b200/lib/libgsm_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
bad = 0
name, before, seen = None, [], False
def flush():
    global bad
    if name and seen and before:
        bad += 1
        print(name)
        for l in before:
            print("   ", l)
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        flush()
        name, before, seen = m.group(1), [], False
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if not m or name is None:
        continue
    ins = m.group(2).strip()
    if "ACQBULK" in ins:
        seen = True
        # only the first wait matters
        name_done = name
        flush()
        name = None
        continue
    if not seen and re.search(r"\b(LDG|LD\.|ATOMG|ATOM\.|REDG|RED\.)", ins):
        before.append(f"{m.group(1)}: {ins}")
print("kernels with loads before the wait:", bad)
sys.exit(1 if bad else 0)
