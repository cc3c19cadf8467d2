"""exp(-0.5h * p) on the XU pipe (MUFU.EX2) against the canonical polynomial, all 65 536 half inputs (gsm_probe_math ops 12, 14-16):
how many inputs the unguarded form gets wrong, how many the guard sends to the polynomial, and that the guarded form is bit-equal."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from gsm_renderer_b200.renderer import probe_math

bits = np.arange(65536, dtype=np.uint16)
h = bits.view(np.float16)
notnan = ~np.isnan(h)
ref = probe_math(12, bits)
guarded, raw, flag = probe_math(14, bits), probe_math(15, bits), probe_math(16, bits)
bad_raw = notnan & (raw != ref)
bad_guarded = notnan & (guarded != ref)
inrange = notnan & (h >= 0) & (h <= 35)
print("inputs in [0, 35]:", int(inrange.sum()))
print("unguarded MUFU differs from the canonical polynomial on", int(bad_raw.sum()), "inputs;", int((bad_raw & inrange).sum()), "in [0, 35]")
print("guard sends", int((flag == 1)[notnan].sum()), "inputs to the polynomial;", int(((flag == 1) & inrange).sum()), "in [0, 35]")
print("unguarded-wrong inputs not caught by the guard:", int((bad_raw & (flag == 0)).sum()))
print("guarded form differs on", int(bad_guarded.sum()), "inputs")
rng = np.random.default_rng(0)
sh = rng.permutation(bits)
ok = np.array_equal(probe_math(14, sh)[~np.isnan(sh.view(np.float16))], probe_math(12, sh)[~np.isnan(sh.view(np.float16))])
print("shuffled pairs equal:", ok)
tuned = probe_math(17, bits)
print("tuned guard-free form differs on", int((notnan & (tuned != ref)).sum()), "inputs; shuffled pairs equal:",
      np.array_equal(probe_math(17, sh)[~np.isnan(sh.view(np.float16))], probe_math(12, sh)[~np.isnan(sh.view(np.float16))]))
for i in np.nonzero(bad_raw)[0][:12]:
    print(f"  p={float(h[i]):.6g} bits={i:#06x} canonical={ref[i]:#06x} mufu={raw[i]:#06x} guard={int(flag[i])}")
