"""A short run of the randomised parity sweep (tools/fuzz_parity.py: random clouds, cameras, resolutions, precisions, key
widths, mono and stereo; every white-box buffer and pixel against the oracle, bit for bit)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_randomised_parity_sweep():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "40", "2026"], capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert r.returncode == 0 and "40 cases, 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_randomised_strip_sweep():
    """tools/fuzz_strips.py: random shard counts and uneven strip partitions, image + depth equal to the single-GPU frame."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_strips.py"), "16", "11"], capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert r.returncode == 0 and "16 cases, 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_randomised_foveated_copy_sweep():
    """tools/fuzz_copy.py: random drawables, formats, viewports and rate maps through gsm_stereo_copy, every byte against the oracle."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_copy.py"), "60", "3"], capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert r.returncode == 0 and "60 cases, 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
