"""The C++ facade (include/gsm/DepthFirstRenderer.hpp) compiles against the C ABI and runs the reference's
testDepthFirstPipelineStages scene. CPU: it must compile, link and fail loudly with deviceNotAvailable."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "facade_driver")
LIBDIR = os.path.join(ROOT, "gsm_renderer_b200", "lib")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "facade_driver.cpp")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                    "-L", LIBDIR, "-lgsm_b200", f"-Wl,-rpath,{LIBDIR}"], check=True)


def test_facade_compiles_and_fails_loudly_without_gpu():
    import torch
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 3 and "RendererError" in r.stdout  # deviceNotAvailable, no silent CPU path
    else:
        assert r.returncode == 0, r.stdout


@pytest.mark.gpu
def test_facade_pipeline_stages_scene_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    assert "visible=872 instances=2656" in r.stdout  # the oracle's counters for this scene (tests/test_oracle_kats.py)
    # the same scene through gsm::GlobalRenderer: the oracle's GlobalRenderer frame gives 872 visible, 2 347 assignments, 499 tiles
    assert "global visible=872 assignments=2347 overflow=0 activeTiles=499 sorted=1" in r.stdout, r.stdout
