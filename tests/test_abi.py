"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/gsm/gsm.h
declares, and reports the reference's RendererError cases without a GPU (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from gsm_renderer_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _native


def declared_symbols():
    syms = set()
    for h in ("gsm.h", "gsm_scene.h"):  # every header of the boundary that declares entry points
        src = open(os.path.join(ROOT, "include", "gsm", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms |= set(re.findall(r"\b(gsm_[a-z0-9_]+)\s*\(", src))
    return sorted(syms)


def test_library_exports_every_declared_symbol(native):
    lib = native.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libgsm_b200.so does not export {s}"
    assert set(syms) == set(native.EXPORTS), "python binding table and header diverged"
    assert lib.gsm_abi_version() == 2


def test_header_compiles_as_c_and_layouts_match():
    # the RendererTypes restatement must be plain C (SwiftPM C target) with the reference's sizes
    code = r'''
#include "gsm/gsm.h"
#include "gsm/gsm_scene.h"
#include "gsm/gsm_types.h"
#include <stdio.h>
#include <stddef.h>
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(GSMPackedWorldGaussian), sizeof(GSMPackedWorldGaussianHalf),
    sizeof(GSMGaussianRenderData), sizeof(GSMStereoTiledRenderData), sizeof(GSMDepthFirstHeader),
    sizeof(GSMStereoCameraUniforms), offsetof(GSMStereoCameraUniforms, rightViewMatrix), offsetof(GSMStereoCameraUniforms, sceneTransform));
  printf("%zu %zu %zu %zu\n", sizeof(gsm_config), sizeof(gsm_camera), sizeof(gsm_ply_info), sizeof(gsm_scene_info));
  return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write(code)
        exe = os.path.join(d, "t")
        subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), p, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    # BridgingTypes.h:57,66,75,206,275 and the offsets documented at :179,:204
    assert out[:8] == ["48", "32", "16", "32", "32", "416", "160", "352"]
    from gsm_renderer_b200 import _native
    assert int(out[8]) == C.sizeof(_native.gsm_config) and int(out[9]) == C.sizeof(_native.gsm_camera)
    assert int(out[10]) == C.sizeof(_native.gsm_ply_info) and int(out[11]) == C.sizeof(_native.gsm_scene_info)


def test_config_defaults_match_reference(native):
    cfg = native.gsm_config()
    native.lib().gsm_config_default(C.byref(cfg))
    # GaussianRendererProtocol.swift:211-218, DepthFirstRenderer.swift:48-49
    assert (cfg.maxGaussians, cfg.maxWidth, cfg.maxHeight) == (6_000_000, 1920, 1080)
    assert cfg.precision == 1 and cfg.gaussianColorSpace == 1
    assert cfg.depthSortKeyPrecision == 32 and cfg.tileIdPrecision == 16


def test_init_errors_without_compute(native):
    from gsm_renderer_b200.renderer import DepthFirstRenderer, RendererConfig, RendererError
    # DepthFirstRenderer.swift:51-56
    with pytest.raises(RendererError) as e:
        DepthFirstRenderer(config=RendererConfig(maxGaussians=30_000_001))
    assert e.value.case == "invalidGaussianCount"
    import torch
    if not torch.cuda.is_available():
        # DepthFirstRenderer.swift:58-61: no device -> deviceNotAvailable, never a silent CPU path
        with pytest.raises(RendererError) as e:
            DepthFirstRenderer(config=RendererConfig(maxGaussians=1000))
        assert e.value.case == "deviceNotAvailable"


def test_no_fallback_when_library_missing(monkeypatch, native):
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", "/nonexistent/libgsm_b200.so")
    with pytest.raises(native.NativeLibraryError):
        native.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gsm_renderer_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.replace("CPU oracle", "").replace("the oracle", "").lower() or \
                    not re.search(r"(import\s+oracle|from\s+oracle|#include\s+[\"<].*oracle|gsm_oracle|gsmo_)", txt), f


def test_no_global_access_before_dependent_launch_wait(native):
    """Every kernel of the chain reads its predecessor's output only after griddepcontrol.wait. nvcc may hoist a load through a
    `const T* __restrict__` pointer (an invariant load) above the inline-asm wait: tools/check_pdl_hoist.py scans the SASS of
    the built library for global loads / atomics scheduled before the first ACQBULK of a kernel (round 2: the depth sort's local
    pass read its bucket count before the scatter kernel had written it; create_instances read the frame header early)."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "check_pdl_hoist.py"),
                          os.path.join(root, "gsm_renderer_b200", "lib", "libgsm_b200.so")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:]
