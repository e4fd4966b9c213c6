"""The C-ABI library loads and exports every symbol include/restir_b200.h declares; struct layouts
match the ctypes mirror. No compute calls (runs without a GPU)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from restir_embree_b200 import abi, renderer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "restir_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z_0-9]+)\s*\(", src)))


def test_header_and_mirror_list_the_same_symbols():
    assert _declared_functions() == sorted(abi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    L = renderer.load_library()
    for name in _declared_functions():
        assert hasattr(L, name), name
    assert L.rb_abi_version() == 2


def test_struct_layouts_match_the_header(built):
    names = ["RbMaterial", "RbSurface", "RbSceneDesc", "RbParams", "RbCamera", "RbTimings", "RbCreateInfo", "RbRay",
             "RbHit", "RbSceneStats", "RbImageStats"]
    prog = '#include <stdio.h>\n#include "restir_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = dict(l.split() for l in subprocess.check_output([exe]).decode().splitlines())
    for n in names:
        assert int(out[n]) == C.sizeof(getattr(abi, n)), n


def test_default_params_match_reference_defaults(built):
    L = renderer.load_library()
    p = abi.RbParams()
    L.rb_default_params(C.byref(p))
    q = abi.default_params()
    for name, _ in abi.RbParams._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if name == "bgColor":
            assert list(a) == list(b)
        else:
            assert a == b, name
    # P/ReSTIRIntegrator.cpp:13-33
    assert (p.M_Area, p.M_Brdf, p.spatialReuseNeighborCount, p.spatialPassCount, p.confidenceCap) == (1, 1, 5, 1, 20)
    assert p.spatialReuseRadius == 30.0 and abs(p.minNormalSimilarity - 0.85) < 1e-7


def test_create_fails_loudly_without_a_gpu(built):
    """No CPU fallback: without a CUDA device rb_create must return an error, not a working handle."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(renderer.RestirError):
        renderer.Renderer(16, 16)


def test_invalid_create_info_is_rejected(built):
    L = renderer.load_library()
    info = abi.RbCreateInfo()
    info.width, info.height, info.band_y0, info.band_y1 = 0, 16, 0, 16
    h = C.c_void_p()
    assert L.rb_create(C.byref(info), C.byref(h)) == -1
    assert b"invalid" in L.rb_last_error(None)
