"""CPU: the C-ABI library builds, loads, and exports every symbol include/mvsb200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "mvsb200.h")).read()
    return sorted(set(re.findall(r"MVSB200_API\s+[\w\s\*]+?\b(mvsb200_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for n in ("mvsb200_build_cost_volume", "mvsb200_conv3d", "mvsb200_depth_regress", "mvsb200_vis_fuse",
              "mvsb200_mvs_relative_proj", "mvsb200_vis_homography_params", "mvsb200_last_error"):
        assert n in names


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md shows, for every symbol of the C ABI, what it replaces in the reference (or that it is a support call)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared() if n not in doc]
    assert not missing, missing


def test_library_builds_and_exports_every_declared_symbol():
    from wild_deep_mvs_b200 import build, _lib
    so = build.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    for name in _declared():
        assert hasattr(lib, name), name
    assert set(_declared()) == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert _lib.load().mvsb200_abi_version() == _lib.ABI_VERSION == 6


def test_struct_layouts_match_header():
    from wild_deep_mvs_b200 import _lib
    # 10 ints + 2*16 ints + one long long (8-byte aligned)
    assert ctypes.sizeof(_lib.CostVolumeDesc) == (10 + 32) * 4 + 8
    assert ctypes.sizeof(_lib.Conv3dDesc) == 15 * 4   # 14 shape / epilogue ints + static_params


def test_argument_validation_without_gpu():
    """Descriptor validation happens before any CUDA call, so it is testable on a CPU box."""
    from wild_deep_mvs_b200 import _lib
    lib = _lib.load()
    d = _lib.Conv3dDesc()
    d.B, d.D, d.H, d.W, d.Cin, d.Cout = 1, 8, 8, 8, 8, 8
    d.kd = d.kh = d.kw = 5
    d.stride = 1
    o = [ctypes.c_int() for _ in range(3)]
    rc = lib.mvsb200_conv3d_out_shape(ctypes.byref(d), *[ctypes.byref(v) for v in o])
    assert rc == -1 and b"kernel extents" in lib.mvsb200_last_error()
    d.kd = d.kh = d.kw = 3
    d.stride, d.transposed = 2, 1
    assert lib.mvsb200_conv3d_out_shape(ctypes.byref(d), *[ctypes.byref(v) for v in o]) == 0
    assert [v.value for v in o] == [16, 16, 16]
    d.transposed = 0
    assert lib.mvsb200_conv3d_out_shape(ctypes.byref(d), *[ctypes.byref(v) for v in o]) == 0
    assert [v.value for v in o] == [4, 4, 4]
    rc = lib.mvsb200_depth_regress(None, 1, 1, 1, 1, 0, None, None, 0, None, None, None, None, None)
    assert rc == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wild_deep_mvs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f
