"""The C-ABI library loads and exports every symbol include/vggsfm_b200.h declares (no GPU compute)."""
import os
import re

from vggsfm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vggsfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vgg_[a-z0-9_]+)\s*\(", src)) - {"vgg_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    decl = declared_symbols()
    assert decl, "no declarations parsed"
    for name in decl:
        assert hasattr(L, name), f"libvggsfm_b200.so does not export {name}"
    assert sorted(_lib.EXPORTS) == decl
    assert L.vgg_version() >= 100


def test_host_only_entry_points():
    import ctypes
    L = _lib.lib()
    dc, ns = ctypes.c_int(), ctypes.c_int()
    assert L.vgg_ba_dims(1, 2, ctypes.byref(dc), ctypes.byref(ns)) == 0 and (dc.value, ns.value) == (6, 2)
    assert L.vgg_ba_dims(0, 1, ctypes.byref(dc), ctypes.byref(ns)) == 0 and (dc.value, ns.value) == (7, 0)
    assert L.vgg_ba_dims(7, 1, ctypes.byref(dc), ctypes.byref(ns)) != 0
    assert L.vgg_ba_camrec_len(1, 1) == 8 + 36
    o = _lib.BAOptions()
    L.vgg_ba_default_options(ctypes.byref(o))
    assert o.max_num_iterations == 100 and o.gradient_tolerance == 1e-4
