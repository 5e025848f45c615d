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


def test_fails_loudly_without_the_library_or_a_gpu(monkeypatch, tmp_path):
    """No fallback path: a missing .so raises NativeLibraryMissing, CPU tensors raise, on every public entry."""
    import pytest
    import torch
    from vggsfm_b200 import bundle_adjustment as ba
    from vggsfm_b200 import corr, pose_refinement as pr, triangulation as tri
    x = torch.zeros(2, 4, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        tri.cam_from_img(x, torch.eye(3).repeat(2, 1, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        tri.triangulate_tracks(torch.zeros(2, 3, 4), x, track_vis=torch.ones(2, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ba.bundle_adjustment(torch.zeros(4, 3), torch.zeros(2, 3, 4), torch.eye(3).repeat(2, 1, 1), None, x, torch.ones(2, 4, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pr.pose_refinement_batched(torch.zeros(2, 3, 4, dtype=torch.float64), torch.zeros(2, 4, dtype=torch.float64),
                                   torch.zeros(4, 3), x, torch.ones(2, 4, dtype=torch.bool), torch.ones(2, dtype=torch.uint8), 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        corr.sample_features4d(torch.zeros(1, 3, 8, 8), torch.zeros(1, 5, 2))
    # missing library
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libvggsfm_b200.so"))
    with pytest.raises(_lib.NativeLibraryMissing, match="no CPU/PyTorch fallback"):
        _lib.lib()
    pcs = _lib.PoseOptions()
    assert pcs.min_inliers == 0                      # ctypes struct layout is importable without the library
