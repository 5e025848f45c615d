"""Host-side mirrors against golden vectors produced by the UNMODIFIED reference (tools/make_golden_host.py runs
vggsfm.utils.align, vggsfm.models.utils.get_EFP and vggsfm.models.triangulator.find_best_initial_pair in the build
container).  CPU only; float results to 1e-12 (float64) / 1e-5 relative (float32 get_EFP), masks and thresholds exact.
Also the reference's own property test for the alignment (vggsfm/utils/align.py:255-300): a random similarity is
recovered."""
import os

import numpy as np
import pytest
import torch

from vggsfm_b200 import align
from vggsfm_b200.triangulator import find_best_initial_pair, get_EFP

G = os.path.join(os.path.dirname(__file__), "golden")


def test_align_matches_reference_goldens():
    d = np.load(os.path.join(G, "host_align.npz"))
    for i in range(int(d["n"])):
        src, tgt = torch.from_numpy(d[f"src{i}"]), torch.from_numpy(d[f"tgt{i}"])
        R, T, s = align.align_camera_extrinsics(src, tgt, estimate_scale=bool(d[f"est{i}"]))
        assert R.shape == (1, 3, 3) and T.shape == (1, 3)
        assert np.allclose(R.numpy(), d[f"R{i}"], atol=1e-12) and np.allclose(T.numpy(), d[f"T{i}"], atol=1e-11)
        assert abs(float(s) - float(d[f"s{i}"])) <= 1e-12 * max(1.0, abs(float(d[f"s{i}"])))
        out = align.apply_transformation(src, R, T, s)
        assert np.allclose(out.numpy(), d[f"applied{i}"], atol=1e-11)
        aR, aT = align.apply_transformation(src, R, T, s, return_extri=False)
        assert torch.equal(torch.cat([aR, aT[..., None]], -1), out)


def test_align_recovers_a_random_similarity():
    torch.manual_seed(3)
    for _ in range(20):
        B = 10
        q, _ = torch.linalg.qr(torch.randn(B, 3, 3, dtype=torch.float64))
        src = torch.cat([q, torch.randn(B, 3, 1, dtype=torch.float64)], -1)
        qa, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64))
        t_true, s_true = torch.randn(1, 3, dtype=torch.float64), float(torch.rand(()) * 100)
        tgt = align.apply_transformation(src, qa[None], t_true, s_true)
        R, T, s = align.align_camera_extrinsics(src, tgt)
        assert torch.allclose(align.apply_transformation(src, R, T, s), tgt, atol=1e-9)
        assert torch.allclose(R[0], qa, atol=1e-10) and abs(float(s) - s_true) < 1e-9 * max(1.0, s_true)


class _Cams:
    pass


def test_get_EFP_matches_reference_goldens():
    d = np.load(os.path.join(G, "host_get_EFP.npz"))
    for i in range(int(d["n"])):
        c = _Cams()
        c.focal_length, c.R, c.T = torch.from_numpy(d[f"focal{i}"]), torch.from_numpy(d[f"R{i}"]), torch.from_numpy(d[f"T{i}"])
        S = c.R.shape[0]
        E, K = get_EFP(c, torch.from_numpy(d[f"size{i}"]), 1, S, default_focal=bool(d[f"default{i}"]))
        assert E.shape == (1, S, 3, 4) and K.shape == (1, S, 3, 3) and K.dtype == torch.float32
        assert np.array_equal(E.numpy(), d[f"E{i}"])
        assert np.allclose(K.numpy(), d[f"K{i}"], rtol=1e-6, atol=0)
        assert torch.equal(c.focal_length, torch.from_numpy(d[f"focal{i}"]))      # the caller's camera is not modified


def test_find_best_initial_pair_matches_reference_goldens():
    d = np.load(os.path.join(G, "host_find_best_initial_pair.npz"))
    seen = set()
    for i in range(int(d["n"])):
        tot, thr = find_best_initial_pair(torch.from_numpy(d[f"geo{i}"]), torch.from_numpy(d[f"che{i}"]),
                                          torch.from_numpy(d[f"tri{i}"]), 16)
        assert int(thr) == int(d[f"thr{i}"]) and np.array_equal(tot.numpy(), d[f"tot{i}"])
        seen.add(int(thr))
    assert {16, 1} <= seen                                                        # first-try accept and the floor are covered
