"""CPU checks of host-side logic that both the CUDA path and the oracle rely on: constant-parameter masks, option
presets, the hypothesis-pair draw (same CPU RNG stream as the reference), shard ranges, the Ozaki plan's pair table."""
import itertools

import numpy as np
import pytest
import torch

from oracle import ba_oracle as bo
from vggsfm_b200 import bundle_adjustment as ba
from vggsfm_b200 import triangulation as tri


@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("rf,rk,gauge", [(True, True, True), (False, True, True), (True, False, False), (False, False, True)])
def test_param_const_matches_oracle(model, mode, rf, rk, gauge):
    S = 5
    const_pose = np.array([False, False, True, False, False])
    got = ba.default_param_const(S, model, mode, "cpu", rf, rk, gauge, torch.from_numpy(const_pose)).numpy().astype(bool)
    ref = bo.default_param_const(S, model, mode, refine_focal=rf, refine_extra=rk, gauge=gauge, const_pose=const_pose)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    dc, ns = ba.dims(model, mode)
    assert (dc, ns) == bo.dims(model, mode)
    if gauge:
        assert got[:6].all() and got[dc + 3]                 # first pose, x of the second translation
    assert got[2 * dc:2 * dc + 6].all()                      # the explicitly constant pose


def test_pair_draw_follows_the_reference_rng_stream():
    """triangulation.py:804-813: all pairs when C(S,2) < max_ransac_iters, else comb[randperm(len)[:iters]] on the CPU
    global generator -- a seeded run sees the pairs the reference would."""
    S = 8
    allp = tri.draw_ransac_pairs(S, 256)
    assert allp.shape == (28, 2) and sorted(map(tuple, allp)) == sorted(itertools.combinations(range(S), 2))
    S = 40
    torch.manual_seed(123)
    got = tri.draw_ransac_pairs(S, 128)
    torch.manual_seed(123)
    comb = np.array(list(itertools.combinations(range(S), 2)))
    ref = comb[torch.randperm(len(comb))[:128].numpy()]
    assert np.array_equal(got, ref) and len({tuple(p) for p in got}) == 128


def test_pad_tracks_and_camera_model_ids():
    assert [ba.pad_tracks(n) for n in (1, 16, 17, 4096)] == [16, 16, 32, 4096]
    assert ba.camera_model_id("SIMPLE_PINHOLE") == 0 and ba.camera_model_id("SIMPLE_RADIAL") == 1
    with pytest.raises(ValueError, match="is not supported yet"):
        ba.camera_model_id("OPENCV")


def test_ozaki_pair_table():
    """The order groups of csrc/syrk_i8.cu (restated): every (p,q) with p+q <= s+1 appears exactly once, at most four
    orders per group, at most 14 operand tiles per group."""
    for s in range(3, 8):
        pairs = []
        for t0 in range(2, s + 2, 4):
            t1 = min(t0 + 3, s + 1)
            grp = [(p, t - p) for t in range(t0, t1 + 1) for p in range(1, s + 1) if 1 <= t - p <= s]
            assert len({p + q for p, q in grp}) <= 4
            assert len({p for p, _ in grp}) + len({q for _, q in grp}) <= 14
            pairs += grp
        assert len(pairs) == len(set(pairs)) == s * (s + 1) // 2
        assert all(p + q <= s + 1 for p, q in pairs)
