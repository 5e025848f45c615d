"""The synthetic video sequence behind the C5 driver (tools/video_c5.py): window layout, lifetimes, repeatable tracks."""
import numpy as np

from vggsfm_b200.synthetic import make_video_scene


def test_window_layout_and_lifetimes():
    sc = make_video_scene(F=200, new_per_window=32, seed=3)
    assert sc.window_range(0) == (0, 32) and sc.window_range(1) == (32, 48)
    assert sc.window_range(sc.num_windows() - 1)[1] == 200
    assert (sc.birth == 0).sum() == 64 and (sc.birth == 2).sum() == 32
    ids = np.nonzero(sc.birth == 2)[0]
    s, _ = sc.window_range(2)
    assert (sc.first_frame[ids] == s).all() and (sc.last_frame[ids] == s + 3 * sc.window).all()
    uv, ok = sc.observe(ids, 0, 200)
    assert not ok[:s].any() and not ok[s + 3 * sc.window:].any()
    assert ok[s:s + 3 * sc.window].mean() > 0.6                       # most points stay in view for their three windows


def test_observations_are_a_pure_function_of_frame_and_point():
    sc = make_video_scene(F=120, new_per_window=16, seed=1)
    ids = np.nonzero(sc.birth == 1)[0]
    a, oa = sc.observe(ids, 30, 80)
    b, ob = sc.observe(ids[3:9], 40, 50)
    assert np.array_equal(a[10:20, 3:9], b) and np.array_equal(oa[10:20, 3:9], ob)
    # noise has the requested scale
    clean = make_video_scene(F=120, new_per_window=16, seed=1, noise_px=0.0)
    c, _ = clean.observe(ids, 30, 80)
    d = (a - c)[oa]
    assert 0.2 < d.std() < 0.4 and abs(d.mean()) < 0.05
