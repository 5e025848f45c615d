"""SceneStore (vggsfm_b200/video.py): the tensor tables behind the video runner's point_dict / frame_dict
(vggsfm/runners/video_runner.py:354-492, :606-638) against a plain-dict restatement of the same bookkeeping.  CPU."""
import numpy as np
import torch

from vggsfm_b200.video import SceneStore


def _dict_add(point_dict, frame_dict, start, valid, track, vis, ids, xyz):
    for k, pid in enumerate(ids):
        tr = point_dict[pid]["track"] if pid in point_dict else {}
        for s in range(valid.shape[0]):
            if valid[s, k]:
                tr[start + s] = (track[s, k].copy(), float(vis[s, k]))
                frame_dict.setdefault(start + s, []).append(pid)
        if pid not in point_dict:
            point_dict[pid] = {"xyz": xyz[k].copy(), "track": tr}


def test_tables_match_dict_bookkeeping():
    rng = np.random.default_rng(0)
    st = SceneStore(torch.device("cpu"))
    pd, fd = {}, {}
    t = torch.from_numpy
    # window 1: frames 0..4, 6 new points
    v1 = rng.uniform(size=(5, 6)) < 0.7
    tr1, vi1, x1 = rng.uniform(0, 100, (5, 6, 2)).astype(np.float32), rng.uniform(size=(5, 6)).astype(np.float32), rng.normal(size=(6, 3)).astype(np.float32)
    ids1 = st.add_points(t(x1), None, t(tr1), t(vi1), t(v1), 0)
    _dict_add(pd, fd, 0, v1, tr1, vi1, list(range(6)), x1)
    # window 2: frames 4..8: points 1,3,4 carried on, 3 new points
    v2 = rng.uniform(size=(4, 3)) < 0.8
    tr2, vi2 = rng.uniform(0, 100, (4, 3, 2)).astype(np.float32), rng.uniform(size=(4, 3)).astype(np.float32)
    st.extend_tracks(torch.tensor([1, 3, 4]), t(tr2), t(vi2), t(v2), 5)
    _dict_add(pd, fd, 5, v2, tr2, vi2, [1, 3, 4], None)
    v3 = rng.uniform(size=(4, 3)) < 0.9
    tr3, vi3, x3 = rng.uniform(0, 100, (4, 3, 2)).astype(np.float32), rng.uniform(size=(4, 3)).astype(np.float32), rng.normal(size=(3, 3)).astype(np.float32)
    ids3 = st.add_points(t(x3), None, t(tr3), t(vi3), t(v3), 5)
    _dict_add(pd, fd, 5, v3, tr3, vi3, [6, 7, 8], x3)
    assert ids1.tolist() == list(range(6)) and ids3.tolist() == [6, 7, 8] and st.num_points == len(pd)
    for f in range(9):
        assert st.visible_points(f).tolist() == sorted(fd.get(f, []))
    xyz, tracks, masks, _ = st.dense(2, 9)
    for pid, p in pd.items():
        assert np.allclose(xyz[pid].numpy(), p["xyz"])
        for f in range(2, 9):
            assert bool(masks[f - 2, pid]) == (f in p["track"])
            if f in p["track"]:
                assert np.array_equal(tracks[f - 2, pid].numpy(), p["track"][f][0])
    # joint-BA style rebuild: points 2 and 7 deleted, one observation filtered
    keep = torch.ones(9, dtype=torch.bool)
    keep[[2, 7]] = False
    xyz_f, tracks_f, masks_f, _ = st.dense(0, 9)
    masks_f[3, 1] = False
    st.replace_from_ba(0, xyz_f * 2.0, torch.zeros(9, 3, 4, dtype=torch.float64), tracks_f, masks_f, keep)
    assert st.num_points == 7
    old = [i for i in range(9) if i not in (2, 7)]
    for new_id, pid in enumerate(old):                       # renumbered in id order (:617-638)
        assert np.allclose(st.xyz[new_id].numpy(), 2.0 * pd[pid]["xyz"])
        frames = sorted(f for f in pd[pid]["track"] if not (pid == 1 and f == 3))
        got = torch.sort(st.obs_frame[st.obs_point == new_id]).values.tolist()
        assert got == frames
    assert bool((st.obs_vis == 1).all()) and bool(st.has_extri[:9].all())
