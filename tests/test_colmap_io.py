"""COLMAP binary export of vggsfm_b200.reconstruction.Reconstruction (vggsfm_b200/colmap_io.py): ids and ordering follow
batch_matrix_to_pycolmap (vggsfm/utils/tensor_to_pycolmap.py:16-160), byte layout follows COLMAP's model format;
checked by an independent byte-level parse and a write -> read round trip.  CPU only."""
import struct

import numpy as np
import pytest
import torch

from vggsfm_b200 import colmap_io as cio
from vggsfm_b200.reconstruction import Reconstruction
from vggsfm_b200.synthetic import make_scene


def _rec(cam, shared, S=5, P=40):
    sc = make_scene(S, P, cam, seed=3, invisible_frac=0.3)
    masks = sc.mask.copy()
    masks[:, 7] = False                       # a track with no observations
    masks[1:, 9] = False                      # a track with a single observation: not a COLMAP point
    xyz = sc.points3d.copy()
    xyz[11] = 0.0                             # a point the BA deleted (reads back as zeros)
    alive = np.ones(P, dtype=bool)
    alive[11] = False
    t = torch.from_numpy
    rec = Reconstruction.from_batch_matrix(t(xyz), t(sc.extrinsics), t(sc.intrinsics), t(sc.tracks), t(masks),
                                           torch.tensor([1024, 768]), shared_camera=shared, camera_type=cam,
                                           extra_params=t(sc.extra_params) if sc.extra_params is not None else None,
                                           alive=t(alive))
    return rec, sc, masks, xyz


@pytest.mark.parametrize("cam,shared", [("SIMPLE_PINHOLE", False), ("SIMPLE_RADIAL", True)])
def test_round_trip_and_ids(tmp_path, cam, shared):
    rec, sc, masks, xyz = _rec(cam, shared)
    rec.set_point_colors(torch.linspace(0, 1, 40)[:, None].repeat(1, 3))       # ids 1..max id
    rec.write(str(tmp_path))
    m = cio.read_model(str(tmp_path))
    S, P = masks.shape
    keep = (masks.sum(0) >= 2) & (np.abs(xyz).sum(1) > 0)
    assert not keep[7] and not keep[9] and not keep[11]
    order = np.nonzero(masks.sum(0) >= 2)[0]                                     # ids 1..P' in track order (:62-70)
    dead = int(np.nonzero(order == 11)[0][0]) + 1                               # the deleted point keeps its id, unused
    assert sorted(m["points3D"]) == [i for i in range(1, len(order) + 1) if i != dead]
    for pid, p in m["points3D"].items():
        assert np.array_equal(p["xyz"], xyz[order[pid - 1]])
        assert p["rgb"][0] == int(round((pid - 1) / 39 * 255)) and p["error"] == -1.0
        assert len(p["track"]) == int(masks[:, order[pid - 1]].sum())
        for (iid, idx2d) in p["track"]:                                          # element -> that image's point2D -> back
            assert m["images"][iid]["point3D_ids"][idx2d] == pid
            assert np.array_equal(m["images"][iid]["xys"][idx2d], sc.tracks[iid, order[pid - 1]].astype(np.float64))
    assert sorted(m["images"]) == list(range(S))
    assert sorted(m["cameras"]) == ([0] if shared else list(range(S)))
    for s in range(S):
        im = m["images"][s]
        assert im["name"] == f"image_{s}" and im["camera_id"] == (0 if shared else s)
        assert np.allclose(cio.qvec_to_rotmat(im["qvec"]), sc.extrinsics[s, :, :3], atol=1e-14)
        assert np.array_equal(im["tvec"], sc.extrinsics[s, :, 3])
        assert len(im["point3D_ids"]) == int((masks[s] & keep).sum())
        assert np.all(np.diff(im["point3D_ids"]) > 0)                            # point order
    c = m["cameras"][0]
    assert (c["width"], c["height"]) == (1024, 768)
    if cam == "SIMPLE_RADIAL":
        assert c["model_id"] == 2 and np.array_equal(c["params"], [1000.0, 512.0, 512.0, 0.05])
    else:
        assert c["model_id"] == 0 and np.array_equal(c["params"], [1000.0, 512.0, 512.0])


def test_byte_layout(tmp_path):
    """Independent parse of the first records with explicit struct formats (COLMAP read_write_model.py layout)."""
    rec, sc, masks, xyz = _rec("SIMPLE_RADIAL", False)
    rec.write(str(tmp_path))
    b = (tmp_path / "cameras.bin").read_bytes()
    assert struct.unpack_from("<Q", b, 0)[0] == 5
    cid, mid, w, h = struct.unpack_from("<iiQQ", b, 8)
    assert (cid, mid, w, h) == (0, 2, 1024, 768)
    assert struct.unpack_from("<4d", b, 32) == (1000.0, 512.0, 512.0, 0.05)
    assert len(b) == 8 + 5 * (24 + 32)
    b = (tmp_path / "images.bin").read_bytes()
    assert struct.unpack_from("<Q", b, 0)[0] == 5
    iid = struct.unpack_from("<i", b, 8)[0]
    q = struct.unpack_from("<4d", b, 12)
    assert iid == 0 and abs(np.linalg.norm(q) - 1) < 1e-15 and q[0] > 0
    assert b[72:80] == b"image_0\x00"
    b = (tmp_path / "points3D.bin").read_bytes()
    n = struct.unpack_from("<Q", b, 0)[0]
    pid, x, y, z, r, g, bl, err = struct.unpack_from("<QdddBBBd", b, 8)
    tl = struct.unpack_from("<Q", b, 8 + 43)[0]
    assert pid == 1 and (x, y, z) == tuple(xyz[0]) and (r, g, bl) == (0, 0, 0) and err == -1.0
    assert tl == int(masks[:, 0].sum()) and n == len(cio.read_model(str(tmp_path))["points3D"])


def test_quaternion_branches():
    """All four branches of the matrix -> quaternion conversion give the rotation back."""
    from oracle.ba_oracle import exp_so3
    for w in ([0.1, 0.2, -0.1], [3.0, 0.1, 0.1], [0.1, 3.0, 0.1], [0.1, 0.1, 3.0], [2.2, 2.2, 0.0]):
        R = exp_so3(np.array(w))
        q = cio.rotmat_to_qvec(R)
        assert abs(np.linalg.norm(q) - 1) < 1e-15 and np.allclose(cio.qvec_to_rotmat(q), R, atol=1e-14)


def test_unsupported_camera_type(tmp_path):
    sc = make_scene(3, 8, "SIMPLE_PINHOLE", seed=3)
    t = torch.from_numpy
    with pytest.raises(ValueError, match="is not supported yet"):      # tensor_to_pycolmap.py:97-100
        Reconstruction.from_batch_matrix(t(sc.points3d), t(sc.extrinsics), t(sc.intrinsics), t(sc.tracks), t(sc.mask),
                                         torch.tensor([1024, 768]), camera_type="OPENCV")
