"""GPU end-to-end slice of Triangulator.triangulate_tracks_and_BA + iterative_global_BA
(models/triangulator.py:365-439, utils/triangulation.py:1076-1209) built from the CUDA pieces:
cam_from_img -> triangulate_tracks -> global_BA -> filter_all_points3D -> iterative_global_BA.
Checked by size-independent properties (reprojection RMS at the noise floor, geometry up to a similarity)."""
import numpy as np
import pytest

from tests.helpers import to_dev

pytestmark = pytest.mark.gpu


def _umeyama(A, B):
    """similarity (s,R,t) minimising |s R A + t - B|."""
    ma, mb = A.mean(0), B.mean(0)
    Ac, Bc = A - ma, B - mb
    U, D, Vt = np.linalg.svd(Bc.T @ Ac / len(A))
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    s = np.trace(np.diag(D) @ S) / (Ac ** 2).sum() * len(A)
    return s, R, mb - s * R @ ma


@pytest.mark.parametrize("cam,shared", [("SIMPLE_PINHOLE", False), ("SIMPLE_RADIAL", True)])
def test_triangulate_then_ba(cuda_dev, cam, shared):
    import torch
    from vggsfm_b200 import bundle_adjustment as ba
    from vggsfm_b200 import triangulation as tri
    from vggsfm_b200.synthetic import make_scene, perturb
    S, N = 16, 512
    sc = make_scene(S, N, cam, seed=4, invisible_frac=0.2, outlier_frac=0.03)
    extr0, K0, extra0, _ = perturb(sc, rot_deg=0.5, trans_frac=0.01, focal_frac=0.03, seed=5)
    dev = cuda_dev
    E, K = to_dev(extr0, dev), to_dev(K0, dev)
    ex = to_dev(extra0, dev) if extra0 is not None else None
    tracks, vis, score = to_dev(sc.tracks, dev), to_dev(sc.vis, dev), to_dev(sc.score, dev)
    image_size = torch.tensor([1024, 1024], device=dev)
    torch.manual_seed(0)
    tn = tri.cam_from_img(tracks, K, ex)
    pts, num, mask = tri.triangulate_tracks(E, tn, track_vis=vis, track_score=score)
    valid = num >= 3                                                       # triangulator.py:399
    p1, E1, K1, ex1, rec = ba.global_BA(pts, valid, tracks, mask, E, K, ex, image_size, shared_camera=shared,
                                        camera_type=cam)
    assert rec.summary.final_cost < rec.summary.initial_cost
    ok, _ = tri.filter_all_points3D(p1, tracks[:, valid], E1, K1, extra_params=ex1, max_reproj_error=4,
                                    check_triangle=False)
    valid2 = valid.clone()
    valid2[valid] = ok
    p2, E2, K2, ex2, valid3, masks3, rec2 = ba.iterative_global_BA(
        tracks, K1, E1, vis, score, valid2, p1[ok], image_size, shared_camera=shared, min_valid_track_length=3,
        max_reproj_error=2, lastBA=True, camera_type=cam, extra_params=ex1)
    assert p2.shape[0] == int(valid3.sum()) and masks3.shape == (S, p2.shape[0])
    assert int(valid3.sum()) > 0.9 * N
    # reprojection RMS over the surviving observations ~ the 0.3 px noise
    uvh = tri.project_3D_points(p2, E2, K2, ex2)
    err = ((uvh - tracks[:, valid3].double()) ** 2).sum(-1)
    rms = torch.sqrt(err[masks3].mean()).item()
    assert rms < 0.6, rms
    # geometry recovered up to a similarity
    gt = sc.points3d[valid3.cpu().numpy()]
    s, R, t = _umeyama(p2.cpu().numpy(), gt)
    res = np.linalg.norm((s * (R @ p2.cpu().numpy().T).T + t) - gt, axis=1)
    assert np.median(res) < 5e-3
    # focal close to the ground truth 1000 px
    assert abs(K2[:, 0, 0].mean().item() - 1000.0) < 10.0
    assert rec2 is not None and ba.get_valid_frame_mask(K2, E2, ex2, 1024).all()
