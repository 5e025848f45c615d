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
    assert ba.get_valid_frame_mask(K2, E2, ex2, 1024).all()
    # the lastBA reconstruction (triangulation.py:1186-1199): P points x S images, normalised once more, serialisable
    import tempfile
    from vggsfm_b200 import colmap_io as cio
    P = p2.shape[0]
    back = rec2.to_batch_matrix(device="cpu", camera_type=cam)
    assert back[0].shape == (P, 3) and back[1].shape == (S, 3, 4) and back[2].shape == (S, 3, 3)
    e_n, p_n = ba.normalize(E2, p2, 5.0, 0.1, 0.9)
    assert (back[0] - p_n.cpu()).abs().max().item() < 1e-12 and (back[1] - e_n.cpu()).abs().max().item() < 1e-12
    assert (back[2] - K2.cpu()).abs().max().item() == 0.0
    with tempfile.TemporaryDirectory() as d:
        rec2.write(d)
        m = cio.read_model(d)
    assert len(m["points3D"]) == P and sorted(m["images"]) == list(range(S))
    assert len(m["cameras"]) == (1 if shared else S)
    xyz = np.stack([m["points3D"][i + 1]["xyz"] for i in range(P)])
    assert np.array_equal(xyz, back[0].numpy())
    obs = masks3.cpu().numpy()
    for s_ in (0, S - 1):
        assert len(m["images"][s_]["point3D_ids"]) == int(obs[s_].sum())
        assert np.allclose(cio.qvec_to_rotmat(m["images"][s_]["qvec"]), back[1][s_, :, :3].numpy(), atol=1e-12)
    # the intermediate (lastBA=False) call also hands back the BA'd reconstruction, like the reference (:1146)
    assert rec.num_points3D() > 0 and rec.num_images() == S


class _Cameras:
    """The three attributes get_EFP reads from the camera predictor's output (models/utils.py:46-55)."""

    def __init__(self, focal_ndc, R, T):
        self.focal_length, self.R, self.T = focal_ndc, R, T


@pytest.mark.parametrize("cam,shared", [("SIMPLE_PINHOLE", False), ("SIMPLE_RADIAL", True)])
def test_triangulator_forward(cuda_dev, cam, shared):
    """Triangulator.forward (models/triangulator.py:44-363) end to end on the CUDA path: same arguments, same
    9-tuple; checked by size-independent properties (reprojection RMS at the noise floor, camera centres equal to
    the ground truth up to a similarity)."""
    import torch
    from vggsfm_b200 import triangulation as tri
    from vggsfm_b200.synthetic import make_scene, perturb
    from vggsfm_b200.triangulator import Triangulator
    S, N = 12, 1024
    sc = make_scene(S, N, cam, seed=11, invisible_frac=0.1, outlier_frac=0.02, k=0.03)
    extr0, K0, _, _ = perturb(sc, rot_deg=0.4, trans_frac=0.01, focal_frac=0.02, seed=12)
    dev = cuda_dev
    W = H = 1024
    cams = _Cameras(to_dev(np.stack([K0[:, 0, 0], K0[:, 1, 1]], -1) * 2.0 / min(W, H), dev, torch.float32),
                    to_dev(extr0[:, :, :3], dev, torch.float32), to_dev(extr0[:, :, 3], dev, torch.float32))
    tracks = to_dev(sc.tracks, dev)[None]
    vis = to_dev(sc.vis, dev)[None]
    score = to_dev(sc.score, dev)[None]
    images = torch.rand(1, S, 3, H, W, device=dev)
    # the two-view stage upstream hands over epipolar inliers: here, pairs whose two observations are not outliers
    from vggsfm_b200.synthetic import project_np
    uv_gt, _ = project_np(sc.extrinsics, 1000.0, np.array([512.0, 512.0]), 0.03 if cam == "SIMPLE_RADIAL" else 0.0, sc.points3d)
    ok = np.linalg.norm(sc.tracks - uv_gt, axis=-1) < 3.0
    prelim = {"fmat_inlier_mask": to_dev(ok[:1] & ok[1:], dev)[None]}
    torch.manual_seed(0)
    out = Triangulator()(cams, tracks, vis, images, prelim, pred_score=score, BA_iters=2, shared_camera=shared,
                         robust_refine=2, camera_type=cam)
    E, K, ex, pts, rgb, rec, vframe, v2d, vtracks = out
    P = int(vtracks.sum())
    assert E.shape == (S, 3, 4) and E.dtype == torch.float64 and K.shape == (S, 3, 3)
    assert pts.shape == (P, 3) and rgb.shape == (P, 3) and v2d.shape == (S, N) and vframe.shape == (S,)
    assert (ex is None) == (cam == "SIMPLE_PINHOLE")
    assert P > 0.85 * N and bool(vframe.all())
    assert not bool(v2d[:, ~vtracks].any())
    assert bool(((rgb >= 0) & (rgb <= 1)).all())
    uvh = tri.project_3D_points(pts, E, K, ex)
    err = ((uvh - tracks[0][:, vtracks].double()) ** 2).sum(-1)
    rms = torch.sqrt(err[v2d[:, vtracks]].mean()).item()
    assert rms < 0.6, rms
    # camera centres vs ground truth up to a similarity
    En = E.cpu().numpy()
    C_est = -np.einsum("sji,sj->si", En[:, :, :3], En[:, :, 3])
    C_gt = -np.einsum("sji,sj->si", sc.extrinsics[:, :, :3], sc.extrinsics[:, :, 3])
    s, R, t = _umeyama(C_est, C_gt)
    resid = np.linalg.norm((s * (R @ C_est.T).T + t) - C_gt, axis=1).max()
    assert resid < 0.02, resid
    if shared:
        assert torch.all(K[:, 0, 0] == K[0, 0, 0])
        assert abs(K[0, 0, 0].item() - 1000.0) < 5.0
