"""vggsfm_b200.pycolmap_compat: the pycolmap-shaped entry points (bundle_adjustment(reconstruction, options),
pose_refinement, absolute_pose_estimation, ObservationManager) give the same numbers as the tensor mirrors they wrap,
on the scene object built by batch_matrix_to_pycolmap -- i.e. the reference's call sequence
batch_matrix_to_pycolmap -> pycolmap.bundle_adjustment -> filter_reconstruction -> pycolmap_to_batch_matrix
(vggsfm/utils/triangulation.py:1033-1063) statement by statement."""
import numpy as np
import pytest

from tests.helpers import rotation_angle_deg, to_dev
from vggsfm_b200.synthetic import make_scene, perturb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cam,shared", [("SIMPLE_PINHOLE", False), ("SIMPLE_RADIAL", True)])
def test_global_ba_call_sequence(cuda_dev, cam, shared):
    import torch
    from vggsfm_b200 import bundle_adjustment as ba
    from vggsfm_b200 import pycolmap_compat as pycolmap
    from vggsfm_b200.reconstruction import batch_matrix_to_pycolmap, pycolmap_to_batch_matrix
    S, N = 7, 150
    sc = make_scene(S, N, cam, seed=21, invisible_frac=0.3)
    extr, K, extra, pts = perturb(sc, seed=22)
    if shared:
        K[:] = K[0]
    t = torch.from_numpy
    size = torch.tensor([1024, 1024])
    # the reference's statements (triangulation.py:1033-1063)
    rec = batch_matrix_to_pycolmap(t(pts), t(extr), t(K), t(sc.tracks), t(sc.mask), size, shared_camera=shared,
                                   camera_type=cam, extra_params=t(extra) if extra is not None else None)
    opt = pycolmap.BundleAdjustmentOptions()
    opt.solver_options.gradient_tolerance *= 10
    opt.solver_options.max_num_iterations = 50
    summ = pycolmap.bundle_adjustment(rec, opt)
    rec.normalize(5.0, 0.1, 0.9, True)                                           # filter_reconstruction
    p_o, e_o, k_o, x_o = pycolmap_to_batch_matrix(rec, device="cpu", camera_type=cam)
    # the tensor mirror of the same sequence
    dev = cuda_dev
    out = ba.bundle_adjustment(to_dev(pts, dev), to_dev(extr, dev), to_dev(K, dev), to_dev(extra, dev) if extra is not None else None,
                               to_dev(sc.tracks, dev), to_dev(sc.mask, dev), shared_camera=shared, camera_type=cam,
                               options=ba.prepare_ba_options())
    assert summ.iterations == out[5].iterations
    assert np.abs(p_o.numpy() - out[0].cpu().numpy()).max() < 1e-9
    assert np.abs(e_o.numpy() - out[1].cpu().numpy()).max() < 1e-9
    assert np.abs(k_o.numpy() - out[2].cpu().numpy()).max() < 1e-8
    if x_o is not None:
        assert np.abs(x_o.numpy() - out[3].cpu().numpy()).max() < 1e-10


def test_pose_calls_and_observation_manager(cuda_dev):
    import torch
    from vggsfm_b200 import pycolmap_compat as pycolmap
    from vggsfm_b200.reconstruction import batch_matrix_to_pycolmap
    S, N = 5, 300
    sc = make_scene(S, N, "SIMPLE_PINHOLE", seed=31, noise_px=0.3, outlier_frac=0.1)
    extr, K, _, _ = perturb(sc, rot_deg=0.3, trans_frac=0.005, focal_frac=0.0, seed=32)
    s = 3
    cam = pycolmap.Camera(model="SIMPLE_PINHOLE", width=1024, height=1024, params=np.array([1000.0, 512.0, 512.0]), camera_id=s)
    ro = pycolmap.AbsolutePoseRefinementOptions()
    ro.refine_focal_length = True
    ans = pycolmap.pose_refinement(pycolmap.Rigid3d(pycolmap.Rotation3d(extr[s][:, :3]), extr[s][:, 3]), sc.tracks[s], sc.points3d,
                                   sc.mask[s], cam, ro)
    E = ans["cam_from_world"].matrix()
    assert rotation_angle_deg(E[None, :, :3], sc.extrinsics[s][None, :, :3]).max() < 0.1
    assert abs(cam.params[0] - 1000.0) < 10.0
    eo = pycolmap.AbsolutePoseEstimationOptions()
    eo.estimate_focal_length = True
    eo.ransac.max_error = 12
    cam2 = pycolmap.Camera(model="SIMPLE_PINHOLE", width=1024, height=1024, params=np.array([1500.0, 512.0, 512.0]), camera_id=s)
    torch.manual_seed(0)
    est = pycolmap.absolute_pose_estimation(sc.tracks[s][sc.mask[s]], sc.points3d[sc.mask[s]], cam2, eo, ro)
    assert est is not None and est["num_inliers"] > 0.7 * int(sc.mask[s].sum())
    assert rotation_angle_deg(est["cam_from_world"].matrix()[None, :, :3], sc.extrinsics[s][None, :, :3]).max() < 0.2
    assert abs(cam2.params[0] - 1000.0) < 30.0
    assert pycolmap.absolute_pose_estimation(sc.tracks[s][:2], sc.points3d[:2], cam2, eo, ro) is None
    # ObservationManager on a ground-truth scene with outliers: the gross outliers go, the good observations stay
    t = torch.from_numpy
    rec = batch_matrix_to_pycolmap(t(sc.points3d), t(sc.extrinsics), t(sc.intrinsics), t(sc.tracks), t(sc.mask),
                                   torch.tensor([1024, 1024]))
    before = sum(p.track.length() for p in rec.points3D.values())
    om = pycolmap.ObservationManager(rec)
    om.filter_all_points3D(2.0, 1.5)
    om.filter_observations_with_negative_depth()
    after = sum(p.track.length() for p in rec.points3D.values())
    assert 0.8 * before < after < 0.97 * before
    for p in rec.points3D.values():
        assert p.track.length() >= 2
