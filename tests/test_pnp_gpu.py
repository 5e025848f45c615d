"""GPU parity of the batched P3P LO-RANSAC absolute-pose kernel (csrc/pnp.cu, vgg_absolute_pose_estimation) against
oracle/pnp_oracle.py on the same host-drawn samples, and the refine_pose fall-back it serves
(vggsfm/utils/triangulation.py:404-433).  Bars: inlier counts / masks / chosen focal factor exact (integer and
selection work), poses 1e-8 (float64, different summation order in the Gauss-Newton reductions)."""
import numpy as np
import pytest

from oracle import pnp_oracle as po
from tests.helpers import rotation_angle_deg, to_dev
from vggsfm_b200.synthetic import make_scene, perturb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cam,est_f,P", [("SIMPLE_PINHOLE", False, 300), ("SIMPLE_RADIAL", False, 257), ("SIMPLE_PINHOLE", True, 200)])
def test_kernel_matches_oracle(cuda_dev, cam, est_f, P):
    import torch
    from vggsfm_b200 import pose_refinement as pr
    S = 5
    sc = make_scene(S, P, cam, seed=7, noise_px=0.3, outlier_frac=0.25, invisible_frac=0.2)
    model = po.SIMPLE_RADIAL if cam == "SIMPLE_RADIAL" else po.SIMPLE_PINHOLE
    k = 0.05 if cam == "SIMPLE_RADIAL" else 0.0
    f0 = 1000.0 * (1.5 if est_f else 1.0)
    intr4 = np.tile(np.array([f0, 512.0, 512.0, k]), (S, 1))
    intr4[3, 0] *= 1.1 if est_f else 1.0
    us = np.random.default_rng(3).uniform(size=(24, 3))
    frames = np.array([True, True, False, True, True])
    dev = cuda_dev
    poses, focal, ninl, inl = pr.absolute_pose_estimation_batched(
        to_dev(sc.tracks, dev), to_dev(sc.points3d, dev), to_dev(sc.mask, dev), to_dev(intr4, dev), model,
        frames=to_dev(frames, dev), estimate_focal_length=est_f, max_error=12.0, u_samples=torch.from_numpy(us))
    poses, focal, ninl, inl = poses.cpu().numpy(), focal.cpu().numpy(), ninl.cpu().numpy(), inl.cpu().numpy()
    assert ninl[2] == 0 and not inl[2].any() and not poses[2].any()
    for s in (0, 1, 3, 4):
        r = po.absolute_pose_estimation(sc.tracks[s], sc.points3d, intr4[s], model, us, estimate_focal_length=est_f,
                                        max_error=12.0, mask=sc.mask[s])
        assert r is not None and ninl[s] == r["num_inliers"], (s, ninl[s], r["num_inliers"])
        assert focal[s] == r["focal"]
        assert np.array_equal(inl[s], r["inliers"])
        assert np.abs(poses[s] - r["pose"]).max() < 1e-8, np.abs(poses[s] - r["pose"]).max()
        E = sc.extrinsics[s]
        if not est_f:
            assert rotation_angle_deg(poses[s][None, :, :3], E[None, :, :3]).max() < 0.1


def test_too_few_points_gives_no_model(cuda_dev):
    from vggsfm_b200 import pose_refinement as pr
    sc = make_scene(2, 40, "SIMPLE_PINHOLE", seed=1)
    mask = np.zeros((2, 40), dtype=bool)
    mask[0, :2] = True                           # two usable points: no minimal sample
    intr4 = np.tile(np.array([1000.0, 512.0, 512.0, 0.0]), (2, 1))
    dev = cuda_dev
    poses, focal, ninl, inl = pr.absolute_pose_estimation_batched(to_dev(sc.tracks, dev), to_dev(sc.points3d, dev),
                                                                  to_dev(mask, dev), to_dev(intr4, dev), 0)
    assert ninl.cpu().tolist() == [0, 0] and not bool(inl.any())


@pytest.mark.parametrize("cam,shared", [("SIMPLE_PINHOLE", False), ("SIMPLE_RADIAL", True)])
def test_refine_pose_force_estimate_recovers_lost_frames(cuda_dev, cam, shared):
    """refine_pose(force_estimate=True): frames whose pose is so wrong that fewer than 100 observations reproject
    within 12 px take the absolute-pose fall-back (triangulation.py:404-433) and come back at the ground truth."""
    import torch
    from vggsfm_b200 import pose_refinement as pr
    S, N = 8, 600
    sc = make_scene(S, N, cam, seed=9, noise_px=0.3, outlier_frac=0.05)
    extr0, K0, ex0, _ = perturb(sc, rot_deg=0.2, trans_frac=0.005, focal_frac=0.0, seed=10)
    lost = [2, 5]
    for s in lost:                                    # 25 degrees off + a shifted centre
        a = np.deg2rad(25.0)
        Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        extr0[s, :, :3] = Ry @ extr0[s, :, :3]
        extr0[s, :, 3] += np.array([0.8, -0.3, 0.5])
    dev = cuda_dev
    E, K = to_dev(extr0, dev), to_dev(sc.intrinsics, dev)
    ex = to_dev(sc.extra_params, dev) if sc.extra_params is not None else None
    tracks, vis = to_dev(sc.tracks, dev), to_dev(sc.vis, dev)
    valid = torch.ones(N, dtype=torch.bool, device=dev)
    isz = torch.tensor([1024, 1024], device=dev)
    torch.manual_seed(0)
    args = (E, K, ex, vis > 0.05, to_dev(sc.points3d, dev), tracks, valid, isz)
    E0, _, _, _ = pr.refine_pose(*args, shared_camera=shared, camera_type=cam, force_estimate=False)
    assert sorted(torch.nonzero(pr.last_report.needs_absolute_pose).flatten().tolist()) == lost
    err0 = rotation_angle_deg(E0.cpu().numpy()[:, :, :3], sc.extrinsics[:, :, :3])
    assert err0[lost].min() > 10.0                    # without the fall-back the lost frames stay lost
    E1, K1, ex1, vmask = pr.refine_pose(*args, shared_camera=shared, camera_type=cam, force_estimate=True)
    assert pr.last_report.absolute_pose_ok[lost].all()
    err1 = rotation_angle_deg(E1.cpu().numpy()[:, :, :3], sc.extrinsics[:, :, :3])
    assert err1.max() < 0.2, err1
    assert np.abs(E1.cpu().numpy()[:, :, 3] - sc.extrinsics[:, :, 3]).max() < 0.02
    assert bool(vmask.all()) and abs(K1[lost[0], 0, 0].item() - 1000.0) < 30.0
