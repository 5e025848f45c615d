"""GPU checks of the sliding-window mirrors (vggsfm_b200/video.py; vggsfm/runners/video_runner.py:800-853,
:907-939, :941-1017): window BA with a fixed first frame, constant carried points and constant intrinsics
against the oracle's LM with the same constant sets; align_next_window against the oracle's pose loop."""
import numpy as np
import pytest

from oracle import ba_oracle as bo
from oracle import pose_oracle as po
from tests.helpers import to_dev
from vggsfm_b200.synthetic import make_scene, perturb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cam", ["SIMPLE_PINHOLE", "SIMPLE_RADIAL"])
def test_window_bundle_adjustment(cuda_dev, cam):
    import torch
    from vggsfm_b200 import video
    S, P, E = 9, 384, 150
    sc = make_scene(S, P, cam, seed=21, invisible_frac=0.15)
    extr0, _, _, pts0 = perturb(sc, rot_deg=0.3, trans_frac=0.01, focal_frac=0.0, point_sigma=0.03, seed=22)
    extr0[0] = sc.extrinsics[0]
    pts0[:E] = sc.points3d[:E]
    K1 = sc.intrinsics[:1]
    ex1 = sc.extra_params[:1] if sc.extra_params is not None else None
    mask = sc.mask.copy()
    mask[:, 200] = False
    mask[0, 200] = True                                   # a track with < 2 inliers: not in the problem
    dev = cuda_dev
    out, extr, summ, ok = video.window_bundle_adjustment(
        to_dev(pts0, dev), to_dev(extr0, dev), to_dev(K1, dev), to_dev(ex1, dev) if ex1 is not None else None,
        to_dev(sc.tracks, dev), to_dev(mask, dev), E, shared_camera=True, camera_type=cam)
    assert ok and summ.final_cost < summ.initial_cost
    out, extr = out.cpu().numpy(), extr.cpu().numpy()
    assert np.array_equal(extr[0], extr0[0])              # fixed frame
    assert np.array_equal(out[:E], pts0[:E])              # carried points are constant
    assert np.array_equal(out[200], pts0[200])
    # oracle with the same constant sets
    model = bo.SIMPLE_RADIAL if cam == "SIMPLE_RADIAL" else bo.SIMPLE_PINHOLE
    vidx = np.nonzero(mask.sum(0) >= 2)[0]
    intr = np.zeros((S, 4))
    intr[:, 0], intr[:, 1], intr[:, 2] = K1[0, 0, 0], K1[0, 0, 2], K1[0, 1, 2]
    if ex1 is not None:
        intr[:, 3] = ex1[0, 0]
    const_pose = np.zeros(S, bool)
    const_pose[0] = True
    pc = bo.default_param_const(S, model, bo.INTR_CONST, gauge=False, const_pose=const_pose)
    m = mask[:, vidx]
    point_const = (vidx < E) | ~m.any(0)
    pe, _, xe, sm = bo.lm_solve(extr0.copy(), intr, pts0[vidx].copy(), sc.tracks[:, vidx].astype(np.float64), m, model,
                                bo.INTR_CONST, param_const=pc, point_const=point_const, options=bo.LMOptions())
    assert summ.iterations == sm["iterations"]
    assert np.allclose(extr, pe, atol=1e-7), np.abs(extr - pe).max()
    assert np.allclose(out[vidx], xe, atol=1e-6), np.abs(out[vidx] - xe).max()
    # the new points moved towards the ground truth
    assert np.linalg.norm(out[E:] - sc.points3d[E:], axis=1).mean() < 0.5 * np.linalg.norm(pts0[E:] - sc.points3d[E:], axis=1).mean()


def test_align_next_window_and_filter(cuda_dev):
    import torch
    from vggsfm_b200 import video
    S, P = 9, 300
    sc = make_scene(S, P, "SIMPLE_RADIAL", seed=31, invisible_frac=0.1)
    extr0, _, _, _ = perturb(sc, rot_deg=0.5, trans_frac=0.02, focal_frac=0.0, seed=32)
    inl = sc.mask.copy()
    inl[4, 30:] = False                                   # <= 50 inliers: all points are used (:973-977)
    dev = cuda_dev
    K1, ex1 = to_dev(sc.intrinsics[:1], dev), to_dev(sc.extra_params[:1], dev)
    got = video.align_next_window(to_dev(extr0, dev), to_dev(sc.tracks, dev), to_dev(inl, dev), to_dev(sc.points3d, dev),
                                  K1, ex1, "SIMPLE_RADIAL").cpu().numpy()
    intr = np.tile(np.array([1000.0, 512, 512, sc.extra_params[0, 0]]), (S, 1))
    inl_o = inl.copy()
    inl_o[4] = True
    active = np.ones(S, bool)
    active[0] = False
    pe = extr0.copy()
    for s in range(1, S):
        pe[s], _, _ = po.pose_refinement(extr0[s], intr[s], sc.points3d, sc.tracks[s].astype(np.float64), inl_o[s], 1, False, False)
    assert np.array_equal(got[0], extr0[0])
    assert np.allclose(got, pe, atol=1e-8), np.abs(got - pe).max()
    assert np.abs(got[1:, :, 3] - sc.extrinsics[1:, :, 3]).max() < 0.01      # frame 0 is the (perturbed) anchor
    # filter_points_and_compute_masks: shapes and the >= 3 inlier rule
    pts, trk, msk, valid = video.filter_points_and_compute_masks(to_dev(sc.points3d, dev), to_dev(sc.tracks, dev),
                                                                 to_dev(got, dev), K1, ex1)
    assert pts.shape[0] == int(valid.sum()) == trk.shape[1] == msk.shape[1]
    assert bool((msk.sum(0) >= 3).all()) and int(valid.sum()) > 0.9 * P


def test_joint_BA(cuda_dev):
    """video.joint_BA (video_runner.py:494-541): shared camera refined, outlier observations cleared by the 2 px filter,
    scene normalised (camera-centre extent 5), reprojection RMS at the noise floor."""
    import torch
    from vggsfm_b200 import triangulation as tri
    from vggsfm_b200 import video
    S, P = 20, 600
    sc = make_scene(S, P, "SIMPLE_RADIAL", seed=41, invisible_frac=0.15)
    # what reaches joint_BA has already passed the 4 px window filter: plant mild outliers (5-8 px) on 1 % of the observations
    rng = np.random.default_rng(43)
    planted = (rng.uniform(size=(S, P)) < 0.01) & sc.mask
    ang = rng.uniform(0, 2 * np.pi, size=int(planted.sum()))
    sc.tracks[planted] += (rng.uniform(5, 8, size=ang.shape)[:, None] * np.stack([np.cos(ang), np.sin(ang)], -1)).astype(np.float32)
    extr0, K0, ex0, pts0 = perturb(sc, rot_deg=0.3, trans_frac=0.01, focal_frac=0.02, point_sigma=0.02, seed=42)
    dev = cuda_dev
    pts, E, K, ex, masks, valid = video.joint_BA(to_dev(pts0, dev), to_dev(extr0, dev), to_dev(K0[:1], dev), to_dev(ex0[:1], dev),
                                                 to_dev(sc.tracks, dev), to_dev(sc.mask, dev), camera_type="SIMPLE_RADIAL")
    assert K.shape == (1, 3, 3) and ex.shape == (1, 1) and masks.shape == (S, P) and valid.shape == (P,)
    assert abs(K[0, 0, 0].item() - 1000.0) < 3.0 and abs(ex[0, 0].item() - 0.05) < 0.01
    assert int(valid.sum()) > 0.97 * P
    assert not bool(masks[:, ~valid].any()) and not bool((masks & ~to_dev(sc.mask, dev)).any())
    uvh = tri.project_3D_points(pts, E, K.expand(S, -1, -1), ex.expand(S, -1))
    err = ((uvh - to_dev(sc.tracks, dev).double()) ** 2).sum(-1)
    assert torch.sqrt(err[masks].mean()).item() < 0.5 and err[masks].max().item() <= 4.0 + 1e-9      # 2 px filter
    dropped = (to_dev(sc.mask, dev) & ~masks & valid[None]).cpu().numpy()
    assert (dropped & planted).sum() > 0.95 * (planted & valid.cpu().numpy()[None]).sum()          # the planted outliers go
    assert (dropped & ~planted).sum() < 0.002 * S * P
    En = E.cpu().numpy()
    C = -np.einsum("sji,sj->si", En[:, :, :3], En[:, :, 3])
    Cs = np.sort(C.astype(np.float32), axis=0)
    ext = np.linalg.norm(Cs[int(0.9 * (S - 1))] - Cs[int(0.1 * (S - 1))])
    assert abs(ext - 5.0) < 1e-3, ext


def test_pose_refinement_degenerate_inputs(cuda_dev):
    """No correspondences / no frames: nothing is touched, every frame reports FEW_INLIERS."""
    import torch
    from vggsfm_b200 import pose_refinement as pr
    poses = torch.eye(3, 4, dtype=torch.float64, device=cuda_dev).repeat(3, 1, 1).contiguous()
    intr = torch.tensor([[500.0, 320, 240, 0]] * 3, dtype=torch.float64, device=cuda_dev)
    rep = pr.pose_refinement_batched(poses, intr, torch.zeros(0, 3, dtype=torch.float64, device=cuda_dev),
                                     torch.zeros(3, 0, 2, device=cuda_dev), torch.zeros(3, 0, dtype=torch.bool, device=cuda_dev),
                                     torch.full((3,), 7, dtype=torch.uint8, device=cuda_dev), 0)
    assert rep.termination.tolist() == [7, 7, 7] and rep.inlier_used.shape == (3, 0)
    assert torch.equal(poses[0], torch.eye(3, 4, dtype=torch.float64, device=cuda_dev))
    # all correspondences masked out: the kernel runs and leaves the frame alone
    X = torch.randn(50, 3, dtype=torch.float64, device=cuda_dev) + torch.tensor([0, 0, 5.0], dtype=torch.float64, device=cuda_dev)
    rep = pr.pose_refinement_batched(poses, intr, X, torch.zeros(3, 50, 2, device=cuda_dev),
                                     torch.zeros(3, 50, dtype=torch.bool, device=cuda_dev),
                                     torch.full((3,), 7, dtype=torch.uint8, device=cuda_dev), 0)
    assert rep.termination.tolist() == [7, 7, 7] and int(rep.num_inliers.sum()) == 0
