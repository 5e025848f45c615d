"""GPU parity of the batched pose-refinement kernel (csrc/pose_refine.cu, vgg_pose_refinement) against
oracle/pose_oracle.py, and of the refine_pose / init_refine_pose mirrors against the oracle's frame loop.
Floating point (f64): the LM trajectory is identical up to reduction order, so poses agree to 1e-8 relative
and the iteration / termination records agree exactly; effective inlier masks are bit-exact."""
import numpy as np
import pytest

from oracle import pose_oracle as po
from oracle.ba_oracle import exp_so3, project
from tests.helpers import to_dev

pytestmark = pytest.mark.gpu


def _frames(model, S=12, P=700, seed=0, shared=False):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(P, 3)) * 0.8 + np.array([0, 0, 6.0])
    poses = np.stack([np.concatenate([exp_so3(rng.normal(size=3) * 0.15), rng.normal(size=(3, 1)) * 0.3], 1) for _ in range(S)])
    intr = np.tile(np.array([800.0, 512, 384, 0.04 if model == 1 else 0.0]), (S, 1))
    if not shared:
        intr[:, 0] += rng.uniform(-30, 30, size=S)
    uv, _ = project(poses, intr, X, model)
    uv = (uv + rng.normal(size=uv.shape) * 0.4).astype(np.float32)
    out = rng.uniform(size=(S, P)) < 0.05
    uv[out] += rng.normal(size=(int(out.sum()), 2)).astype(np.float32) * 60
    inl = rng.uniform(size=(S, P)) < 0.8
    inl[3, 40:] = False                          # a frame with too few inliers
    p0 = poses.copy()
    for s in range(S):
        p0[s, :, :3] = exp_so3(rng.normal(size=3) * 0.01) @ p0[s, :, :3]
        p0[s, :, 3] += rng.normal(size=3) * 0.03
    i0 = intr.copy()
    i0[:, 0] *= 1.03 if shared else rng.uniform(0.96, 1.04, size=S)
    return X, uv, inl, p0, i0


@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("max_err,min_inl", [(0.0, 0), (12.0, 100)])
def test_kernel_matches_oracle(cuda_dev, model, max_err, min_inl):
    import torch
    from vggsfm_b200 import pose_refinement as pr
    X, uv, inl, p0, i0 = _frames(model, seed=model)
    S, P = inl.shape
    active = np.ones(S, bool)
    active[5] = False
    rf = np.ones(S, bool)
    rf[6] = False
    rk = np.ones(S, bool)
    rk[7] = False
    flags = active * 1 + rf * 2 + rk * 4
    poses = to_dev(p0, cuda_dev)
    intr = to_dev(i0, cuda_dev)
    opt = pr.default_pose_options()
    opt.max_reproj_error = max_err
    opt.min_inliers = min_inl
    rep = pr.pose_refinement_batched(poses, intr, to_dev(X, cuda_dev), to_dev(uv, cuda_dev), to_dev(inl, cuda_dev),
                                     to_dev(flags.astype(np.uint8), cuda_dev), model, opt)
    torch.cuda.synchronize()
    got_p, got_i = poses.cpu().numpy(), intr.cpu().numpy()
    used = rep.inlier_used.cpu().numpy()
    term = rep.termination.cpu().numpy()
    its = rep.iterations.cpu().numpy()
    for s in range(S):
        _, _, use_o, _ = po.frame_loop(p0[s:s + 1], i0[s:s + 1], X, uv[s:s + 1].astype(np.float64), inl[s:s + 1], [False], model,
                                       False, max_err, min_inl)
        assert np.array_equal(used[s], use_o[0]), s
        if not active[s]:
            assert term[s] == 6 and np.array_equal(got_p[s], p0[s]) and np.array_equal(got_i[s], i0[s])
            continue
        if use_o[0].sum() <= min_inl:
            assert term[s] == 7 and np.array_equal(got_p[s], p0[s])
            continue
        pe, ie, sm = po.pose_refinement(p0[s], i0[s], X, uv[s].astype(np.float64), use_o[0], model, bool(rf[s]), bool(rk[s]))
        assert term[s] == sm["termination"] and its[s] == sm["iterations"], (s, term[s], sm)
        assert np.allclose(got_p[s], pe, rtol=0, atol=1e-8), (s, np.abs(got_p[s] - pe).max())
        assert np.allclose(got_i[s], ie, rtol=1e-9, atol=1e-9), (s, got_i[s], ie)
        assert abs(rep.final_cost[s].item() - sm["final_cost"]) <= 1e-9 * sm["final_cost"]
        assert abs(rep.initial_cost[s].item() - sm["initial_cost"]) <= 1e-9 * sm["initial_cost"]
        if not rf[s]:
            assert got_i[s, 0] == i0[s, 0]
        if not rk[s] or model == 0:
            assert got_i[s, 3] == i0[s, 3]


def _K(intr):
    S = len(intr)
    K = np.zeros((S, 3, 3))
    K[:, 0, 0] = K[:, 1, 1] = intr[:, 0]
    K[:, 0, 2] = intr[:, 1]
    K[:, 1, 2] = intr[:, 2]
    K[:, 2, 2] = 1
    return K


@pytest.mark.parametrize("cam,shared", [("SIMPLE_PINHOLE", False), ("SIMPLE_RADIAL", True), ("SIMPLE_RADIAL", False)])
def test_refine_pose_mirror(cuda_dev, cam, shared):
    """refine_pose (triangulation.py:260-479) vs the oracle's frame loop, including the empty-point compaction,
    the 12 px / depth pre-filter, the > 100 inlier rule and the shared-camera order."""
    import torch
    from vggsfm_b200 import pose_refinement as pr
    model = 1 if cam == "SIMPLE_RADIAL" else 0
    X, uv, inl, p0, i0 = _frames(model, seed=7, shared=shared)
    S, P = inl.shape
    N = P + 50                                           # tracks that are not valid
    rng = np.random.default_rng(1)
    valid = np.ones(N, bool)
    valid[rng.choice(N, 50, replace=False)] = False
    Xv = X.copy()
    Xv[5] = 0.0                                          # an "empty" point: dropped at :289-295
    tracks = rng.uniform(0, 1000, size=(S, N, 2)).astype(np.float32)
    tracks[:, valid] = uv
    inl_full = rng.uniform(size=(S, N)) < 0.5
    inl_full[:, valid] = inl
    ex = to_dev(i0[:, 3:4], cuda_dev) if model == 1 else None
    E, K, exo, vmask = pr.refine_pose(to_dev(p0, cuda_dev), to_dev(_K(i0), cuda_dev), ex, to_dev(inl_full, cuda_dev),
                                      to_dev(Xv, cuda_dev), to_dev(tracks, cuda_dev), to_dev(valid, cuda_dev),
                                      torch.tensor([1024, 768], device=cuda_dev), shared_camera=shared, camera_type=cam)
    keep = np.ones(P, bool)
    keep[5] = False
    pe, ie, used, summ = po.frame_loop(p0, i0, X[keep], uv[:, keep].astype(np.float64), inl[:, keep], np.ones(S, bool), model,
                                       shared, 12.0, 100)
    assert vmask.all()
    assert np.allclose(E.cpu().numpy(), pe, atol=1e-8)
    assert np.allclose(K.cpu().numpy(), _K(ie), rtol=1e-9)
    if model == 1:
        assert np.allclose(exo.cpu().numpy()[:, 0], ie[:, 3], atol=1e-9)
    else:
        assert exo is None
    rep = pr.last_report
    assert np.array_equal(rep.inlier_used.cpu().numpy(), used)
    assert [int(t) for t in rep.termination.cpu()] == [sm["termination"] for sm in summ]
    assert summ[3]["termination"] == 7 and bool(rep.needs_absolute_pose[3])
    if shared:
        assert torch.all(K[:, 0, 0] == K[0, 0, 0])


def test_init_refine_pose_mirror(cuda_dev):
    """init_refine_pose (triangulation.py:482-647): query frame counts all tracks, the init pair is skipped, > 50 rule."""
    import torch
    from vggsfm_b200 import pose_refinement as pr
    X, uv, inl, p0, i0 = _frames(0, seed=9)
    S, P = inl.shape
    init_idx = 4
    valid = np.ones(P, bool)
    valid[::7] = False
    E, K, exo, vmask = pr.init_refine_pose(to_dev(p0, cuda_dev), to_dev(_K(i0), cuda_dev), None, to_dev(inl[1:], cuda_dev),
                                           to_dev(X[valid], cuda_dev), to_dev(uv, cuda_dev), to_dev(valid, cuda_dev),
                                           torch.tensor([1024, 768], device=cuda_dev), init_idx)
    active = np.ones(S, bool)
    active[0] = active[init_idx + 1] = False
    inl2 = np.concatenate([np.ones((1, P), bool), inl[1:]], 0)[:, valid]
    pe, ie, _, summ = po.frame_loop(p0, i0, X[valid], uv[:, valid].astype(np.float64), inl2, active, 0, False, 0.0, 50)
    assert np.allclose(E.cpu().numpy(), pe, atol=1e-8) and np.allclose(K.cpu().numpy(), _K(ie), rtol=1e-9)
    assert np.array_equal(E.cpu().numpy()[0], p0[0]) and np.array_equal(E.cpu().numpy()[init_idx + 1], p0[init_idx + 1])
    assert exo is None and vmask.all()
    assert [int(t) for t in pr.last_report.termination.cpu()] == [sm["termination"] for sm in summ]


def test_sample_features4d(cuda_dev):
    """vgg_sample_features4d vs torch grid_sample (align_corners=True, border), the reference's own formulation
    (models/utils.py:347-447), in fp32: tolerance 2e-6 of the value range."""
    import torch
    import torch.nn.functional as F
    from vggsfm_b200.corr import sample_features4d
    torch.manual_seed(0)
    B, C, H, W, R = 3, 35, 37, 53, 500
    inp = torch.randn(B, C, H, W)
    coords = torch.rand(B, R, 2) * torch.tensor([W + 6.0, H + 6.0]) - 3.0       # some points outside -> border
    coords[0, :4] = torch.tensor([[0.0, 0.0], [W - 1.0, H - 1.0], [W - 1.0, 0.0], [10.0, H - 1.0]])
    g = coords.unsqueeze(2) * torch.tensor([2 / (W - 1), 2 / (H - 1)]) - 1
    ref = F.grid_sample(inp, g, align_corners=True, padding_mode="border").permute(0, 2, 1, 3).reshape(B, R, C)
    got = sample_features4d(inp.to(cuda_dev), coords.to(cuda_dev)).cpu()
    assert got.shape == (B, R, C)
    assert (got - ref).abs().max().item() < 2e-5


def test_sample_features4d_reference_golden(cuda_dev):
    """Same kernel against the output of the reference's own sample_features4d (tools/make_golden_host.py)."""
    import os
    import torch
    from vggsfm_b200.corr import sample_features4d
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "host_sample_features4d.npz"))
    got = sample_features4d(torch.from_numpy(d["input"]).to(cuda_dev), torch.from_numpy(d["coords"]).to(cuda_dev)).cpu().numpy()
    assert got.shape == d["out"].shape and np.abs(got - d["out"]).max() < 2e-5
