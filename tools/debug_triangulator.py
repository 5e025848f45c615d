"""Stage-by-stage log of Triangulator.forward on a synthetic scene (GPU)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from vggsfm_b200.synthetic import make_scene, perturb
from vggsfm_b200.triangulator import Triangulator

cam = sys.argv[1] if len(sys.argv) > 1 else "SIMPLE_RADIAL"
shared = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
S, N = int(sys.argv[3]) if len(sys.argv) > 3 else 12, int(sys.argv[4]) if len(sys.argv) > 4 else 1024
sc = make_scene(S, N, cam, seed=11, invisible_frac=0.1, outlier_frac=0.02, k=0.03)
extr0, K0, _, _ = perturb(sc, rot_deg=0.4, trans_frac=0.01, focal_frac=0.02, seed=12)
from vggsfm_b200.synthetic import project_np
uv_gt, _ = project_np(sc.extrinsics, 1000.0, np.array([512.0, 512.0]), 0.03 if cam == "SIMPLE_RADIAL" else 0.0, sc.points3d)
ok = np.linalg.norm(sc.tracks - uv_gt, axis=-1) < 3.0
fmat = ok[:1] & ok[1:]
dev = torch.device("cuda:0")
t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev if dt is None else dev, dt) if dt else torch.from_numpy(np.ascontiguousarray(a)).to(dev)


class C:
    pass


c = C()
c.focal_length = t(np.stack([K0[:, 0, 0], K0[:, 1, 1]], -1) * 2.0 / 1024, torch.float32)
c.R = t(extr0[:, :, :3], torch.float32)
c.T = t(extr0[:, :, 3], torch.float32)
tr = Triangulator()
tr.verbose = True
images = torch.rand(1, S, 3, 64, 64, device=dev)
torch.manual_seed(0)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = tr(c, t(sc.tracks)[None], t(sc.vis)[None], images.new_zeros(1, S, 3, 1024, 1024) if rep == 0 else images.new_zeros(1, S, 3, 1024, 1024),
             {"fmat_inlier_mask": t(fmat)[None]}, pred_score=t(sc.score)[None],
             BA_iters=2, shared_camera=shared, robust_refine=2, camera_type=cam)
    torch.cuda.synchronize(); print("forward s:", time.time() - t0)
    tr.verbose = False
E = out[0].cpu().numpy()
C_est = -np.einsum("sji,sj->si", E[:, :, :3], E[:, :, 3])
C_gt = -np.einsum("sji,sj->si", sc.extrinsics[:, :, :3], sc.extrinsics[:, :, 3])
print("centres est\n", C_est.round(3), "\ngt\n", C_gt.round(3))
