"""Short driver for ncu captures: a few launches of every hot kernel at the C3 size (400 x 4096).

    ncu --set full --clock-control none --import-source on -k regex:'ba_blocks|syrk|tri_main' -c 6 \
        -o gpurun_out/prof python tools/profile_kernels.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vggsfm_b200 import bundle_adjustment as ba      # noqa: E402
from vggsfm_b200 import triangulation as tri         # noqa: E402
from vggsfm_b200.synthetic import make_scene, perturb  # noqa: E402

S, N = int(os.environ.get("PROF_S", 400)), int(os.environ.get("PROF_N", 4096))
dev = torch.device("cuda:0")
sc = make_scene(S, N, "SIMPLE_RADIAL", seed=0)
extr, K, extra, pts = perturb(sc, seed=1)
t = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).to(dev).contiguous()
intr = np.zeros((S, 4))
intr[:, 0], intr[:, 1], intr[:, 2], intr[:, 3] = K[0, 0, 0], K[0, 0, 2], K[0, 1, 2], extra[0, 0]
uv, mask = t(sc.tracks, torch.float32), t(sc.mask.astype(np.uint8))
poses, intr_t, X = t(extr), t(intr), t(pts)
model, mode = ba.SIMPLE_RADIAL, ba.INTR_SHARED
for _ in range(2):
    ba.build_blocks(uv, mask, poses, intr_t, X, model, mode)
torch.cuda.synchronize()
if os.environ.get("PROF_MODE", "all") == "blocks":
    print("blocks only")
    sys.exit(0)
opt = ba.default_options()
opt.max_num_iterations = 2
opt.gradient_tolerance = 0.0
s = ba.lm_solve(uv, mask, poses.clone(), intr_t.clone(), X.clone(), model, mode, options=opt)
print("BA", s)
tn = tri.cam_from_img(t(sc.tracks), t(sc.intrinsics))
torch.manual_seed(0)
pairs = tri.draw_ransac_pairs(S, 256)
for _ in range(2):
    p3, num, m = tri.triangulate_tracks(t(sc.extrinsics), tn, track_vis=t(sc.vis), track_score=t(sc.score), ransac_pairs=pairs)
torch.cuda.synchronize()
print("tri median err", float(np.median(np.linalg.norm(p3.cpu().numpy() - sc.points3d, axis=1))), "mean inliers", num.float().mean().item())
