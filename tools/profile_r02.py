"""Driver for the round-2 ncu captures (one workload per mode, a few launches each):
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> -o gpurun_out/<name> \
        python tools/profile_r02.py chol|corr_tc|corr_fine|blocks
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vggsfm_b200 import _lib, bundle_adjustment as ba   # noqa: E402
from vggsfm_b200.corr import CorrBlock                  # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1]
if mode == "chol":
    n = 2403
    rng = np.random.default_rng(0)
    B = rng.normal(size=(n, n + 8))
    A = np.tril(B @ B.T + n * 1e-3 * np.eye(n))
    lda = (n + 127) // 128 * 128
    src = torch.zeros(n, lda, dtype=torch.float64, device=dev)
    src[:, :n] = torch.from_numpy(A).to(dev)
    ws = torch.empty(((n + 127) // 128) * 131072 + 1024, dtype=torch.uint8, device=dev)
    L = _lib.lib()
    for _ in range(2):
        buf = src.clone()
        _lib.check(L.vgg_cholesky_lower(n, lda, buf.data_ptr(), ws.data_ptr(), ws.numel(), None, torch.cuda.current_stream().cuda_stream), "chol")
    torch.cuda.synchronize()
elif mode in ("corr_tc", "corr_fine"):
    B, S, C, H, W, N, Lv, r = (1, 128, 128, 128, 128, 1024, 5, 4) if mode == "corr_tc" else (256, 128, 32, 31, 31, 1, 3, 3)
    g = torch.Generator(device=dev).manual_seed(0)
    fm = torch.randn(B, S, C, H, W, device=dev, dtype=torch.float16, generator=g)
    tg = torch.randn(B, S, N, C, device=dev, generator=g)
    co = torch.rand(B, S, N, 2, device=dev, generator=g) * torch.tensor([W - 9.0, H - 9.0], device=dev) + 4.0
    cb = CorrBlock(fm, num_levels=Lv, radius=r, half=True)
    for _ in range(2):
        cb.corr(tg)
        out = cb.sample(co)
    torch.cuda.synchronize()
elif mode == "blocks":
    from vggsfm_b200.synthetic import make_scene, perturb
    S, N = 400, int(os.environ.get("PROF_N", 4096))
    sc = make_scene(S, N, "SIMPLE_RADIAL", seed=0)
    extr, K, extra, pts = perturb(sc, seed=1)
    t = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).to(dev).contiguous()
    intr = np.zeros((S, 4))
    intr[:, 0], intr[:, 1], intr[:, 2], intr[:, 3] = K[0, 0, 0], K[0, 0, 2], K[0, 1, 2], extra[0, 0]
    for _ in range(3):
        ba.build_blocks(t(sc.tracks, torch.float32), t(sc.mask.astype(np.uint8)), t(extr), t(intr), t(pts), ba.SIMPLE_RADIAL, ba.INTR_SHARED)
    torch.cuda.synchronize()
print("done", mode)
