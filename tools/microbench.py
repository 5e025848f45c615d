"""CUDA-event micro-benchmarks for A/B runs (env switches are read once per process):
   python tools/microbench.py chol | ba | blocks [N]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vggsfm_b200 import _lib, bundle_adjustment as ba       # noqa: E402
from vggsfm_b200.synthetic import make_scene, perturb       # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "ba"
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VGG_"))


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if mode == "chol":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2402
    rng = np.random.default_rng(0)
    B = rng.normal(size=(n, n + 8))
    A = np.tril(B @ B.T + n * 1e-3 * np.eye(n))
    lda = (n + 127) // 128 * 128
    src = torch.zeros(n, lda, dtype=torch.float64, device=dev)
    src[:, :n] = torch.from_numpy(A).to(dev)
    buf = src.clone()
    ws = torch.empty(((n + 63) // 64) * 32768 + 256, dtype=torch.uint8, device=dev)
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream

    def run():
        buf.copy_(src)
        _lib.check(L.vgg_cholesky_lower(n, lda, buf.data_ptr(), ws.data_ptr(), ws.numel(), None, st), "chol")
    t_all = timeit(run)
    t_copy = timeit(lambda: buf.copy_(src))
    Afull = src[:, :n] + torch.tril(src[:, :n], -1).T
    t_torch = timeit(lambda: torch.linalg.cholesky(Afull))
    print(f"[{tag}] cholesky n={n}: own {t_all - t_copy:.3f} ms   torch.linalg.cholesky {t_torch:.3f} ms")
else:
    S = 400
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    sc = make_scene(S, N, "SIMPLE_RADIAL", seed=0)
    extr, K, extra, pts = perturb(sc, seed=1)
    t = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).to(dev).contiguous()
    intr = np.zeros((S, 4))
    intr[:, 0], intr[:, 1], intr[:, 2], intr[:, 3] = K[0, 0, 0], K[0, 0, 2], K[0, 1, 2], extra[0, 0]
    uv, mask = t(sc.tracks, torch.float32), t(sc.mask.astype(np.uint8))
    poses, intr_t, X = t(extr), t(intr), t(pts)
    model, mode_i = ba.SIMPLE_RADIAL, ba.INTR_SHARED
    if mode == "blocks":
        ms = timeit(lambda: ba.build_blocks(uv, mask, poses, intr_t, X, model, mode_i), reps=10)
        nbytes = S * N * 153
        print(f"[{tag}] build_blocks 400x{N}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s (153 B/obs)")
    else:
        opt = ba.default_options()
        opt.max_num_iterations = 10
        opt.gradient_tolerance = 0.0
        res = []

        def run():
            res.append(ba.lm_solve(uv, mask, poses.clone(), intr_t.clone(), X.clone(), model, mode_i, options=opt))
        ms = timeit(run, reps=5)
        s = res[-1]
        print(f"[{tag}] lm_solve 400x{N} 10 its: {ms:.2f} ms/solve -> {10e3 / ms:.1f} it/s  cost={s.final_cost:.6f} launches={s.kernel_launches}")
