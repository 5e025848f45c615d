"""Deterministic synthetic scenes for the parity tests and bench.py.

Points X ~ N((0,0,4), 0.5^2 I); cameras on an arc (yaw 0.6*i/S rad, t = (-2*i/S, 0, 0));
K = diag(f, f, 1) with f = 1000 px and principal point (512, 512) on a 1024^2 image;
SIMPLE_RADIAL k = 0.05; uv = pi(X) + N(0, noise^2 px).  numpy only (no torch, no CUDA) so the same
arrays feed the CUDA path, the oracle and the CPU baseline.
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class Scene:
    extrinsics: np.ndarray      # [S,3,4] float64 ground truth
    intrinsics: np.ndarray      # [S,3,3] float64
    extra_params: np.ndarray | None  # [S,1] float64 or None
    points3d: np.ndarray        # [N,3] float64 ground truth
    tracks: np.ndarray          # [S,N,2] float32 pixel observations
    vis: np.ndarray             # [S,N] float32 in (0,1]
    score: np.ndarray           # [S,N] float32
    mask: np.ndarray            # [S,N] bool: observation usable (vis>0.05)
    image_size: np.ndarray      # [2]
    camera_type: str


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    R = np.zeros(a.shape + (3, 3))
    R[..., 0, 0] = c
    R[..., 0, 2] = s
    R[..., 1, 1] = 1.0
    R[..., 2, 0] = -s
    R[..., 2, 2] = c
    return R


def project_np(extr, f, pp, k, X):
    """pixel projection of X[N,3] into every frame: [S,N,2], depth [S,N]."""
    p = np.einsum("sij,nj->sni", extr[:, :, :3], X) + extr[:, None, :, 3]
    u = p[..., 0] / p[..., 2]
    v = p[..., 1] / p[..., 2]
    d = 1.0 + k * (u * u + v * v)
    return np.stack([f * d * u + pp[0], f * d * v + pp[1]], -1), p[..., 2]


def make_scene(S, N, camera_type="SIMPLE_PINHOLE", noise_px=0.3, seed=0, invisible_frac=0.0,
               outlier_frac=0.0, focal=1000.0, k=0.05) -> Scene:
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, 3)) * 0.5 + np.array([0.0, 0.0, 4.0])
    i = np.arange(S) / float(S)
    R = _rot_y(0.6 * i)
    t = np.stack([-2.0 * i, np.zeros(S), np.zeros(S)], -1)
    extr = np.concatenate([R, t[:, :, None]], axis=2)
    pp = np.array([512.0, 512.0])
    kk = k if camera_type == "SIMPLE_RADIAL" else 0.0
    uv, _ = project_np(extr, focal, pp, kk, X)
    uv = uv + rng.normal(size=uv.shape) * noise_px
    vis = rng.uniform(0.06, 1.0, size=(S, N))
    if invisible_frac > 0:
        vis[rng.uniform(size=(S, N)) < invisible_frac] = 0.01
    if outlier_frac > 0:
        out = rng.uniform(size=(S, N)) < outlier_frac
        uv[out] = rng.uniform(0, 1024, size=(int(out.sum()), 2))
    K = np.zeros((S, 3, 3))
    K[:, 0, 0] = K[:, 1, 1] = focal
    K[:, 0, 2] = pp[0]
    K[:, 1, 2] = pp[1]
    K[:, 2, 2] = 1.0
    extra = np.full((S, 1), kk) if camera_type == "SIMPLE_RADIAL" else None
    return Scene(extr, K, extra, X, uv.astype(np.float32), vis.astype(np.float32),
                 np.ones((S, N), np.float32), vis > 0.05, np.array([1024, 1024]), camera_type)


def _exp_so3(w):
    th = np.linalg.norm(w, axis=-1, keepdims=True)
    th = np.where(th < 1e-12, 1e-12, th)
    a = w / th
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -a[..., 2], a[..., 1]
    K[..., 1, 0], K[..., 1, 2] = a[..., 2], -a[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -a[..., 1], a[..., 0]
    s = np.sin(th)[..., None]
    c = np.cos(th)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def perturb(scene: Scene, rot_deg=1.0, trans_frac=0.02, focal_frac=0.05, point_sigma=0.02, seed=1):
    """BA starting point: GT cameras perturbed by rot_deg / trans_frac / focal_frac, points by
    point_sigma.  Returns (extrinsics, intrinsics, extra_params, points3d)."""
    rng = np.random.default_rng(seed)
    S = scene.extrinsics.shape[0]
    w = rng.normal(size=(S, 3))
    w = w / np.linalg.norm(w, axis=1, keepdims=True) * np.deg2rad(rot_deg) * rng.uniform(0.5, 1.0, size=(S, 1))
    extr = scene.extrinsics.copy()
    extr[:, :, :3] = _exp_so3(w) @ extr[:, :, :3]
    extr[:, :, 3] += rng.normal(size=(S, 3)) * trans_frac
    K = scene.intrinsics.copy()
    fs = 1.0 + focal_frac * rng.uniform(-1, 1)
    K[:, 0, 0] *= fs
    K[:, 1, 1] *= fs
    extra = None if scene.extra_params is None else scene.extra_params * 0.5
    pts = scene.points3d + rng.normal(size=scene.points3d.shape) * point_sigma
    return extr, K, extra, pts


@dataclasses.dataclass
class VideoScene:
    """Forward-moving camera over a long corridor of points (BASELINE.json configs[4]: the sequential video runner).
    Point j is created in window ``birth[j]`` and observed in the frames of ``lifetime`` consecutive windows."""
    extrinsics: np.ndarray      # [F,3,4] ground truth
    focal: float
    pp: np.ndarray              # [2]
    points3d: np.ndarray        # [P,3] ground truth
    birth: np.ndarray           # [P] window index
    first_frame: np.ndarray     # [P]
    last_frame: np.ndarray      # [P] exclusive
    window: int
    init_window: int
    noise_px: float
    seed: int

    def window_range(self, w):
        """frames [start, end) of window w (window 0 is the init window)."""
        if w == 0:
            return 0, self.init_window
        s = self.init_window + (w - 1) * self.window
        return s, min(s + self.window, self.extrinsics.shape[0])

    def num_windows(self):
        F = self.extrinsics.shape[0]
        return 1 + (F - self.init_window + self.window - 1) // self.window

    def observe(self, ids, f0, f1):
        """noisy pixel tracks [f1-f0, len(ids), 2] float32 and the usable mask (inside the point's lifetime, in front of the
        camera, inside the 1024^2 image).  Noise is a pure function of (seed, frame, point): every call agrees."""
        ids = np.asarray(ids)
        uv, z = project_np(self.extrinsics[f0:f1], self.focal, self.pp, 0.0, self.points3d[ids])
        fr = np.arange(f0, f1)[:, None]
        h = (fr.astype(np.uint64) * np.uint64(2654435761) + ids[None].astype(np.uint64) * np.uint64(40503) + np.uint64(self.seed)) % np.uint64(2 ** 31)
        rng_u = (h.astype(np.float64) + 0.5) / 2 ** 31
        h2 = (h * np.uint64(1103515245) + np.uint64(12345)) % np.uint64(2 ** 31)
        rng_v = (h2.astype(np.float64) + 0.5) / 2 ** 31
        # Box-Muller from the two hashed uniforms
        g0 = np.sqrt(-2.0 * np.log(rng_u)) * np.cos(2 * np.pi * rng_v)
        g1 = np.sqrt(-2.0 * np.log(rng_u)) * np.sin(2 * np.pi * rng_v)
        uv = uv + np.stack([g0, g1], -1) * self.noise_px
        ok = (fr >= self.first_frame[ids][None]) & (fr < self.last_frame[ids][None]) & (z > 0.1)
        ok &= (uv[..., 0] > 0) & (uv[..., 0] < 1024) & (uv[..., 1] > 0) & (uv[..., 1] < 1024)
        return uv.astype(np.float32), ok


def make_video_scene(F=1000, window=16, init_window=32, new_per_window=512, lifetime=3, step=0.06, noise_px=0.3,
                     focal=1000.0, seed=0) -> VideoScene:
    """Camera i at (i*step, 0, 0) with a slow yaw wobble, looking down +z; the points of window w sit around the camera
    positions of the middle of their lifetime at depth ~4, so each is seen in about ``lifetime`` windows."""
    rng = np.random.default_rng(seed)
    i = np.arange(F, dtype=np.float64)
    R = _rot_y(0.05 * np.sin(i / 40.0))
    C = np.stack([i * step, 0.02 * np.sin(i / 25.0), np.zeros(F)], -1)
    t = -np.einsum("fij,fj->fi", R, C)
    extr = np.concatenate([R, t[:, :, None]], axis=2)
    sc = VideoScene(extr, focal, np.array([512.0, 512.0]), np.zeros((0, 3)), np.zeros(0, int), np.zeros(0, int), np.zeros(0, int),
                    window, init_window, noise_px, seed)
    pts, birth, f0s, f1s = [], [], [], []
    for w in range(sc.num_windows()):
        s, e = sc.window_range(w)
        last = min(F, s + lifetime * window) if w else min(F, init_window + (lifetime - 1) * window)
        mid = 0.5 * (s + last) * step
        n = new_per_window * (2 if w == 0 else 1)
        X = np.stack([mid + rng.uniform(-1.2, 1.2, n), rng.normal(0, 0.5, n), rng.normal(4.0, 0.5, n)], -1)
        pts.append(X)
        birth.append(np.full(n, w))
        f0s.append(np.full(n, s))
        f1s.append(np.full(n, last))
    sc.points3d = np.concatenate(pts)
    sc.birth = np.concatenate(birth)
    sc.first_frame = np.concatenate(f0s)
    sc.last_frame = np.concatenate(f1s)
    return sc
