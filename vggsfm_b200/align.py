"""Similarity alignment of camera sets -- the window-to-window alignment of the video runner.

Close transcription of ``align_camera_extrinsics`` / ``apply_transformation`` (vggsfm/utils/align.py:145-252; the same
dozen tensor statements, kept statement for statement because the semantics must match), OpenCV convention,
extrinsics [B,3,4] = R|t.  A dozen 3x3 operations on at most window_size+1 cameras: host-side glue, written with
torch so that it runs wherever the caller's tensors live.
"""
from __future__ import annotations

import torch


def align_camera_extrinsics(cameras_src, cameras_tgt, estimate_scale=True, eps=1e-9):
    """Rotation by orthogonal Procrustes on mean(R_tgt^T R_src), scale by matching the covariances of the camera
    translations expressed in the source frames, translation from the means (align.py:145-205).
    Returns (align_t_R [1,3,3], align_t_T [1,3], align_t_s)."""
    R_src, R_tgt = cameras_src[:, :, :3], cameras_tgt[:, :, :3]
    M = torch.matmul(R_tgt.transpose(1, 2), R_src).mean(dim=0)
    U, _, V = torch.svd(M)
    rot = V @ U.t()
    a = torch.einsum("bi,bij->bj", cameras_src[:, :, 3], R_src)        # t^T R per camera
    b = torch.einsum("bi,bij->bj", cameras_tgt[:, :, 3], R_src)
    a_mu, b_mu = a.mean(dim=0, keepdim=True), b.mean(dim=0, keepdim=True)
    if estimate_scale and a.shape[0] > 1:
        ac, bc = a - a_mu, b - b_mu
        scale = (ac * bc).mean() / (ac ** 2).mean().clamp(eps)
    else:
        scale = 1.0
    return rot[None], b_mu - scale * a_mu, scale


def apply_transformation(cameras_src, align_t_R, align_t_T, align_t_s, return_extri=True):
    """R' = R A, t' = R T + s t (align.py:208-252)."""
    R_src, T_src = cameras_src[:, :, :3], cameras_src[:, :, 3]
    n = R_src.shape[0]
    aligned_R = torch.matmul(R_src, align_t_R.expand(n, 3, 3))
    aligned_T = torch.matmul(R_src, align_t_T[..., None].expand(n, 3, 1))[..., 0] + T_src * align_t_s
    if return_extri:
        return torch.cat([aligned_R, aligned_T.unsqueeze(-1)], dim=-1)
    return aligned_R, aligned_T
