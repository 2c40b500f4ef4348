"""Synthetic DTU-shaped inputs (SURVEY.md 8d / BASELINE.md 3): random images, pinhole cameras with
fx = fy = 1.8 W, principal point at the centre, reference extrinsic = identity, source views rotated
by a few degrees about y and translated by ~0.1 depth_min along +-x / +-y, depth range 425..935.
Used by bench.py, __graft_entry__.smoke() and the tests (there is no network for real datasets)."""
from __future__ import annotations

import math
from typing import Dict, List

import torch

DEPTH_MIN, DEPTH_MAX = 425.0, 935.0

DEFAULT_NET_KWARGS = dict(  # reference eval.py:326-337 defaults
    patchmatch_interval_scale=[0.005, 0.0125, 0.025],
    propagation_range=[6, 4, 2],
    patchmatch_iteration=[1, 2, 2],
    patchmatch_num_sample=[8, 8, 16],
    propagate_neighbors=[0, 8, 16],
    evaluate_neighbors=[9, 9, 9],
)


def make_cameras(B: int, N: int, H: int, W: int, jitter: float = 0.0, gen: torch.Generator = None):
    """intrinsics [B,N,3,3], extrinsics [B,N,4,4] (float32, CPU)."""
    K = torch.zeros(B, N, 3, 3)
    K[:, :, 0, 0] = 1.8 * W
    K[:, :, 1, 1] = 1.8 * W
    K[:, :, 0, 2] = W / 2
    K[:, :, 1, 2] = H / 2
    K[:, :, 2, 2] = 1.0
    E = torch.eye(4).repeat(B, N, 1, 1)
    for b in range(B):
        for i in range(1, N):
            sign = 1.0 if i % 2 else -1.0
            ang = math.radians(min(5.0, 1.5 * i)) * sign
            if jitter:
                ang += jitter * float(torch.rand((), generator=gen) - 0.5) * 0.02
            c, s = math.cos(ang), math.sin(ang)
            E[b, i, :3, :3] = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
            base = 0.1 * DEPTH_MIN * (1.0 + 0.1 * b)
            tx = base * sign if i <= 2 else 0.0
            ty = 0.0 if i <= 2 else base * sign
            E[b, i, :3, 3] = torch.tensor([tx, ty, 0.0])
    return K, E


def make_inputs(B: int, n_views: int, H: int, W: int, seed: int = 0) -> Dict[str, object]:
    """images: list of n_views [B,3,H,W] U[0,1); cameras; depth range.  CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    images: List[torch.Tensor] = [torch.rand(B, 3, H, W, generator=g) for _ in range(n_views)]
    K, E = make_cameras(B, n_views, H, W)
    return {
        "images": images,
        "intrinsics": K,
        "extrinsics": E,
        "depth_min": torch.full((B,), DEPTH_MIN),
        "depth_max": torch.full((B,), DEPTH_MAX),
    }


def stage_projections(intrinsics: torch.Tensor, extrinsics: torch.Tensor, stage: int):
    """Per-stage 4x4 projection matrices as the caller builds them (reference net.py:226-231):
    returns (ref_proj [B,4,4], [src_proj [B,4,4], ...])."""
    scale = {3: 0.125, 2: 0.25, 1: 0.5}[stage]
    K = intrinsics.clone()
    K[:, :, :2] *= scale
    proj = extrinsics.clone()
    proj[:, :, :3, :4] = torch.matmul(K, extrinsics[:, :, :3, :4])
    mats = torch.unbind(proj, 1)
    return mats[0], list(mats[1:])
