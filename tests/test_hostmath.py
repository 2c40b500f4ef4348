"""CPU checks of the per-element formulas the kernels inline (patchmatchnet_b200/csrc/pm_math.cuh,
compiled for the host by tests/hostmath.cpp) against the oracle.  The kernels themselves are
checked on the GPU box (tests/test_gpu_parity.py); this catches convention bugs -- tap order,
the align_corners mismatch, (dy,dx) tables, sentinel handling -- without spending GPU time."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import pm_oracle
from patchmatchnet_b200 import synthetic
from tests import pm_cases


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _gather_from_cells(fmap, w, key, cols):
    """fmap [C,rows*cols] -> bilinear value per cell [C,N] using the packed keys (numpy)."""
    r0 = key & ((1 << 29) - 1)
    dx = (key >> 29) & 1
    dy = (key >> 30) & 1
    none = key == -2
    r0 = np.where(none, 0, r0)
    dx = np.where(none, 0, dx)
    dy = np.where(none, 0, dy)
    taps = [r0, r0 + dx, r0 + dy * cols, r0 + dy * cols + dx]
    out = np.zeros((fmap.shape[0], key.shape[0]), dtype=np.float64)
    for t in range(4):
        out += fmap[:, taps[t]].astype(np.float64) * w[:, t][None, :]
    out[:, none] = 0.0
    return out


@pytest.mark.parametrize("Hs,Ws", [(13, 21), (9, 17)])
def test_warp_cells_match_oracle_warp(hostmath, Hs, Ws):
    torch.manual_seed(3)
    B, C, H, W, D = 1, 8, 13, 21, 6
    Kc, Ec = synthetic.make_cameras(B, 2, H * 8, W * 8)
    ref_proj, src_projs = synthetic.stage_projections(Kc, Ec, 3)
    src = torch.randn(B, C, Hs, Ws)
    depth = 300.0 + 900.0 * torch.rand(B, D, H, W)
    depth[:, 0, :3] = -10.0  # behind the camera
    depth[:, 1, 5:8] = 1e-6  # z ~ 0
    want = pm_oracle.homography_warp(src, src_projs[0], ref_proj, depth)[0]  # [C,D,H,W]
    rot, trans = pm_oracle.relative_projection(src_projs[0], ref_proj)
    rt = np.concatenate([rot[0].reshape(-1).numpy(), trans[0].reshape(-1).numpy()]).astype(np.float32)
    dn = depth[0].reshape(-1).numpy().astype(np.float32).copy()
    w = np.zeros((D * H * W, 4), dtype=np.float32)
    key = np.zeros(D * H * W, dtype=np.int32)
    hostmath.hm_warp_cells(_fptr(rt), _fptr(dn), H, W, Hs, Ws, D, _fptr(w), _iptr(key))
    got = _gather_from_cells(src[0].reshape(C, -1).numpy(), w, key, Ws).reshape(C, D, H, W)
    err = np.abs(got - want.numpy()).max()
    assert err < 5e-4 * max(1.0, float(want.abs().max())), err
    if (Hs, Ws) == (H, W):  # with a smaller source map the (W,H) sentinel can land on the last texel, in the reference too
        assert np.abs(got[:, 0, :3]).max() == 0.0


@pytest.mark.parametrize("kind,K,dil", [("evaluation", 9, 2), ("evaluation", 17, 4), ("evaluation", 9, 6),
                                        ("propagation", 4, 2), ("propagation", 8, 4), ("propagation", 16, 2)])
def test_neighbour_cells_match_grid_sample(hostmath, kind, K, dil):
    torch.manual_seed(5)
    B, C, H, W = 1, 3, 11, 14
    off = torch.randn(B, 2 * K, H * W) * 3.0  # large enough to leave the map on every side
    fmap = torch.randn(B, C, H, W)
    table = pm_oracle.neighbour_table(kind, K, dil)
    grid = pm_oracle.sampling_grid(table, off, H, W)
    want = pm_oracle._border_sample(fmap, grid).view(C, K, H * W)
    dy = np.zeros(K, dtype=np.int32)
    dx = np.zeros(K, dtype=np.int32)
    assert hostmath.hm_neighbour_table(int(kind == "evaluation"), K, dil, _iptr(dy), _iptr(dx)) == 0
    assert [(int(a), int(b)) for a, b in zip(dy, dx)] == [tuple(t) for t in table]
    w = np.zeros((K * H * W, 4), dtype=np.float32)
    key = np.zeros(K * H * W, dtype=np.int32)
    on = off[0].numpy().astype(np.float32).copy()
    assert hostmath.hm_neighbour_cells(_fptr(on), int(kind == "evaluation"), K, dil, H, W, _fptr(w), _iptr(key)) == 0
    got = _gather_from_cells(fmap[0].reshape(C, -1).numpy(), w, key, W).reshape(C, K, H * W)
    assert np.abs(got - want.numpy()).max() < 2e-5


def test_unsupported_neighbour_counts(hostmath):
    dy = np.zeros(32, dtype=np.int32)
    dx = np.zeros(32, dtype=np.int32)
    assert hostmath.hm_neighbour_table(1, 10, 2, _iptr(dy), _iptr(dx)) == -1
    assert hostmath.hm_neighbour_table(0, 5, 2, _iptr(dy), _iptr(dx)) == -1


def test_hypothesis_formulas(hostmath):
    dmin, dmax = torch.tensor([425.0]), torch.tensor([935.0])
    inv_min, inv_max = float(1.0 / dmin), float(1.0 / dmax)
    u = torch.rand(1, 48, 2, 3)
    want = pm_oracle.init_hypotheses(dmin, dmax, 2, 3, 0.025, 16, torch.empty(0), u.device, lambda size, device: u)
    for k in (0, 24, 47):
        got = hostmath.hm_random_hypothesis(float(u[0, k, 1, 2]), k, inv_min, inv_max)
        assert abs(got - float(want[0, k, 1, 2])) <= 1e-4
    for ns in (8, 16, 3):
        depth = torch.tensor([430.0, 600.0, 930.0]).view(1, 1, 1, 3)
        want = pm_oracle.init_hypotheses(dmin, dmax, 1, 3, 0.025, ns, depth, depth.device)
        for k in range(ns):
            for j in range(3):
                got = hostmath.hm_perturbed_hypothesis(float(depth[0, 0, 0, j]), k, ns, inv_min, inv_max, 0.025)
                assert abs(got - float(want[0, k, 0, j])) <= 2e-4, (ns, k, j)


def test_depth_similarity_formula(hostmath):
    for xc, xn in ((0.3, 0.31), (0.3, 0.9), (0.5, 0.5), (0.1, 0.0)):
        t = min(max(abs(xn - xc) / 0.025, 0.0), 4.0)
        want = float(torch.sigmoid(torch.tensor(4.0 - 2.0 * t)))
        assert abs(hostmath.hm_depth_similarity(xc, xn, 0.025) - want) < 1e-6
