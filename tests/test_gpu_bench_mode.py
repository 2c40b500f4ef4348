"""Parity of the network in the configuration bench.py TIMES: torch's default flags (cuDNN TF32
allowed -> the native convs run with TF32 operands, `ops.conv_precision() == 1`), `cudnn.benchmark` on, the forward
replayed from the DepthEngine's CUDA graph -- against the fp32 oracle (the reference's op sequence in full fp32 on the
same device, the stage-3 random draw shared through `rand_source`).

Sizes: BASELINE config 2 (640x512, 1+4 views), config 3 (1600x1184, 1+4 views), config 4's per-GPU batch taken to 8
reference views on one GPU (640x512).  Bound: north_star's 1e-3 relative L1 on the final depth map; the observed values
are printed and appended to gpurun_out/bench_mode_parity.json so the evidence travels back from the GPU box.

The other GPU parity tests run the library and native convs in full fp32 (3xTF32); this file is the one that holds the
timed mode itself to the bound."""
import json
import os

import pytest
import torch

from oracle import pm_oracle
from patchmatchnet_b200 import PatchmatchNet, load_reference_state, ops, synthetic
from tests import pm_cases

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
NORTH_STAR_TOL = 1e-3
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net(weights, cls=None):
    net = PatchmatchNet(**pm_cases.NET_KWARGS) if cls is None else PatchmatchNet(**pm_cases.NET_KWARGS, patchmatch_cls=cls)
    load_reference_state(net, weights)
    return net.eval().to(DEV)


def _record(row):
    path = os.path.join(REPO, "gpurun_out", "bench_mode_parity.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        rows = json.load(open(path)) if os.path.exists(path) else []
        rows.append(row)
        json.dump(rows, open(path, "w"), indent=1)
    except OSError:
        pass


@pytest.mark.parametrize("B,H,W,n_views,slots", [(1, 512, 640, 5, 3), (1, 1184, 1600, 5, 1), (8, 512, 640, 5, 1)],
                         ids=["cfg2_640x512", "cfg3_1600x1184", "batch8_640x512"])
def test_timed_configuration_matches_fp32_oracle(golden_weights, B, H, W, n_views, slots):
    from patchmatchnet_b200.engine import DepthEngine

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark,
           torch.backends.cudnn.deterministic)
    inp = synthetic.make_inputs(B, n_views, H, W, seed=5)
    g = torch.Generator().manual_seed(1234)
    rand48 = torch.rand(B, 48, H // 8, W // 8, generator=g).to(DEV)
    draw = lambda size, device: rand48
    try:
        # ---- the timed configuration: exactly what bench.py sets up ----
        torch.backends.cudnn.allow_tf32 = True          # torch's default
        torch.backends.cuda.matmul.allow_tf32 = False   # torch's default
        torch.backends.cudnn.benchmark = True           # bench.py / reference eval.py:301
        torch.backends.cudnn.deterministic = False
        assert ops.conv_precision() == 1, "native convs must run with TF32 operands in the timed mode"
        mine = _net(golden_weights)
        mine.patchmatch_3.rand_source = draw
        eng = DepthEngine(mine, B, n_views, H, W, device=DEV, use_graph=True, n_slots=slots)
        d_timed, c_timed = eng.infer(inp["images"], inp["intrinsics"], inp["extrinsics"], inp["depth_min"], inp["depth_max"])
        d_timed, c_timed = d_timed.clone(), c_timed.clone()
        d_again, _ = eng.infer(inp["images"], inp["intrinsics"], inp["extrinsics"], inp["depth_min"], inp["depth_max"])
        assert eng.use_graph and eng._slots[0]["graph"] is not None, "the timed mode replays a CUDA graph"
        replay_drift = pm_cases.rel_l1(d_again, d_timed)
        # ---- the same network with every conv in full fp32 (3xTF32 native, fp32 cuDNN): the parity configuration ----
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cudnn.benchmark = False
        mine32 = _net(golden_weights)
        mine32.patchmatch_3.rand_source = draw
        args = lambda: ([i.to(DEV) for i in inp["images"]], inp["intrinsics"].to(DEV), inp["extrinsics"].to(DEV),
                        inp["depth_min"].to(DEV), inp["depth_max"].to(DEV))
        with torch.no_grad():
            d_fp32, _, _ = mine32(*args())
            # ---- the fp32 oracle: reference op sequence, full fp32, same device, same random draw ----
            orc = _net(golden_weights, pm_oracle.PatchMatchOracle)
            orc.patchmatch_3.rand_source = draw
            d_orc, c_orc, _ = orc(*args())
        torch.cuda.synchronize()
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark,
         torch.backends.cudnn.deterministic) = old
    err_timed = pm_cases.rel_l1(d_timed, d_orc)
    err_fp32 = pm_cases.rel_l1(d_fp32, d_orc)
    conf_mean = float((c_timed.cpu() - c_orc.cpu()).abs().mean())
    row = {"case": f"B{B} {W}x{H} 1+{n_views - 1} views", "timed_mode_rel_l1_vs_fp32_oracle": err_timed,
           "fp32_mode_rel_l1_vs_fp32_oracle": err_fp32, "graph_replay_drift": replay_drift,
           "confidence_mean_abs_diff": conf_mean, "bound": NORTH_STAR_TOL}
    print("bench-mode parity:", json.dumps(row))
    _record(row)
    assert torch.isfinite(d_timed).all()
    assert replay_drift <= 1e-6, replay_drift
    assert err_fp32 <= 2e-4, err_fp32
    assert err_timed <= NORTH_STAR_TOL, err_timed


def test_eval_mode_with_grad_enabled_is_differentiable(golden_weights):
    """eval() with autograd on must not take the folded / native inference branches (constants built
    under no_grad, cudnn_convolution_relu has no derivative): every FeatureNet / Refinement parameter gets a gradient,
    as in the reference, which is fully differentiable in eval mode."""
    net = _net(golden_weights)
    inp = synthetic.make_inputs(1, 3, 64, 80, seed=2)
    fixed = torch.rand(1, 48, 8, 10, device=DEV)
    net.patchmatch_3.rand_source = lambda size, device: fixed
    for p in net.parameters():
        p.grad = None
    with torch.enable_grad():
        depth, conf, _ = net([i.to(DEV) for i in inp["images"]], inp["intrinsics"].to(DEV), inp["extrinsics"].to(DEV),
                             inp["depth_min"].to(DEV), inp["depth_max"].to(DEV))
        assert depth.requires_grad
        depth.mean().backward()
    missing = [n for n, p in net.named_parameters()
               if p.grad is None and (n.startswith("feature.") or n.startswith("upsample_net."))]
    assert not missing, missing
    with torch.no_grad():  # and the inference fast path still agrees with it
        d2, _, _ = net([i.to(DEV) for i in inp["images"]], inp["intrinsics"].to(DEV), inp["extrinsics"].to(DEV),
                       inp["depth_min"].to(DEV), inp["depth_max"].to(DEV))
    assert pm_cases.rel_l1(d2, depth) <= NORTH_STAR_TOL
