"""Gradient parity of the training path (native forward + native backward kernels through
patchmatchnet_b200.autograd) against torch autograd run on the oracle (CPU).  Needs a GPU: -m gpu.

Tolerance: gradients agree to ~1e-5 relative to the largest gradient entry (fp32, different summation
order; d_src is accumulated with float atomics so its last bits vary run to run)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import pm_oracle
from patchmatchnet_b200 import PatchMatch, autograd as ag, ops, synthetic
from tests import pm_cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _full_fp32_library_ops():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = old


def close(got, want, tol=2e-5):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = max(1e-12, float(want.abs().max()))
    err = float((got - want).abs().max())
    assert err <= tol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def compare_param_grads(mine, want, rtol, atol_of_global):
    """Per-parameter comparison with an absolute floor tied to the largest gradient of the model: parameters in
    front of a train-mode BatchNorm have (analytically) zero gradient along some directions, what is left of
    them is rounding noise of either implementation and must not be compared relatively."""
    assert set(mine) == set(want)
    glob = max(float(w.abs().max()) for w in want.values() if w is not None)
    bad = []
    for k, w in want.items():
        if w is None:  # parameters the reference graph never reaches
            assert mine[k] is None or float(mine[k].abs().max()) == 0.0, k
            continue
        assert mine[k] is not None, k
        err = float((mine[k].double() - w.double()).abs().max())
        lim = rtol * float(w.abs().max()) + atol_of_global * glob
        if err > lim:
            bad.append((k, err, float(w.abs().max())))
    assert not bad, f"global grad scale {glob:.3e}; offenders (name, max err, tensor scale): {bad[:8]}"


def _warp_case(B, V, C, H, W, D, seed):
    g = torch.Generator().manual_seed(seed)
    ref = torch.randn(B, C, H, W, generator=g)
    srcs = [torch.randn(B, C, H, W, generator=g) for _ in range(V)]
    Kc, Ec = synthetic.make_cameras(B, V + 1, H * 8, W * 8)
    ref_proj, src_projs = synthetic.stage_projections(Kc, Ec, 3)
    depth = torch.sort(350.0 + 700.0 * torch.rand(B, D, H, W, generator=g), dim=1)[0]
    depth[:, 0, :2] = -20.0
    vw = torch.rand(B, V, H, W, generator=g)
    return ref, srcs, ref_proj.contiguous(), [m.contiguous() for m in src_projs], depth, vw


@pytest.mark.parametrize("C,G,H,W,D,B,V", [(64, 8, 9, 13, 12, 2, 3), (32, 8, 11, 14, 16, 1, 2), (16, 4, 13, 18, 8, 2, 4)])
def test_warp_corr_backward(C, G, H, W, D, B, V):
    ref, srcs, ref_proj, src_projs, depth, vw = _warp_case(B, V, C, H, W, D, seed=C + D)
    gw = torch.randn(V, B, G, D, H, W, generator=torch.Generator().manual_seed(1))
    # oracle
    r = ref.clone().requires_grad_(True)
    ss = [s.clone().requires_grad_(True) for s in srcs]
    sims = torch.stack([pm_oracle.groupwise_correlation(pm_oracle.homography_warp(s, sp, ref_proj, depth), r, G) for s, sp in zip(ss, src_projs)])
    (sims * gw).sum().backward()
    # ours
    rt = ops.relative_projection(ref_proj.to(DEV), [m.to(DEV) for m in src_projs])
    rn = nhwc(ref.to(DEV)).requires_grad_(True)
    sn = torch.stack([nhwc(s.to(DEV)) for s in srcs]).requires_grad_(True)
    out = ag.WarpCorr.apply(rn, sn, rt, depth.to(DEV), None, G)
    close(out, sims, 2e-5)
    (out * gw.to(DEV)).sum().backward()
    close(rn.grad.permute(0, 3, 1, 2), r.grad)
    for v in range(V):
        close(sn.grad[v].permute(0, 3, 1, 2), ss[v].grad)
    # weighted-average mode (view weights detached, as on every iteration that uses it)
    gw2 = torch.randn(B, G, D, H, W, generator=torch.Generator().manual_seed(2))
    r.grad = None
    for s in ss:
        s.grad = None
    sims = torch.stack([pm_oracle.groupwise_correlation(pm_oracle.homography_warp(s, sp, ref_proj, depth), r, G) for s, sp in zip(ss, src_projs)])
    wsum = 1e-5 + vw.sum(1)
    agg = (sims * vw.permute(1, 0, 2, 3)[:, :, None, None]).sum(0) / wsum[:, None, None]
    (agg * gw2).sum().backward()
    rn.grad = None
    sn.grad = None
    out = ag.WarpCorr.apply(rn, sn, rt, depth.to(DEV), vw.to(DEV), G)
    (out * gw2.to(DEV)).sum().backward()
    close(rn.grad.permute(0, 3, 1, 2), r.grad)
    for v in range(V):
        close(sn.grad[v].permute(0, 3, 1, 2), ss[v].grad)


def test_pack_nhwc_backward():
    maps = [torch.randn(2, 16, 5, 7, device=DEV, requires_grad=True) for _ in range(3)]
    pack = ag.PackNHWC.apply(*maps)
    w = torch.randn_like(pack)
    (pack * w).sum().backward()
    for i, m in enumerate(maps):
        close(m.grad, w[i].permute(0, 3, 1, 2), 1e-7)


def test_aggregate_views_backward():
    g = torch.Generator().manual_seed(3)
    V, B, G, D, H, W = 3, 2, 8, 6, 5, 7
    sims = torch.randn(V, B, G, D, H, W, generator=g)
    vw = torch.rand(B, V, H, W, generator=g)
    gw = torch.randn(B, G, D, H, W, generator=g)
    s1, w1 = sims.clone().requires_grad_(True), vw.clone().requires_grad_(True)
    agg = (s1 * w1.permute(1, 0, 2, 3)[:, :, None, None]).sum(0) / (1e-5 + w1.sum(1))[:, None, None]
    (agg * gw).sum().backward()
    s2, w2 = sims.to(DEV).requires_grad_(True), vw.to(DEV).requires_grad_(True)
    out = ag.AggregateViews.apply(s2, w2)
    (out * gw.to(DEV)).sum().backward()
    close(s2.grad, s1.grad)
    close(w2.grad, w1.grad)


@pytest.mark.parametrize("C,G,K,dil,H,W,B", [(64, 8, 9, 2, 9, 13, 2), (32, 8, 9, 4, 11, 14, 1), (16, 4, 17, 4, 12, 16, 1)])
def test_offset_corr_backward(C, G, K, dil, H, W, B):
    g = torch.Generator().manual_seed(K + C)
    ref = torch.randn(B, C, H, W, generator=g)
    off = torch.randn(B, 2 * K, H, W, generator=g) * 2.5  # leaves the map at the borders: clamped positions get zero gradient
    gw = torch.randn(B, G, K, H, W, generator=g)
    o1 = off.clone().requires_grad_(True)
    grid = pm_oracle.sampling_grid(pm_oracle.neighbour_table("evaluation", K, dil), o1.view(B, 2 * K, H * W), H, W)
    corr = pm_oracle._FeatureWeightHead(K, G).neighbour_correlation(ref, grid)
    (corr * gw).sum().backward()
    o2 = off.to(DEV).requires_grad_(True)
    out = ag.OffsetCorr.apply(nhwc(ref.to(DEV)), o2, G, K, dil)
    (out * gw.to(DEV)).sum().backward()
    close(o2.grad, o1.grad, 5e-5)


@pytest.mark.parametrize("mode,Ns,Kp,dil,H,W,B", [("random", 48, 16, 2, 9, 13, 2), ("perturb", 16, 16, 2, 9, 13, 1), ("perturb", 8, 8, 4, 11, 14, 2), ("perturb", 8, 4, 4, 7, 9, 1)])
def test_init_propagate_backward(mode, Ns, Kp, dil, H, W, B):
    g = torch.Generator().manual_seed(Ns + Kp)
    dmin, dmax = torch.full((B,), 425.0), torch.full((B,), 935.0)
    scale = 0.025
    off = torch.randn(B, 2 * Kp, H, W, generator=g) * 2.0
    gw = torch.randn(B, Ns + Kp, H, W, generator=g)
    o1 = off.clone().requires_grad_(True)
    if mode == "random":
        u = torch.rand(B, 48, H, W, generator=g)
        init = pm_oracle.init_hypotheses(dmin, dmax, H, W, scale, 16, torch.empty(0), u.device, lambda size, device: u)
        seed, m = u, ops.MODE_RANDOM
    else:
        depth = 430.0 + 500.0 * torch.rand(B, 1, H, W, generator=g)
        init = pm_oracle.init_hypotheses(dmin, dmax, H, W, scale, Ns, depth, depth.device)
        seed, m = depth, ops.MODE_PERTURB
    grid = pm_oracle.sampling_grid(pm_oracle.neighbour_table("propagation", Kp, dil), o1.view(B, 2 * Kp, H * W), H, W)
    want = pm_oracle.propagate(init, grid)
    (want * gw).sum().backward()
    o2 = off.to(DEV).requires_grad_(True)
    hyp, xnorm = ag.InitPropagate.apply(seed.to(DEV), o2, dmin.to(DEV), dmax.to(DEV), m, Ns, Kp, dil, scale)
    assert not xnorm.requires_grad
    (hyp * gw.to(DEV)).sum().backward()
    close(o2.grad, o1.grad, 1e-4)


@pytest.mark.parametrize("D,K,dil,H,W,B,inverse", [(16, 9, 2, 9, 13, 2, False), (8, 9, 6, 11, 14, 1, True), (8, 17, 4, 10, 12, 1, False)])
def test_adaptive_eval_backward(D, K, dil, H, W, B, inverse):
    g = torch.Generator().manual_seed(D + K)
    dmin, dmax = torch.full((B,), 425.0), torch.full((B,), 935.0)
    scale = 0.0125
    depth = torch.sort(430.0 + 500.0 * torch.rand(B, D, H, W, generator=g), dim=1, descending=inverse)[0]
    score0 = torch.randn(B, D, H, W, generator=g) * 2.0
    off = torch.randn(B, 2 * K, H, W, generator=g) * 1.5
    fw = torch.rand(B, K, H, W, generator=g) + 0.05
    gdepth = torch.randn(B, H, W, generator=g)
    gprob = torch.randn(B, D, H, W, generator=g)

    def oracle(use_prob):
        s1, d1, o1, f1 = [t.clone().requires_grad_(True) for t in (score0, depth, off, fw)]
        grid = pm_oracle.sampling_grid(pm_oracle.neighbour_table("evaluation", K, dil), o1.view(B, 2 * K, H * W), H, W)
        w = pm_oracle.depth_similarity_weight(d1.detach(), dmin, dmax, grid.detach(), scale, K) * f1.unsqueeze(1)
        w = w / torch.sum(w, dim=2).unsqueeze(2)
        s = torch.sum(pm_oracle._border_sample(s1, grid).view(B, D, K, H, W) * w, dim=2)
        prob = torch.exp(F.log_softmax(s, dim=1))
        out = pm_oracle._Evaluation.regress(d1, prob, inverse)
        loss = (out * gdepth).sum() + ((prob * gprob).sum() if use_prob else 0.0)
        loss.backward()
        return s1.grad, d1.grad, o1.grad, f1.grad

    for use_prob in (False, True):
        want = oracle(use_prob)
        s2, d2, o2, f2 = [t.to(DEV).requires_grad_(True) for t in (score0, depth, off, fw)]
        inv_min, inv_max = (1.0 / dmin).view(B, 1, 1, 1), (1.0 / dmax).view(B, 1, 1, 1)
        xnorm = ((1.0 / depth - inv_max) / (inv_min - inv_max)).to(DEV)
        out, prob = ag.AdaptiveEval.apply(s2, d2, xnorm, o2, f2, dmin.to(DEV), dmax.to(DEV), dil, scale, inverse)
        loss = (out * gdepth.to(DEV)).sum() + ((prob * gprob.to(DEV)).sum() if use_prob else 0.0)
        loss.backward()
        for got, w_, tol in zip((s2.grad, d2.grad, o2.grad, f2.grad), want, (5e-5, 5e-5, 2e-4, 1e-4)):
            close(got, w_, tol)


@pytest.mark.parametrize("name", list(pm_cases.STAGE_CASES))
def test_stage_training_gradients_match_oracle(golden_weights, name):
    """A whole PatchMatch stage in train() mode (BatchNorm on batch statistics): loss = sum of smooth-L1 of every
    iteration's depth against a target, as in the reference's patchmatchnet_loss (net.py:336-342); parameter and
    input-feature gradients against torch autograd on the oracle (CPU)."""
    spec = pm_cases.STAGE_CASES[name]
    case = pm_cases.make_stage_inputs(spec)
    state = pm_cases.stage_state(golden_weights, spec["stage"])
    target = 500.0 + 300.0 * torch.rand(case["ref_feature"].shape[0], 1, *case["ref_feature"].shape[2:], generator=torch.Generator().manual_seed(9))

    def run(mod, dev):
        mod.load_state_dict(state, strict=True)
        mod = mod.to(dev).train()
        if case["rand48"] is not None:
            mod.rand_source = lambda size, device: case["rand48"].to(device)
        ref = case["ref_feature"].detach().clone().to(dev).requires_grad_(True)
        srcs = [s.detach().clone().to(dev).requires_grad_(True) for s in case["src_features"]]
        depths, score, vw = mod(
            ref_feature=ref, src_features=srcs, ref_proj=case["ref_proj"].to(dev), src_projs=[m.to(dev) for m in case["src_projs"]],
            depth_min=case["depth_min"].to(dev), depth_max=case["depth_max"].to(dev), depth=case["depth"].to(dev),
            view_weights=case["view_weights"].to(dev),
        )
        loss = sum(F.smooth_l1_loss(d, target.to(dev), reduction="mean") for d in depths)
        loss.backward()
        grads = {k: (p.grad.detach().cpu() if p.grad is not None else None) for k, p in mod.named_parameters()}
        return loss.item(), [d.detach().cpu() for d in depths], ref.grad.cpu(), [s.grad.cpu() for s in srcs], grads

    lo, do, gro, gso, po = run(pm_oracle.PatchMatchOracle(**pm_cases.stage_ctor_kwargs(spec["stage"])), "cpu")
    lm, dm, grm, gsm, pmine = run(PatchMatch(**pm_cases.stage_ctor_kwargs(spec["stage"])), DEV)
    for a, b in zip(dm, do):
        assert pm_cases.rel_l1(a, b) <= 1e-4
    assert abs(lm - lo) <= 1e-4 * abs(lo)
    close(grm, gro, 2e-3)
    for a, b in zip(gsm, gso):
        close(a, b, 2e-3)
    compare_param_grads(pmine, po, rtol=5e-3, atol_of_global=1e-4)


def test_network_training_step_matches_oracle(golden_weights):
    """BASELINE.json configs[4] in miniature: full cascade in train() mode, reference loss (net.py:321-342),
    backward; every parameter gradient against torch autograd on the oracle behind the same shell (CPU)."""
    from patchmatchnet_b200 import PatchmatchNet, load_reference_state, patchmatchnet_loss

    spec = dict(B=2, n_views=3, H=64, W=80, seed=41)
    inp = synthetic.make_inputs(spec["B"], spec["n_views"], spec["H"], spec["W"], seed=spec["seed"])
    g = torch.Generator().manual_seed(5)
    rand48 = torch.rand(spec["B"], 48, spec["H"] // 8, spec["W"] // 8, generator=g)
    gts, masks = [], []
    for lvl in range(4):
        h, w = spec["H"] >> lvl, spec["W"] >> lvl
        gts.append(500.0 + 350.0 * torch.rand(spec["B"], 1, h, w, generator=g))
        masks.append(torch.rand(spec["B"], 1, h, w, generator=g) > 0.2)

    def run(cls, dev):
        net = PatchmatchNet(**pm_cases.NET_KWARGS) if cls is None else PatchmatchNet(**pm_cases.NET_KWARGS, patchmatch_cls=cls)
        load_reference_state(net, golden_weights)
        net = net.to(dev).train()
        net.patchmatch_3.rand_source = lambda size, device: rand48.to(device)
        depth, conf, per_stage = net([i.to(dev) for i in inp["images"]], inp["intrinsics"].to(dev), inp["extrinsics"].to(dev),
                                     inp["depth_min"].to(dev), inp["depth_max"].to(dev))
        assert conf.numel() == 0  # train mode returns an empty confidence (net.py:286-287)
        loss = patchmatchnet_loss(per_stage, [t.to(dev) for t in gts], [m.to(dev) for m in masks])
        loss.backward()
        return loss.item(), {k: (p.grad.detach().cpu() if p.grad is not None else None) for k, p in net.named_parameters()}

    lo, go = run(pm_oracle.PatchMatchOracle, "cpu")
    lm, gm = run(None, DEV)
    assert abs(lm - lo) <= 2e-4 * abs(lo)
    never = sorted(k for k, v in go.items() if v is None)
    assert any("patchmatch_1.propa_conv" in k for k in never) and any("patchmatch_2.evaluation.pixel_wise_net" in k for k in never)
    # gradients flow through ~40 conv/BN layers computed by different libraries on CPU and GPU
    compare_param_grads(gm, go, rtol=2e-2, atol_of_global=2e-3)
