"""Kernel micro-benchmark: capture the K-A / K-B / K-C launches of one real 640x512 forward (so the
hypotheses have the cascade's real clustering) and time each one in isolation -- L2 flushed, CUDA events --
for a sweep of kernel variants selected through environment variables (read per call by libpmb200.so).

    python tools/kbench.py > gpurun_out/kbench.json
"""
import itertools
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from patchmatchnet_b200 import ops, synthetic  # noqa: E402

dev = "cuda:0"
torch.backends.cudnn.benchmark = True
H, W = int(os.environ.get("KB_H", 512)), int(os.environ.get("KB_W", 640))
net, _ = bench.build_net()
net = net.to(dev)
inp = synthetic.make_inputs(1, 5, H, W, seed=0)
args = lambda: ([i.to(dev) for i in inp["images"]], inp["intrinsics"].to(dev), inp["extrinsics"].to(dev), inp["depth_min"].to(dev), inp["depth_max"].to(dev))
names = ("warp_corr_score", "warp_corr_view_weights", "aggregate_views_score", "adaptive_eval", "init_propagate", "offset_corr_weight")
calls = []
origs = {n: getattr(ops, n) for n in names}
for n in names:
    def mk(n):
        def spy(*a, **k):
            calls.append((n, a, k))
            return origs[n](*a, **k)
        return spy
    setattr(ops, n, mk(n))
with torch.no_grad():
    for _ in range(2):
        calls.clear()
        torch.manual_seed(0)
        net(*args())
for n in names:
    setattr(ops, n, origs[n])
torch.cuda.synchronize()
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=12):
    """(cold mean us, cold min us, warm us): cold = L2 flushed before each launch; warm = 10 back-to-back launches."""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    torch.cuda._sleep(2_000_000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    b.synchronize()
    return round(statistics.mean(ts), 2), round(min(ts), 2), round(a.elapsed_time(b) * 1e2, 2)


out = {"shape": f"{W}x{H}", "rows": []}
variants = ([dict(PMB200_WARP_CORR_V1="1"), dict(PMB200_KA_GEN="2")]
            + [dict(PMB200_KA_GEN="3", PMB200_KA_DC=str(d), PMB200_KA_PIPE=str(pp)) for d, pp in itertools.product((0, 4, 8, 16), (0, 1))]
            + [dict(PMB200_KA_GEN="3", PMB200_KA_DC=str(d), PMB200_KA_PIPE="0", PMB200_KA_MINB="8") for d in (0, 4, 8, 16)]
            # C32 only (8 pixels per warp): pipelined gather capped at 96 registers -> 5 resident CTAs instead of 4
            + [dict(PMB200_KA_GEN="3", PMB200_KA_DC=str(d), PMB200_KA_PIPE="1", PMB200_KA_MINB="5") for d in (0, 8, 16)])
for (n, a, k) in calls:
    desc = n
    if n.startswith("warp_corr"):
        ref, src, depth = a[0], a[1], a[3]
        desc += f" C{ref.shape[3]} D{depth.shape[1]} {ref.shape[1]}x{ref.shape[2]} V{src.shape[0]}"
    elif n == "adaptive_eval":
        desc += f" D{a[1].shape[1]} {a[1].shape[2]}x{a[1].shape[3]}"
    elif n == "init_propagate":
        desc += f" Ns{a[5]} Kp{a[6]} {a[0].shape[2]}x{a[0].shape[3]}"
    elif n == "offset_corr_weight":
        desc += f" C{a[0].shape[3]} {a[0].shape[1]}x{a[0].shape[2]}"
    row = {"call": desc, "default_us": timeit(lambda: origs[n](*a, **k))}
    if n == "adaptive_eval" and os.environ.get("KB_SWEEP_EVAL", "1") == "1":
        for tp, dy in ((32, 8), (32, 4), (16, 16), (16, 8), (8, 32), (8, 16), (64, 4), (32, 2)):
            os.environ["PMB200_KB_TP"], os.environ["PMB200_KB_DY"] = str(tp), str(dy)
            row[f"TP={tp},DY={dy}"] = timeit(lambda: origs[n](*a, **k))
            os.environ.pop("PMB200_KB_TP", None)
            os.environ.pop("PMB200_KB_DY", None)
    if n == "warp_corr_view_weights" and os.environ.get("KB_SWEEP_KA", "1") == "1":
        for dc, pp in itertools.product((8, 16), (0, 1)):  # rows per warp pass / gather pipeline of the view-weights epilogue
            os.environ["PMB200_KA_DC_VW"], os.environ["PMB200_KA_PIPE"] = str(dc), str(pp)
            row[f"DC_VW={dc},PIPE={pp}"] = timeit(lambda: origs[n](*a, **k))
            os.environ.pop("PMB200_KA_DC_VW", None)
            os.environ.pop("PMB200_KA_PIPE", None)
    if n == "warp_corr_score" and os.environ.get("KB_SWEEP_KA", "1") == "1":
        for v in variants:
            for kk, vv in v.items():
                os.environ[kk] = vv
            tag = ",".join(f"{kk[7:]}={vv}" for kk, vv in v.items())
            row[tag] = timeit(lambda: origs[n](*a, **k))
            for kk in v:
                os.environ.pop(kk, None)
    out["rows"].append(row)
print(json.dumps(out, indent=1))
