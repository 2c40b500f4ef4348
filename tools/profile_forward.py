"""Run ONE eager forward of the bench workload between cudaProfilerStart/Stop so that
`ncu --profile-from-start off ...` sees exactly one step (see tools/gpu_*.sh)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from patchmatchnet_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=512)
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--views", type=int, default=5)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
dev = "cuda:0"
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = os.environ.get("PM_TF32", "0") == "1"  # default: the fp32-accurate mode bench.py times
net, _ = bench.build_net()
net = net.to(dev)
inp = synthetic.make_inputs(a.batch, a.views, a.height, a.width, seed=0)
args = lambda: ([i.to(dev) for i in inp["images"]], inp["intrinsics"].to(dev), inp["extrinsics"].to(dev),
                inp["depth_min"].to(dev), inp["depth_max"].to(dev))
with torch.no_grad():
    for _ in range(a.warmup):
        net(*args())
    torch.cuda.synchronize()
    x = args()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net(*x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one forward")
