"""Run a few conv layers of the 640x512 workload once each between cudaProfilerStart/Stop (for ncu --set full)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchmatchnet_b200 import ops  # noqa: E402

dev = "cuda:0"
LAYERS = [  # N, cin, cout, ks, S, pad, dil, h, w
    ("conv1", 5, 8, 8, 3, 1, 1, 1, 512, 640),
    ("conv3", 5, 16, 16, 3, 1, 1, 1, 256, 320),
    ("conv6", 5, 32, 32, 3, 1, 1, 1, 128, 160),
    ("conv5", 5, 16, 32, 5, 2, 2, 1, 256, 320),
    ("conv8", 5, 32, 64, 5, 2, 2, 1, 128, 160),
    ("conv9", 5, 64, 64, 3, 1, 1, 1, 64, 80),
    ("inner2", 5, 16, 64, 1, 1, 0, 1, 256, 320),
]
g = torch.Generator().manual_seed(0)
runs = []
for (name, N, cin, cout, ks, S, pad, dil, h, w) in LAYERS:
    x = torch.randn(N, cin, h, w, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, ks, ks, generator=g).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    frags = {1: ops.pack_conv_filter(wt, 1), 3: ops.pack_conv_filter(wt, 3)}
    frag = frags[1]
    runs.append((x, frag, b, cout, ks, S, pad, dil))
for r in runs:
    for _ in range(2):
        ops.conv2d_nhwc(*r, relu=True, precision=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for r in runs:
    ops.conv2d_nhwc(*r, relu=True, precision=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("probed", [l[0] for l in LAYERS])
