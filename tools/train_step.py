"""BASELINE.json configs[4]: training step (forward + reference loss + backward + gradient all-reduce + Adam)
of the full cascade, 1 ref + 4 src, 640x512, B reference views per GPU.  One process per GPU:

    python tools/train_step.py --batch 2 --steps 10                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_step.py --batch 2

Reports samples/s (max over ranks, CUDA events) and the share of the step spent in the one collective of the
path: a flat fp32 all-reduce of the 222,632 gradients (NCCL over NVLink).  The reference's counterpart is
train.py:127-171 under nn.DataParallel (Adam lr 1e-3, train.py:284)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from patchmatchnet_b200 import distributed as pmd, patchmatchnet_loss, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--height", type=int, default=512)
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--views", type=int, default=5)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.backends.cudnn.benchmark = True
dist = None
if world > 1:
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

net, _ = bench.build_net()
net = net.to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))
reducer = pmd.FlatGradAllReduce(net.parameters())  # .grad tensors become views of one flat buffer (also at N = 1)
B, N, H, W = a.batch, a.views, a.height, a.width
inp = synthetic.make_inputs(B, N, H, W, seed=rank)
images = [i.to(dev) for i in inp["images"]]
K, E = inp["intrinsics"].to(dev), inp["extrinsics"].to(dev)
dmin, dmax = inp["depth_min"].to(dev), inp["depth_max"].to(dev)
g = torch.Generator().manual_seed(100 + rank)
gts = [(500.0 + 350.0 * torch.rand(B, 1, H >> l, W >> l, generator=g)).to(dev) for l in range(4)]
masks = [(torch.rand(B, 1, H >> l, W >> l, generator=g) > 0.1).to(dev) for l in range(4)]


def step():
    reducer.zero_()  # one memset instead of per-parameter zero_grad; never-reached parameters keep a zero gradient
    _, _, per_stage = net([i.clone() for i in images], K.clone(), E, dmin, dmax)
    loss = patchmatchnet_loss(per_stage, gts, masks)
    loss.backward()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    detached = reducer() if dist is not None else 0
    e1.record()
    opt.step()
    return loss, e0, e1, detached


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
comm = []
for _ in range(a.steps):
    loss, e0, e1, missing = step()
    comm.append((e0, e1))
e.record()
torch.cuda.synchronize()
sec = s.elapsed_time(e) * 1e-3
comm_s = sum(x.elapsed_time(y) for x, y in comm) * 1e-3
if dist is not None:
    t = torch.tensor([sec, comm_s], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec, comm_s = float(t[0]), float(t[1])
if rank == 0:
    print(json.dumps({
        "metric": "training samples/s (1ref+4src, 640x512, fwd+loss+bwd+allreduce+Adam)", "value": a.steps * B * world / sec,
        "unit": "samples/s", "n_gpus": world, "steps": a.steps, "ms_per_step": 1e3 * sec / a.steps,
        "allreduce_ms_per_step": 1e3 * comm_s / a.steps, "allreduce_bytes": 4 * sum(p.numel() for p in net.parameters()),
        "params_without_grad": reducer.never_reached if dist is not None else sum(1 for v in reducer.views if not bool(v.any())),
        "grads_not_views_of_the_flat_buffer": missing, "collectives_per_step": 1 if dist is not None else 0,
        "batch_per_gpu": B, "loss": float(loss), "data": "synthetic",
    }), flush=True)
if dist is not None:
    dist.destroy_process_group()
