"""Multi-GPU plumbing for the path: one process per GPU (torch.distributed), reference views sharded
by rank, NO collective on the inference data path (every reference view is independent end to end,
SURVEY.md 8e).  The reference's only multi-device mechanism is nn.DataParallel (eval.py:33, train.py:282):
per-forward parameter broadcast, scatter, gather on device 0 -- none of which is needed here.

For the training configuration the one real exchange is the gradient all-reduce: 222,632 fp32 values
(0.89 MB), done as ONE flat-buffer all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).
Three parameter groups never receive a gradient with the default flags; they are
zero-filled in the flat buffer instead of tripping the reducer.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(num_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `num_items` reference views owned by `rank`; sizes differ by at most 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(num_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    """Slice dim 0 (the reference-view batch) of every tensor for this rank."""
    lo, hi = shard_range(tensors[0].shape[0], rank, world)
    return [t[lo:hi] for t in tensors]


def gather_depth_maps(local: torch.Tensor, num_items: int, group=None) -> torch.Tensor:
    """Optional: collect every rank's depth maps on all ranks (ragged last shard allowed).  Not on the
    timed inference path -- outputs normally stay on their GPU / are written out per rank."""
    world = dist.get_world_size(group)
    sizes = [shard_range(num_items, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


class FlatGradAllReduce:
    """Gradient averaging across ranks as ONE collective and ZERO copies: every parameter's `.grad` is a view into one flat
    fp32 buffer, autograd accumulates straight into it, and the step's exchange is a single all-reduce of that buffer
    (0.89 MB for the 222,632 parameters of the network; reference counterpart: nn.DataParallel's per-step gather of
    replica gradients, train.py:282).

        reducer = FlatGradAllReduce(net.parameters())      # attaches the gradient views
        reducer.zero_()                                    # instead of optimizer.zero_grad(): one memset
        loss.backward(); reducer(); optimizer.step()

    Parameters the graph never reaches (three groups with the default flags) simply keep zeros in their
    slice: every rank contributes to and receives every slice, so replicas cannot drift apart when a parameter has a
    gradient on some ranks only.  A parameter whose `.grad` was replaced or set to None behind the reducer's
    back (e.g. zero_grad(set_to_none=True)) is copied in / re-attached -- correct, just not copy-free."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None) -> None:
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=first.device)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat[off:off + n].view_as(p))
            off += n
        self.never_reached = None  # filled by the first call: parameters whose slice is exactly zero after backward
        self.attach()

    def attach(self) -> None:
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero_(self) -> None:
        """Start a step: one memset of the flat buffer (and re-attach views somebody dropped)."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v

    def __call__(self) -> int:
        """All-reduce (average) in place.  Returns the number of parameters that were not attached to the flat buffer on
        this rank (0 in the copy-free regime)."""
        detached = 0
        for p, v in zip(self.params, self.views):
            if p.grad is v:
                continue
            detached += 1
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
            p.grad = v
        if self.never_reached is None:  # one-off bookkeeping (host sync): which parameters the graph does not reach
            self.never_reached = sum(1 for v in self.views if not bool(v.any()))
        world = dist.get_world_size(self.group)
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)  # one kernel
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(world)
        return detached
