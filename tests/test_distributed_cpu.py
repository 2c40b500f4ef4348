"""world_size-2 gloo tests (CPU) of the N>1 path: rank sharding of reference views (no data-path
collective), the optional ragged gather, and the flat-buffer gradient all-reduce used by the training
configuration, including parameters that never receive a gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from patchmatchnet_b200 import distributed as pmd


def test_shard_ranges_partition_exactly():
    for n in (1, 7, 8, 9, 16):
        for world in (1, 2, 3, 8):
            got = [pmd.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pmd.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. sharded "inference": each rank handles its slice, results gathered only for the check
        n_items = 5
        images = torch.arange(n_items * 6, dtype=torch.float32).view(n_items, 6)
        (mine,) = pmd.shard_batch([images], rank, world)
        local = mine.sum(dim=1, keepdim=True) * 2.0  # stand-in for the per-view computation
        full = pmd.gather_depth_maps(local, n_items)
        assert torch.equal(full, images.sum(dim=1, keepdim=True) * 2.0)
        # 2. gradient all-reduce: .grad tensors are views of ONE flat buffer, a never-used parameter keeps zeros
        torch.manual_seed(0)
        lin = torch.nn.Linear(4, 3)
        unused = torch.nn.Parameter(torch.ones(5))
        reducer = pmd.FlatGradAllReduce(list(lin.parameters()) + [unused])
        assert lin.weight.grad.data_ptr() == reducer.flat.data_ptr()  # first slice of the flat buffer: no copy anywhere
        ref_w = sum(torch.full((3, 4), float(r + 1)) * 2 for r in range(world)) / world
        for step in range(2):  # second step: zero_() + accumulate-in-place again
            reducer.zero_()
            x = torch.full((2, 4), float(rank + 1))
            lin(x).sum().backward()
            assert lin.weight.grad is reducer.views[0], "autograd must accumulate into the view, not replace it"
            detached = reducer()
            assert detached == 0 and reducer.never_reached == 1
            assert torch.allclose(lin.weight.grad, ref_w)
            assert torch.allclose(lin.bias.grad, torch.full((3,), 2.0))
            assert unused.grad is reducer.views[2] and float(unused.grad.abs().max()) == 0.0
        # a gradient on one rank only: every rank still receives the average (replicas cannot drift apart)
        reducer.zero_()
        if rank == 0:
            (unused * 3.0).sum().backward()
        reducer()
        assert torch.allclose(unused.grad, torch.full((5,), 3.0 / world))
        # somebody dropped the views (optimizer.zero_grad(set_to_none=True)): copied in and re-attached
        for p in list(lin.parameters()) + [unused]:
            p.grad = None
        lin(torch.full((2, 4), float(rank + 1))).sum().backward()
        assert reducer() == 3  # two fresh autograd tensors and one None: copied in / zero-filled, then re-attached
        assert torch.allclose(lin.weight.grad, ref_w) and lin.weight.grad is reducer.views[0]
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
