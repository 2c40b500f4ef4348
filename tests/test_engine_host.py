"""Host-side logic of the serving engine that needs no GPU: the packed layout of a request's small inputs (cameras + depth
range in one flat buffer whose named entries are views) and the adjacency test that decides between one image upload and one
per view."""
import torch

from patchmatchnet_b200.engine import DepthEngine
from patchmatchnet_b200.net import _adjacent_views


def _engine(B, N, H, W):
    eng = DepthEngine.__new__(DepthEngine)  # no CUDA: only the shape-dependent host helpers are exercised
    eng.shape = (B, N, H, W)
    return eng


def test_packed_request_parameters_land_in_the_named_views():
    B, N = 2, 5
    eng = _engine(B, N, 16, 24)
    g = torch.Generator().manual_seed(0)
    req = dict(intrinsics=torch.randn(B, N, 3, 3, generator=g), extrinsics=torch.randn(B, N, 4, 4, generator=g),
               depth_min=torch.tensor([425.0, 400.0], dtype=torch.float64), depth_max=torch.tensor([935.0, 900.0]))
    blk = eng._input_block("cpu")
    assert blk["params"].numel() == B * N * 25 + 2 * B
    assert float(blk["depth_min"].min()) == 1.0 and float(blk["depth_max"].max()) == 2.0  # valid range before the first request
    pack = torch.full((blk["params"].numel(),), float("nan"))
    nbytes = eng._pack_params(pack, req)
    assert nbytes == 4 * pack.numel() and torch.isfinite(pack).all()
    blk["params"].copy_(pack)  # what the one upload + one device copy do
    assert torch.equal(blk["intrinsics"], req["intrinsics"]) and torch.equal(blk["extrinsics"], req["extrinsics"])
    assert torch.equal(blk["depth_min"], req["depth_min"].float()) and torch.equal(blk["depth_max"], req["depth_max"])
    for k in ("intrinsics", "extrinsics", "depth_min", "depth_max"):  # views of the one buffer, not copies
        assert blk[k].untyped_storage().data_ptr() == blk["params"].untyped_storage().data_ptr()
    assert blk["images"].shape == (N, B, 3, 16, 24)


def test_adjacent_views_decides_the_single_image_upload():
    stack = torch.arange(5 * 2 * 3 * 4 * 6, dtype=torch.float32).view(5, 2, 3, 4, 6)
    views = [stack[i] for i in range(5)]
    assert _adjacent_views(views)
    first = views[0]
    stacked = torch.as_strided(first, (5,) + tuple(first.shape), (first.numel(),) + tuple(first.stride()))
    assert torch.equal(stacked, stack)  # the view infer_stream uploads in one copy
    assert not _adjacent_views([v.clone() for v in views])
    assert not _adjacent_views([views[0], views[2], views[1], views[3], views[4]])
    assert not _adjacent_views([stack[i].transpose(-1, -2) for i in range(5)])
