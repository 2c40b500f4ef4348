"""patchmatchnet_b200 -- B200-native (sm_100a) learned-PatchMatch hot path of PatchmatchNet.

Public surface (mirrors the reference's ``models`` package for this path):

    PatchMatch      drop-in for reference models/patchmatch.py:PatchMatch
    PatchmatchNet   caller-side shell with the reference constructor/forward (models/net.py)
    ops             tensor-level wrappers over the C ABI (include/patchmatch_b200.h)
"""
from . import distributed, engine, ops, synthetic  # noqa: F401
from .net import PatchmatchNet, load_reference_state, patchmatchnet_loss  # noqa: F401
from .patchmatch import PatchMatch  # noqa: F401

__all__ = ["PatchMatch", "PatchmatchNet", "load_reference_state", "patchmatchnet_loss"]
