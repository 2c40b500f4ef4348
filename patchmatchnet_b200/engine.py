"""Inference engine: the public call a user makes to get depth maps out of host images.

``DepthEngine`` owns static device buffers for one input shape, captures the whole
``PatchmatchNet.forward`` (cuDNN feature pyramid, the native PatchMatch kernels, refinement) into ONE
CUDA graph -- the cascade is ~300 short kernels, so at 640x512 launch latency, not bandwidth, is the
enemy -- and moves data with pinned-memory async copies on a side stream so the
next request's host->device copy overlaps the current request's compute.

    eng = DepthEngine(net, batch=1, n_views=5, height=512, width=640, device="cuda:0")
    depth, confidence = eng.infer(images, intrinsics, extrinsics, depth_min, depth_max)   # host in, host out

There is no CPU path: construction fails without CUDA.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


class DepthEngine:
    def __init__(
        self,
        net: torch.nn.Module,
        batch: int,
        n_views: int,
        height: int,
        width: int,
        device: str = "cuda:0",
        use_graph: bool = True,
        warmup: int = 3,
    ) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("DepthEngine needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device(device)
        self.net = net.eval().to(self.device)
        self.shape = (batch, n_views, height, width)
        B, N, H, W = self.shape
        dev = self.device
        # static device inputs (two sets: the copy stream fills one while the graph reads the other)
        self._slots = []
        for _ in range(2):
            self._slots.append(
                dict(
                    images=torch.zeros(N, B, 3, H, W, device=dev),
                    intrinsics=torch.zeros(B, N, 3, 3, device=dev),
                    extrinsics=torch.zeros(B, N, 4, 4, device=dev),
                    depth_min=torch.ones(B, device=dev),
                    depth_max=torch.full((B,), 2.0, device=dev),
                )
            )
        self._host_out = dict(
            depth=torch.empty(B, 1, H, W).pin_memory(),
            confidence=torch.empty(B, H, W).pin_memory(),
        )
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.compute_stream = torch.cuda.Stream(device=dev)
        self.use_graph = use_graph
        self._graphs: List[Optional[torch.cuda.CUDAGraph]] = [None, None]
        self._outs: List[Optional[Tuple[Tensor, Tensor]]] = [None, None]
        self._warmup = warmup
        self._ready = False

    # ------------------------------------------------------------------
    def _forward(self, slot: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
        imgs = [slot["images"][i] for i in range(self.shape[1])]
        depth, conf, _ = self.net(imgs, slot["intrinsics"].clone(), slot["extrinsics"], slot["depth_min"], slot["depth_max"])
        return depth, conf

    def set_device_inputs(self, slot_idx: int, images, intrinsics, extrinsics, depth_min, depth_max) -> None:
        """Fill a slot from tensors that already live on the device (kernel-only timing path)."""
        s = self._slots[slot_idx]
        for i, im in enumerate(images):
            s["images"][i].copy_(im)
        s["intrinsics"].copy_(intrinsics)
        s["extrinsics"].copy_(extrinsics)
        s["depth_min"].copy_(depth_min.reshape(-1))
        s["depth_max"].copy_(depth_max.reshape(-1))

    def prepare(self) -> None:
        """Warm up (cuDNN autotune, allocator) and capture one graph per slot.  Slots must hold valid cameras."""
        with torch.no_grad():
            with torch.cuda.device(self.device), torch.cuda.stream(self.compute_stream):
                for slot in self._slots:
                    for _ in range(self._warmup):
                        self._forward(slot)
                self.compute_stream.synchronize()
                if self.use_graph:
                    for i, slot in enumerate(self._slots):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=self.compute_stream):
                            self._outs[i] = self._forward(slot)
                        self._graphs[i] = g
        torch.cuda.synchronize(self.device)
        self._ready = True

    def run_slot(self, slot_idx: int) -> Tuple[Tensor, Tensor]:
        """One forward on the inputs currently in the slot, enqueued on the CURRENT stream."""
        if not self._ready:
            self.prepare()
        if self.use_graph:
            self._graphs[slot_idx].replay()
            return self._outs[slot_idx]
        with torch.no_grad():
            return self._forward(self._slots[slot_idx])

    # ------------------------------------------------------------------
    def upload(self, slot_idx: int, host: Dict[str, object], stream: torch.cuda.Stream) -> int:
        """Async pinned-host -> device copy of one request; returns the bytes moved."""
        s = self._slots[slot_idx]
        n = 0
        with torch.cuda.stream(stream):
            for i, im in enumerate(host["images"]):
                s["images"][i].copy_(im, non_blocking=True)
                n += im.numel() * im.element_size()
            for k in ("intrinsics", "extrinsics", "depth_min", "depth_max"):
                src = host[k]
                s[k].copy_(src.reshape(s[k].shape), non_blocking=True)
                n += src.numel() * src.element_size()
        return n

    def infer(self, images: Sequence[Tensor], intrinsics: Tensor, extrinsics: Tensor, depth_min: Tensor, depth_max: Tensor):
        """Host tensors in, host (pinned) tensors out; one request, synchronous."""
        host = dict(images=list(images), intrinsics=intrinsics, extrinsics=extrinsics,
                    depth_min=depth_min.float(), depth_max=depth_max.float())
        if not self._ready:
            self.upload(0, host, self.compute_stream)
            self.upload(1, host, self.compute_stream)
            self.compute_stream.synchronize()
            self.prepare()
        with torch.cuda.device(self.device), torch.cuda.stream(self.compute_stream):
            self.upload(0, host, self.compute_stream)
            depth, conf = self.run_slot(0)
            self._host_out["depth"].copy_(depth, non_blocking=True)
            self._host_out["confidence"].copy_(conf, non_blocking=True)
        self.compute_stream.synchronize()
        return self._host_out["depth"], self._host_out["confidence"]

    def infer_stream(self, requests: Sequence[Dict[str, object]], on_result=None) -> Tuple[int, int]:
        """Pipelined serving loop over pinned-host requests: request i+1 uploads on the copy stream while
        request i computes; every result is read back to pinned host memory.  Returns (h2d_bytes, d2h_bytes)
        per request."""
        if not self._ready:
            self.upload(0, requests[0], self.compute_stream)
            self.upload(1, requests[0], self.compute_stream)
            self.compute_stream.synchronize()
            self.prepare()
        h2d = d2h = 0
        uploaded = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.device(self.device):
            h2d = self.upload(0, requests[0], self.copy_stream)
            uploaded[0].record(self.copy_stream)
            for i in range(len(requests)):
                cur, nxt = i % 2, (i + 1) % 2
                if i + 1 < len(requests):
                    if i >= 1:
                        self.copy_stream.wait_event(consumed[nxt])  # slot nxt was read by request i-1
                    self.upload(nxt, requests[i + 1], self.copy_stream)
                    uploaded[nxt].record(self.copy_stream)
                with torch.cuda.stream(self.compute_stream):
                    self.compute_stream.wait_event(uploaded[cur])
                    depth, conf = self.run_slot(cur)
                    consumed[cur].record(self.compute_stream)
                    self._host_out["depth"].copy_(depth, non_blocking=True)
                    self._host_out["confidence"].copy_(conf, non_blocking=True)
                    d2h = depth.numel() * 4 + conf.numel() * 4
                if on_result is not None:
                    self.compute_stream.synchronize()
                    on_result(i, self._host_out["depth"], self._host_out["confidence"])
            self.compute_stream.synchronize()
        return h2d, d2h
