"""Depth / confidence map files (PFM, COLMAP .bin) -- host-side mirror of the map functions of the reference's
``datasets/data_io.py`` over the native readers / writers of libpmb200.so
(``csrc/pm_mapio.cpp``, C ABI ``pmb200_map_{probe,read,write}``).

Same names, arguments, return values, exception texts and -- byte for byte -- the same files:

    read_map  data_io.py:128-145      save_map  data_io.py:148-162
    read_bin  data_io.py:165-191      save_bin  data_io.py:194-223
    read_pfm  data_io.py:226-288      save_pfm  data_io.py:291-322

plus the two calls the B200 path itself uses so that a map never passes through a pageable temporary:

    read_map_to_device(path, device)   file -> pinned host buffer -> one async copy to the GPU
    save_map_from_device(path, tensor) GPU -> pinned host buffer -> file (row flip / planar order folded into the write)

There is no Python fallback: without libpmb200.so every call raises ``NativeLibraryMissing``.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch

from . import _native

PFM = 1  # PMB200_MAP_PFM
COLMAP_BIN = 2  # PMB200_MAP_COLMAP_BIN
_UNSUPPORTED = "Invalid input format; only pfm and bin are supported"  # data_io.py:145, :162
_BAD_SHAPE = "Image must have H x W x 3, H x W x 1 or H x W dimensions."


def _raise(rc: int) -> None:
    msg = _native.lib().pmb200_last_error().decode("utf-8", "replace")
    if rc == -3:  # PMB200_EIO: "<strerror>: '<path>'"
        if msg.startswith("No such file or directory"):
            raise FileNotFoundError(2, "No such file or directory", msg.split(": ", 1)[1].strip("'"))
        raise OSError(msg)
    if rc == -4 and (msg.startswith("cannot reshape") or msg.startswith("could not convert")):
        raise ValueError(msg)  # what np.reshape / float() raise in the reference
    raise Exception(msg)  # the reference raises bare Exception("Not a PFM file.") etc.


def _format_of(path: str) -> int:
    if path.endswith(".bin"):
        return COLMAP_BIN
    if path.endswith(".pfm"):
        return PFM
    raise Exception(_UNSUPPORTED)


def _probe(path: str, fmt: int) -> "_native.MapInfo":
    info = _native.MapInfo()
    rc = _native.lib().pmb200_map_probe(os.fsencode(path), fmt, info)
    if rc != 0:
        _raise(rc)
    return info


def _read_into(path: str, fmt: int, pinned: bool) -> Tuple[torch.Tensor, "_native.MapInfo"]:
    info = _probe(path, fmt)
    n = info.width * info.height * info.channels
    buf = torch.empty((info.height, info.width, info.channels), dtype=torch.float32, pin_memory=pinned)
    rc = _native.lib().pmb200_map_read(os.fsencode(path), fmt, buf.data_ptr(), n, info)
    if rc != 0:
        _raise(rc)
    return buf, info


def read_pfm(filename: str) -> Tuple[np.ndarray, float]:
    """data_io.py:226-288 -> (array [H,W,1] ('Pf') or [H,W,3] ('PF'), float32, top row first; scale)."""
    buf, info = _read_into(filename, PFM, pinned=False)
    return buf.numpy(), float(info.scale)


def read_bin(path: str) -> np.ndarray:
    """data_io.py:165-191 -> array [H,W,C] float32."""
    return _read_into(path, COLMAP_BIN, pinned=False)[0].numpy()


def scale_to_max_dim(image: np.ndarray, max_dim: int) -> Tuple[np.ndarray, int, int]:
    """data_io.py:13-31 (cv2.resize is the reference's own library call; only taken when max_dim shrinks the map)."""
    original_height, original_width = image.shape[0], image.shape[1]
    scale = max_dim / max(original_height, original_width)
    if 0 < scale < 1:
        import cv2

        image = cv2.resize(image, (int(scale * original_width), int(scale * original_height)), interpolation=cv2.INTER_LINEAR)
    return image, original_height, original_width


def read_map(path: str, max_dim: int = -1) -> np.ndarray:
    """data_io.py:128-145."""
    if path.endswith(".bin"):
        in_map = read_bin(path)
    elif path.endswith(".pfm"):
        in_map, _ = read_pfm(path)
    else:
        raise Exception(_UNSUPPORTED)
    return scale_to_max_dim(in_map, max_dim)[0]


def _hwc(data: np.ndarray) -> Tuple[int, int, int]:
    if data.ndim == 2:
        return data.shape[0], data.shape[1], 1
    if data.ndim == 3 and data.shape[2] in (1, 3):
        return data.shape[0], data.shape[1], data.shape[2]
    raise Exception(_BAD_SHAPE)


def _write(path: str, fmt: int, data: np.ndarray, scale: float) -> None:
    h, w, c = _hwc(data)
    arr = np.ascontiguousarray(data)
    rc = _native.lib().pmb200_map_write(os.fsencode(path), fmt, arr.ctypes.data, h, w, c, scale)
    if rc != 0:
        _raise(rc)


def save_pfm(filename: str, image: np.ndarray, scale: float = 1) -> None:
    """data_io.py:291-322.  (The reference opens the file before it validates, leaving an empty file behind a failed
    call; here nothing is created unless the arguments are valid.)"""
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.dtype.byteorder == ">":
        raise Exception("big-endian arrays are not supported on the B200 host path")
    # the library prints "%f" of -scale; keep Python's own sign arithmetic for the written value (int 0 stays "0.000000")
    _write(filename, PFM, image, -float(-scale))


def save_bin(filename: str, data: np.ndarray) -> None:
    """data_io.py:194-223."""
    if data.dtype != np.float32:
        raise Exception("Image data type must be float32.")
    _write(filename, COLMAP_BIN, data, 1.0)


def save_map(path: str, data: np.ndarray) -> None:
    """data_io.py:148-162."""
    if path.endswith(".bin"):
        save_bin(path, data)
    elif path.endswith(".pfm"):
        save_pfm(path, data)
    else:
        raise Exception(_UNSUPPORTED)


def read_map_to_device(path: str, device, non_blocking: bool = True) -> torch.Tensor:
    """File -> pinned host buffer -> device, [H,W,C] float32.  The copy is enqueued on the current stream of `device`."""
    buf, _ = _read_into(path, _format_of(path), pinned=True)
    return buf.to(device, non_blocking=non_blocking)


def save_map_from_device(path: str, tensor: torch.Tensor, scale: float = 1) -> None:
    """Device map ([H,W], [H,W,1] or [H,W,3], float32) -> pinned host buffer -> file."""
    fmt = _format_of(path)
    if tensor.dtype != torch.float32:
        raise Exception("Image dtype must be float32." if fmt == PFM else "Image data type must be float32.")
    if tensor.dim() not in (2, 3) or (tensor.dim() == 3 and tensor.shape[2] not in (1, 3)):
        raise Exception(_BAD_SHAPE)
    src = tensor.detach().contiguous()
    host = torch.empty(src.shape, dtype=torch.float32, pin_memory=src.is_cuda)
    host.copy_(src)  # synchronous with respect to the host: the data is there when copy_ returns
    h, w = host.shape[0], host.shape[1]
    c = host.shape[2] if host.dim() == 3 else 1
    rc = _native.lib().pmb200_map_write(os.fsencode(path), fmt, host.data_ptr(), h, w, c, -float(-scale) if fmt == PFM else 1.0)
    if rc != 0:
        _raise(rc)


def save_ply(path: str, body) -> None:
    """Write the reference's fused.ply (eval.py:283-296: PlyData([PlyElement.describe(vertex_all, "vertex")]).write -- binary
    little endian, one `vertex` element with float x, y, z and uchar red, green, blue) from vertex records produced by
    ops.fuse_points: uint8 [n,15] tensor(s) (a list concatenates the per-reference-view clouds, eval.py:283-284)."""
    import torch

    parts = list(body) if isinstance(body, (list, tuple)) else [body]
    parts = [b.detach().cpu().contiguous() for b in parts]
    for b in parts:
        if b.dtype != torch.uint8 or b.dim() != 2 or b.shape[1] != 15:
            raise ValueError("save_ply: vertex records must be uint8 [n,15] tensors (ops.fuse_points)")
    n = sum(int(b.shape[0]) for b in parts)
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % n)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        for b in parts:
            f.write(b.numpy().tobytes())
