"""Tensor-level wrappers over the C ABI (``include/patchmatch_b200.h``).

PyTorch is plumbing here: it owns the device buffers and the CUDA stream.  Each
wrapper checks dtype / device / layout, allocates the output with the caching
allocator, and enqueues ONE native kernel on the tensor's current stream via
ctypes -- capturable in a CUDA graph, no host synchronisation.  A CPU tensor is
an error (there is no fallback path).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch

from . import _native

Tensor = torch.Tensor


def _on_device(t: Tensor) -> bool:
    """The one place that decides whether a tensor may be handed to the native library.  (tests/emu_backend.py replaces
    this, `_device_guard` and `_stream` to run the wrappers against the CPU-emulated kernels; the product has no such path.)"""
    return t.is_cuda


def _device_guard(t: Tensor):
    """Context that makes `t`'s GPU current for the duration of a native call."""
    return torch.cuda.device(t.device)


def _require(t: Tensor, name: str, ndim: Optional[int] = None) -> Tensor:
    if not _on_device(t):
        raise RuntimeError(f"{name}: the B200 path needs a CUDA tensor (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def _offsets_arg(off: Tensor, shape, name: str):
    """(tensor, channels_last flag) for a raw offset-conv output: planar NCHW memory or channels-last memory are both
    consumed in place; anything else is made NCHW-contiguous."""
    if not _on_device(off) or off.dtype != torch.float32:
        raise RuntimeError(f"{name}: offsets must be a CUDA float32 tensor (got {off.dtype} on {off.device}); there is no CPU fallback")
    if tuple(off.shape) != tuple(shape):
        raise RuntimeError(f"{name}: offsets must be {tuple(shape)}, got {tuple(off.shape)}")
    if off.is_contiguous():
        return off, 0
    if off.is_contiguous(memory_format=torch.channels_last):
        return off, 1
    return off.contiguous(), 0


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _rowmajor_4x4(m: Tensor, name: str) -> Tensor:
    """[B,4,4] with contiguous 4x4 blocks; the batch stride may be anything (torch.unbind views)."""
    if not _on_device(m) or m.dtype != torch.float32 or m.dim() != 3 or m.shape[1:] != (4, 4):
        raise RuntimeError(f"{name}: expected a CUDA float32 [B,4,4] tensor, got {m.dtype} {tuple(m.shape)} on {m.device}")
    if m.stride(1) != 4 or m.stride(2) != 1:
        m = m.contiguous()
    return m


def set_tuning(key: str, value: int = 0) -> None:
    """Measurement aid (tools/kbench.py, A/B tests): override a launch-configuration knob of the library for this process;
    `set_tuning("reset")` restores every default.  Keys: include/patchmatch_b200.h, pmb200_set_tuning.  Results never depend
    on these knobs, only launch shapes do."""
    _native.check(_native.lib().pmb200_set_tuning(key.encode(), int(value)), "set_tuning")


def relative_projection(ref_proj: Tensor, src_projs: Sequence[Tensor]) -> Tensor:
    """[V,B,12] rotation/translation of src_proj @ inv(ref_proj); reference module.py:148-150."""
    B = ref_proj.shape[0]
    V = len(src_projs)
    ref = _rowmajor_4x4(ref_proj, "ref_proj")
    mats = [_rowmajor_4x4(m, "src_proj") for m in src_projs]
    if any(m.shape[0] != B for m in mats):
        raise RuntimeError("relative_projection: batch size mismatch")
    if B > 1 and len({m.stride(0) for m in mats}) != 1:
        mats = [m.contiguous() for m in mats]
    out = torch.empty((V, B, 12), dtype=torch.float32, device=ref.device)
    arr, keep = _native.pointer_array([m.data_ptr() for m in mats])
    with _device_guard(ref):
        rc = _native.lib().pmb200_relative_projection(
            ref.data_ptr(), ref.stride(0), arr, mats[0].stride(0), V, B, out.data_ptr(), _stream(ref)
        )
    _native.check(rc, "relative_projection")
    return out


def _is_packed_nhwc(maps: Sequence[Tensor]) -> bool:
    """True when the maps are channels-last views laid out back to back in ONE storage (e.g. slices of one
    stacked channels-last FeatureNet output), so that packing would be a no-op."""
    first = maps[0]
    B, C, H, W = first.shape
    want = (H * W * C, 1, W * C, C)
    store = first.untyped_storage().data_ptr()
    if first.dtype != torch.float32 or first.data_ptr() % 32 != 0:  # the kernels read a lane's 8 channels with one 256-bit load
        return False
    for i, m in enumerate(maps):
        if (m.shape != first.shape or m.stride() != want or m.untyped_storage().data_ptr() != store
                or m.storage_offset() != first.storage_offset() + i * first.numel()):
            return False
    return True


def pack_nhwc(maps: Sequence[Tensor]) -> Tensor:
    """[n,B,H,W,C] channels-last copy of n NCHW maps of equal shape (one launch)."""
    first = maps[0]
    B, C, H, W = first.shape
    if _is_packed_nhwc(maps):
        return torch.as_strided(first, (len(maps), B, H, W, C), (B * H * W * C, H * W * C, W * C, C, 1))
    srcs = [_require(m, "feature map", 4) for m in maps]
    for m in srcs:
        if m.shape != first.shape:
            raise RuntimeError("pack_nhwc: maps must share one shape")
    out = torch.empty((len(srcs), B, H, W, C), dtype=torch.float32, device=first.device)
    arr, keep = _native.pointer_array([m.data_ptr() for m in srcs])
    with _device_guard(first):
        rc = _native.lib().pmb200_pack_nhwc(arr, len(srcs), B, C, H, W, out.data_ptr(), _stream(first))
    _native.check(rc, "pack_nhwc")
    return out


def photometric_confidence(prob: Tensor, out_h: int, out_w: int) -> Tensor:
    """[B,D,h,w] probabilities -> [B,out_h,out_w] confidence (reference net.py:289-299)."""
    pr = _require(prob, "prob", 4)
    B, D, h, w = pr.shape
    out = torch.empty((B, out_h, out_w), dtype=torch.float32, device=pr.device)
    with _device_guard(pr):
        rc = _native.lib().pmb200_photometric_confidence(pr.data_ptr(), out.data_ptr(), B, D, h, w, out_h, out_w, _stream(pr))
    _native.check(rc, "photometric_confidence")
    return out


def upsample2x_add(x: Tensor, y: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """bilinear x2 upsample of `x` plus `y` (plus a per-channel `bias`), channels-last 4-D CUDA tensors (logical NCHW)."""
    N, C, h, w = x.shape
    if y.shape != (N, C, 2 * h, 2 * w):
        raise RuntimeError("upsample2x_add: y must be [N,C,2h,2w]")
    for t, nm in ((x, "x"), (y, "y")):
        if not _on_device(t) or t.dtype != torch.float32:
            raise RuntimeError(f"upsample2x_add: {nm} must be a CUDA float32 tensor (no CPU fallback)")
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    if not y.is_contiguous(memory_format=torch.channels_last):
        y = y.contiguous(memory_format=torch.channels_last)
    b_ptr = None
    if bias is not None:
        bias = _require(bias, "bias", 1)
        if bias.numel() != C:
            raise RuntimeError("upsample2x_add: bias must have C elements")
        b_ptr = bias.data_ptr()
    out = torch.empty_like(y, memory_format=torch.channels_last)
    with _device_guard(x):
        rc = _native.lib().pmb200_upsample2x_add_nhwc(x.data_ptr(), y.data_ptr(), b_ptr, out.data_ptr(), N, h, w, C, _stream(x))
    _native.check(rc, "upsample2x_add_nhwc")
    return out


# The small learned convs (offset convs of the hot path, FeatureNet, Refinement) run through csrc/pm_conv.cu in eval
# mode.  False (or PMB200_NATIVE_CONVS=0) hands them back to cuDNN -- kept only for A/B measurements.
NATIVE_CONVS = os.environ.get("PMB200_NATIVE_CONVS", "1") != "0"
NATIVE_CONVS_ALL = os.environ.get("PMB200_NATIVE_CONVS", "1") == "all"  # A/B aid: native even where the library is faster


def conv_prefers_native(cin: int, cout: int, ks: int) -> bool:
    """Which implementation runs a small conv in eval mode.  Measured per layer on B200 (profiles/r1_run17_convbench.json,
    tools/convbench.py): the native channels-last kernel wins on the memory-bound layers -- every 1x1 conv and every
    layer with at most 3200 multiply-adds per output pixel (full-resolution 3->8 / 8->8 / 8->16 convs, the refinement
    head, the stage-1 offset conv: 1.2-3.7x faster than the library) -- while the FLOP-heavy 16..64-channel 3x3 / 5x5
    layers stay with cuDNN, whose tcgen05 implicit-GEMM kernels reach 100-150 TFLOP/s there against ~50 for the
    legacy mma.sync path this kernel issues (measured peak of that path: ~245 TFLOP/s TF32)."""
    if not NATIVE_CONVS:
        return False
    # fp32-accurate mode (cudnn.allow_tf32 off): the library's fp32 kernels for the FLOP-bound layers do not use the tensor
    # cores at all (value_fp32 of profiles/r2_run1_bench.json: 2.62 ms per forward against 0.755), the native 3xTF32 kernel
    # does -> every conv native.  The parity configuration and the bench default are this mode (the depth maps of the TF32
    # mode are 2.6e-3 .. 3.3e-3 off the fp32 reference, outside north_star's 1e-3: tests/test_gpu_bench_mode.py).
    return NATIVE_CONVS_ALL or conv_precision() == 3 or ks == 1 or cin * cout * ks * ks <= 3200


def _round_kcin(cin: int) -> int:
    return 8 if cin <= 8 else (16 if cin <= 16 else (32 if cin <= 32 else 64))


def _round_nt(cout: int) -> int:
    nt = (cout + 7) // 8
    return nt if nt <= 4 else 8


def _tf32_round(x: Tensor) -> Tensor:
    """fp32 -> nearest TF32 value (10-bit mantissa), ties away from zero -- the arithmetic of PTX cvt.rna.tf32.f32 for
    finite inputs: add half an ulp to the magnitude bits, clear the 13 low bits."""
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_conv_filter(weight: Tensor, precision: int, transposed: bool = False) -> Tensor:
    """Conv filter [Cout,Cin,KS,KS] -> the tensor-core fragment order `pmb200_conv2d_nhwc` reads for `precision`
    (include/patchmatch_b200.h): [tap][k-slice][n-tile][lane][2 or 4] with lane = 4*g + t holding the weights of
    output channel nt*8+g for the adjacent input channels ks*8+2t, ks*8+2t+1 of that tap (the MMA's k slots t, t+4),
    zero padded.  precision 1: values rounded to TF32; precision 3: (b0_hi, b1_hi, b0_lo, b1_lo) with
    hi = tf32(w), lo = tf32(w - hi).  `transposed`: `weight` is a ConvTranspose2d filter [Cin,Cout,KS,KS]; the
    equivalent direct filter is its spatial flip with the channel axes swapped.  Pure layout/rounding work on the
    weight's own device (host logic, CPU-testable)."""
    if weight.dim() != 4 or weight.shape[2] != weight.shape[3]:
        raise RuntimeError(f"pack_conv_filter: expected [Cout,Cin,KS,KS], got {tuple(weight.shape)}")
    if precision not in (1, 3):
        raise RuntimeError("pack_conv_filter: precision must be 1 (TF32) or 3 (3xTF32)")
    w = weight.detach().float()
    if transposed:
        w = w.flip(2, 3).permute(1, 0, 2, 3)
    cout, cin, ks, _ = w.shape
    if not (1 <= cin <= 64 and 1 <= cout <= 64):
        raise RuntimeError("pack_conv_filter: 1 <= Cin, Cout <= 64")
    kc, nt = _round_kcin(cin), _round_nt(cout)
    wp = torch.zeros((ks * ks, kc, nt * 8), dtype=torch.float32, device=w.device)
    wp[:, :cin, :cout] = w.permute(2, 3, 1, 0).reshape(ks * ks, cin, cout)
    # channel = ks8*8 + 2*t + half ; n = j*8 + g  ->  [tap][ks8][j][g][t][half]
    frag = wp.view(ks * ks, kc // 8, 4, 2, nt, 8).permute(0, 1, 4, 5, 2, 3).contiguous()
    hi = _tf32_round(frag)
    if precision == 1:
        return hi.view(-1)
    lo = _tf32_round(frag - hi)
    return torch.cat((hi, lo), dim=-1).contiguous().view(-1)


# K-D5 (csrc/pm_conv5.cu): tcgen05 / TMEM / TMA implicit-GEMM convolution, fp32-accurate (3xTF32).  Used in the
# fp32-accurate mode for the layers it serves when TC5_CONVS is on (PMB200_TC5=0 hands them back to the mma.sync kernel).
TC5_CONVS = os.environ.get("PMB200_TC5", "1") != "0"


# (Cin, Cout, KS, stride) of the layers a tcgen05 form beats the mma.sync kernel's 3xTF32 mode on (cold us per launch at the
# 640x512 sizes, tools/convbench.py, profiles/r2_run16_convbench.json).  Per-tap form K-D5 (the only one for stride 2):
# conv5 70 vs 93, conv8 50 vs 63; it loses on conv2 (137 vs 91: Cin = 8, one K slice per tap).  PMB200_TC5=all takes every
# supported layer (A/B measurements), PMB200_TC5=0 none.
TC5_LAYERS = {(16, 32, 5, 2), (32, 64, 5, 2)}
TC5_ALL = os.environ.get("PMB200_TC5", "1") == "all"
# Stride-1 layers served by the halo-tile form K-D5h (one TMA box + one split per output tile, taps as shifted descriptors,
# [w_hi | w_lo] as one operand): mma.sync 3xTF32 -> K-D5h, cold us: conv3/4 58 -> 36, conv6/7 67 -> 29, conv9/10 47 -> 29,
# output2 27 -> 25, output3 77 -> 44, stage-3 offset convs 39 / 40 -> 17 / 18, stage-2 25 / 27 -> 18 / 21, stage-1 32 -> 30.
# It loses where Cin = 8 (conv1 93 vs 70: one K slice per tap, the instruction count is the cost) and ties on Refinement.
# PMB200_TC5H=all: every stride-1 layer the kernel takes (A/B measurements); =0: none.
TC5H_LAYERS = {(16, 16, 3, 1), (32, 32, 3, 1), (64, 64, 3, 1), (64, 32, 1, 1), (64, 16, 1, 1),
               (64, 32, 3, 1), (64, 18, 3, 1), (32, 16, 3, 1), (32, 18, 3, 1), (16, 18, 3, 1)}
TC5H_ALL = os.environ.get("PMB200_TC5H", "1") == "all"
TC5H_OFF = os.environ.get("PMB200_TC5H", "1") == "0"


def conv_tc5_mode(cin: int, cout: int, ks: int, stride: int = 1, transposed: bool = False, fused_add: bool = False) -> str:
    """"" (mma.sync kernel), "halo" (K-D5h) or "tap" (K-D5): which tcgen05 form, if any, runs this conv.  Only in the
    fp32-accurate mode (the arithmetic of both IS the 3xTF32 split), never with a fused transposed / upsample-add epilogue, and
    only for layers on the measured lists above."""
    if not (NATIVE_CONVS and TC5_CONVS) or transposed or fused_add or conv_precision() != 3:
        return ""
    if not _native.lib().pmb200_conv2d_tc5_supported(cin, cout, ks, stride):
        return ""
    key = (cin, cout, ks, stride)
    if stride == 1 and not TC5H_OFF and (TC5H_ALL or key in TC5H_LAYERS):
        return "halo"
    if TC5_ALL or key in TC5_LAYERS:
        return "tap"
    return ""


def conv_uses_tc5(cin: int, cout: int, ks: int, stride: int = 1, transposed: bool = False, fused_add: bool = False) -> bool:
    return conv_tc5_mode(cin, cout, ks, stride, transposed, fused_add) != ""


def pack_conv_filter_tc5_for(weight: Tensor, mode: str) -> Tensor:
    return pack_conv_filter_tc5h(weight) if mode == "halo" else pack_conv_filter_tc5(weight)


def pack_conv_filter_tc5(weight: Tensor) -> Tensor:
    """Conv filter [Cout,Cin,KS,KS] -> the shared-memory image `pmb200_conv2d_tc5` copies per tap (include/patchmatch_b200.h):
    [tap][Cin/32 blocks][2 Npad rows][min(Cin,32) channels], rows 0..Npad-1 = w_hi (TF32-rounded weight), Npad..2Npad-1 = w_lo
    (TF32-rounded remainder) -- ONE K-major tcgen05 operand of N = 2 Npad whose first Npad rows double as the N = Npad operand --
    in the swizzle of the tensor core for its row size.  Pure layout / rounding work on the weight's device (CPU-testable)."""
    if weight.dim() != 4 or weight.shape[2] != weight.shape[3]:
        raise RuntimeError(f"pack_conv_filter_tc5: expected [Cout,Cin,KS,KS], got {tuple(weight.shape)}")
    w = weight.detach().float()
    cout, cin, ks, _ = w.shape
    if cin not in (8, 16, 32, 64) or not (1 <= cout <= 64):
        raise RuntimeError("pack_conv_filter_tc5: Cin in {8,16,32,64}, 1 <= Cout <= 64")
    npad = (cout + 15) // 16 * 16
    cblk = min(cin, 32)
    kb = cin // cblk
    hi = _tf32_round(w)
    lo = _tf32_round(w - hi)
    both = torch.stack((hi, lo))  # [2, Cout, Cin, KS, KS]
    t = torch.zeros((ks * ks, kb, 2, npad, cblk), dtype=torch.float32, device=w.device)
    t[:, :, :, :cout] = both.permute(3, 4, 0, 1, 2).reshape(ks * ks, 2, cout, kb, cblk).permute(0, 3, 1, 2, 4)
    # swizzle: 16-byte chunk j of row r moves to chunk j ^ f(r); Npad is a multiple of 8, so f(Npad + r) = f(r)
    chunks = cblk // 4
    rows = torch.arange(npad, device=w.device)
    f = {8: rows % 8, 4: (rows // 2) % 4, 2: (rows // 4) % 2}[chunks]
    j = torch.arange(chunks, device=w.device)
    src = (j.view(1, chunks) ^ f.view(npad, 1))  # physical chunk p of row r holds logical chunk p ^ f(r)
    t = t.view(ks * ks, kb, 2, npad, chunks, 4)
    idx = src.view(1, 1, 1, npad, chunks, 1).expand(ks * ks, kb, 2, npad, chunks, 4)
    return torch.gather(t, 4, idx).contiguous().view(-1)


def pack_conv_filter_tc5h(weight: Tensor) -> Tensor:
    """Conv filter [Cout,Cin,KS,KS] -> the image `pmb200_conv2d_tc5h` streams per tap: [tap][Cin/4 chunks][2 Npad rows][4 floats],
    rows 0..Npad-1 = w_hi, Npad..2Npad-1 = w_lo: ONE tcgen05 operand of N = 2 Npad (a_hi x [w_hi | w_lo] is a single MMA), whose
    first Npad rows double as the N = Npad operand of a_lo x w_hi.  No swizzle: 8 rows of one chunk are one core matrix."""
    if weight.dim() != 4 or weight.shape[2] != weight.shape[3]:
        raise RuntimeError(f"pack_conv_filter_tc5h: expected [Cout,Cin,KS,KS], got {tuple(weight.shape)}")
    w = weight.detach().float()
    cout, cin, ks, _ = w.shape
    if cin not in (8, 16, 32, 64) or not (1 <= cout <= 64):
        raise RuntimeError("pack_conv_filter_tc5h: Cin in {8,16,32,64}, 1 <= Cout <= 64")
    npad = (cout + 15) // 16 * 16
    hi = _tf32_round(w)
    lo = _tf32_round(w - hi)
    both = torch.stack((hi, lo))  # [2, Cout, Cin, KS, KS]
    t = torch.zeros((ks * ks, cin // 4, 2, npad, 4), dtype=torch.float32, device=w.device)
    t[:, :, :, :cout] = both.permute(3, 4, 0, 1, 2).reshape(ks * ks, 2, cout, cin // 4, 4).permute(0, 3, 1, 2, 4)
    return t.contiguous().view(-1)


# K-S (csrc/pm_stem.cu): FeatureNet's conv0 -> conv1 as one exact-fp32 launch; PMB200_STEM=0 keeps the two K-D launches.
STEM_FUSED = os.environ.get("PMB200_STEM", "1") != "0"


def conv_stem(x: Tensor, w0: Tensor, b0: Tensor, w1: Tensor, b1: Tensor) -> Tensor:
    """relu(conv1(relu(conv0(x)))) of FeatureNet (3 -> 8 -> 8, 3x3, pad 1; reference models/net.py:18-19, :44) in one launch.
    `x` [N,3,H,W] contiguous NCHW on the device; the BatchNorm-folded weights `w0` [8,3,3,3], `b0` [8], `w1` [8,8,3,3], `b1` [8]
    are HOST float32 tensors (they travel in the kernel's parameter block).  Returns [N,8,H,W] in channels-last memory."""
    if not _on_device(x) or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError(f"conv_stem: x must be a CUDA float32 [N,3,H,W] tensor (got {x.dtype} {tuple(x.shape)} on {x.device}); no CPU fallback")
    if not x.is_contiguous():
        x = x.contiguous()
    for name, t, shape in (("w0", w0, (8, 3, 3, 3)), ("b0", b0, (8,)), ("w1", w1, (8, 8, 3, 3)), ("b1", b1, (8,))):
        if t.device.type != "cpu" or t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous():
            raise RuntimeError(f"conv_stem: {name} must be a contiguous float32 HOST tensor of shape {shape}")
    N, _, H, W = x.shape
    y = torch.empty((N, 8, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    with _device_guard(x):
        rc = _native.lib().pmb200_conv_stem(x.data_ptr(), w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr(), y.data_ptr(), N, H, W,
                                            _stream(x))
    _native.check(rc, "conv_stem")
    return y


# K-R (csrc/pm_refine.cu + the one-plane form of csrc/pm_stem.cu): Refinement as two exact-fp32 launches; PMB200_REFINE=0 keeps
# the six conv-family launches and the ATen tail.
REFINE_FUSED = os.environ.get("PMB200_REFINE", "1") != "0"


def _host_weights(name: str, pairs) -> None:
    for nm, t, shape in pairs:
        if t.device.type != "cpu" or t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous():
            raise RuntimeError(f"{name}: {nm} must be a contiguous float32 HOST tensor of shape {shape}")


def refine_low(depth_half: Tensor, depth_min: Tensor, depth_max: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> Tensor:
    """relu(conv2(relu(conv1((depth - lo) / span)))) of Refinement at half resolution (reference models/net.py:104-110), one launch.
    depth_half [N,1,h,w], depth_min / depth_max [N] on the device; BatchNorm-folded HOST weights.  -> [N,8,h,w], channels-last."""
    if not _on_device(depth_half) or depth_half.dtype != torch.float32 or depth_half.dim() != 4 or depth_half.shape[1] != 1:
        raise RuntimeError(f"refine_low: depth_half must be a CUDA float32 [N,1,h,w] tensor (got {depth_half.dtype} {tuple(depth_half.shape)} on {depth_half.device}); no CPU fallback")
    _host_weights("refine_low", (("w1", w1, (8, 1, 3, 3)), ("b1", b1, (8,)), ("w2", w2, (8, 8, 3, 3)), ("b2", b2, (8,))))
    N, _, h, w = depth_half.shape
    dh = depth_half if depth_half.is_contiguous() else depth_half.contiguous()
    lo, hi = _require(depth_min.reshape(-1), "depth_min", 1), _require(depth_max.reshape(-1), "depth_max", 1)
    if lo.numel() != N or hi.numel() != N:
        raise RuntimeError("refine_low: depth_min / depth_max must have one value per image")
    y = torch.empty((N, 8, h, w), dtype=torch.float32, device=dh.device, memory_format=torch.channels_last)
    with _device_guard(dh):
        rc = _native.lib().pmb200_refine_low(dh.data_ptr(), lo.data_ptr(), hi.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                             y.data_ptr(), N, h, w, _stream(dh))
    _native.check(rc, "refine_low")
    return y


def refine_full(low: Tensor, img: Tensor, depth_half: Tensor, depth_min: Tensor, depth_max: Tensor, wd: Tensor, bd: Tensor, w0: Tensor, b0: Tensor,
                w3: Tensor, b3: Tensor, wr: Tensor) -> Tensor:
    """The full-resolution half of Refinement (reference models/net.py:103, :112-120) in one launch: transposed conv of `low`,
    conv0 of the image, conv3 over their concatenation, the residual conv, `(nearest_up2x(d) + res) * span + lo`.
    low [N,8,h,w] channels-last (refine_low), img [N,3,2h,2w] NCHW, depth_half [N,1,h,w]; HOST weights.  -> depth [N,1,2h,2w]."""
    if not _on_device(low) or low.dtype != torch.float32 or low.dim() != 4 or low.shape[1] != 8 or not low.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError("refine_full: low must be a channels-last CUDA float32 [N,8,h,w] tensor; no CPU fallback")
    N, _, h, w = low.shape
    if not _on_device(img) or img.dtype != torch.float32 or tuple(img.shape) != (N, 3, 2 * h, 2 * w):
        raise RuntimeError(f"refine_full: img must be a CUDA float32 [{N},3,{2 * h},{2 * w}] tensor, got {tuple(img.shape)}")
    if tuple(depth_half.shape) != (N, 1, h, w) or depth_half.dtype != torch.float32 or not _on_device(depth_half):
        raise RuntimeError(f"refine_full: depth_half must be a CUDA float32 [{N},1,{h},{w}] tensor")
    _host_weights("refine_full", (("wd", wd, (8, 8, 3, 3)), ("bd", bd, (8,)), ("w0", w0, (8, 3, 3, 3)), ("b0", b0, (8,)),
                                  ("w3", w3, (8, 16, 3, 3)), ("b3", b3, (8,)), ("wr", wr, (1, 8, 3, 3))))
    img = img if img.is_contiguous() else img.contiguous()
    dh = depth_half if depth_half.is_contiguous() else depth_half.contiguous()
    lo, hi = _require(depth_min.reshape(-1), "depth_min", 1), _require(depth_max.reshape(-1), "depth_max", 1)
    out = torch.empty((N, 1, 2 * h, 2 * w), dtype=torch.float32, device=low.device)
    with _device_guard(low):
        rc = _native.lib().pmb200_refine_full(low.data_ptr(), img.data_ptr(), dh.data_ptr(), lo.data_ptr(), hi.data_ptr(), wd.data_ptr(), bd.data_ptr(),
                                              w0.data_ptr(), b0.data_ptr(), w3.data_ptr(), b3.data_ptr(), wr.data_ptr(), out.data_ptr(), N, 2 * h, 2 * w,
                                              _stream(low))
    _native.check(rc, "refine_full")
    return out


def conv2d_tc5(x: Tensor, filter_tc5: Tensor, bias: Optional[Tensor], cout: int, ks: int, stride: int = 1, pad: int = 0, dil: int = 1,
               relu: bool = False, out: Optional[Tensor] = None, out_channel_offset: int = 0, halo: bool = False) -> Tensor:
    """Channels-last convolution on the 5th-generation tensor cores (csrc/pm_conv5.cu), fp32-accurate.  Same calling
    convention as conv2d_nhwc; `filter_tc5` comes from pack_conv_filter_tc5, or from pack_conv_filter_tc5h with `halo=True`
    (stride 1 only: the halo-tile form K-D5h)."""
    if halo and stride != 1:
        raise RuntimeError("conv2d_tc5: the halo-tile form serves stride 1")
    if not _on_device(x) or x.dtype != torch.float32 or x.dim() != 4:
        raise RuntimeError(f"conv2d_tc5: x must be a 4-D CUDA float32 tensor (got {x.dtype} {tuple(x.shape)} on {x.device}); no CPU fallback")
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    N, cin, H, W = x.shape
    want = _native.lib().pmb200_conv2d_tc5_filter_floats(cin, cout, ks)
    if want <= 0 or filter_tc5.numel() != want or filter_tc5.dtype != torch.float32 or filter_tc5.device != x.device:
        raise RuntimeError(f"conv2d_tc5: filter must be {want} float32 values on {x.device} (pack_conv_filter_tc5)")
    Ho = (H + 2 * pad - dil * (ks - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (ks - 1) - 1) // stride + 1
    if out is None:
        out = torch.empty((N, cout, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        ycs, yco = cout, 0
    else:
        if (out.dim() != 4 or out.shape[0] != N or out.shape[2:] != (Ho, Wo) or out.dtype != torch.float32 or out.device != x.device
                or not out.is_contiguous(memory_format=torch.channels_last)):
            raise RuntimeError("conv2d_tc5: `out` must be a channels-last CUDA float32 [N,C,Ho,Wo] tensor")
        ycs, yco = out.shape[1], out_channel_offset
    b_ptr = None
    if bias is not None:
        bias = _require(bias, "bias", 1)
        if bias.numel() != cout:
            raise RuntimeError("conv2d_tc5: bias must have Cout elements")
        b_ptr = bias.data_ptr()
    with _device_guard(x):
        if halo:
            rc = _native.lib().pmb200_conv2d_tc5h(x.data_ptr(), filter_tc5.data_ptr(), b_ptr, out.data_ptr(), N, H, W, cin, cout, ks, pad,
                                                  dil, 1 if relu else 0, ycs, yco, _stream(x))
        else:
            rc = _native.lib().pmb200_conv2d_tc5(x.data_ptr(), filter_tc5.data_ptr(), b_ptr, out.data_ptr(), N, H, W, cin, cout, ks, stride,
                                                 pad, dil, 1 if relu else 0, ycs, yco, _stream(x))
    _native.check(rc, "conv2d_tc5h" if halo else "conv2d_tc5")
    return out


def conv_precision() -> int:
    """1 (TF32 operands) when the library would use TF32 for convolutions (torch.backends.cudnn.allow_tf32, torch's
    default), else 3 (3xTF32, fp32-accurate): the native convs honour the same switch as the cuDNN ones they replace."""
    return 1 if torch.backends.cudnn.allow_tf32 else 3


def conv2d_nhwc(x: Tensor, filter_frag: Tensor, bias: Optional[Tensor], cout: int, ks: int, stride: int = 1, pad: int = 0,
                dil: int = 1, relu: bool = False, transposed2x: bool = False, out: Optional[Tensor] = None,
                out_channel_offset: int = 0, precision: Optional[int] = None, rows_per_warp: int = 0,
                add_up2x: Optional[Tensor] = None) -> Tensor:
    """Channels-last convolution on the tensor cores (csrc/pm_conv.cu).  `x` is a logical-NCHW CUDA tensor in
    channels-last memory (made so if not); returns a logical-NCHW channels-last tensor [N,cout,Ho,Wo], or writes the
    channels out_channel_offset..+cout of `out` (a wider channels-last tensor: concat fusion) and returns `out`.
    `add_up2x`: a coarser channels-last map [N,cout,Ho/2,Wo/2] whose bilinear x2 upsample is added in the epilogue."""
    if not _on_device(x) or x.dtype != torch.float32 or x.dim() != 4:
        raise RuntimeError(f"conv2d_nhwc: x must be a 4-D CUDA float32 tensor (got {x.dtype} {tuple(x.shape)} on {x.device}); no CPU fallback")
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    N, cin, H, W = x.shape
    prec = conv_precision() if precision is None else precision
    up_ptr = None
    if add_up2x is not None:
        if (not _on_device(add_up2x) or add_up2x.dtype != torch.float32 or add_up2x.dim() != 4
                or not add_up2x.is_contiguous(memory_format=torch.channels_last)):
            raise RuntimeError("conv2d_nhwc: add_up2x must be a channels-last CUDA float32 tensor")
        up_ptr = add_up2x.data_ptr()
    want = _native.lib().pmb200_conv2d_filter_floats(cin, cout, ks, prec)
    if want <= 0 or filter_frag.numel() != want or filter_frag.dtype != torch.float32 or filter_frag.device != x.device:
        raise RuntimeError(f"conv2d_nhwc: filter must be {want} float32 values on {x.device} in fragment order for precision {prec} "
                           "(pack_conv_filter)")
    Hv, Wv = (2 * H, 2 * W) if transposed2x else (H, W)
    Ho = (Hv + 2 * pad - dil * (ks - 1) - 1) // stride + 1
    Wo = (Wv + 2 * pad - dil * (ks - 1) - 1) // stride + 1
    if add_up2x is not None and tuple(add_up2x.shape) != (N, cout, Ho // 2, Wo // 2):
        raise RuntimeError(f"conv2d_nhwc: add_up2x must be [N,cout,Ho/2,Wo/2] = {(N, cout, Ho // 2, Wo // 2)}, got {tuple(add_up2x.shape)}")
    if out is None:
        out = torch.empty((N, cout, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        ycs, yco = cout, 0
    else:
        if (out.dim() != 4 or out.shape[0] != N or out.shape[2:] != (Ho, Wo) or out.dtype != torch.float32 or out.device != x.device
                or not out.is_contiguous(memory_format=torch.channels_last)):
            raise RuntimeError("conv2d_nhwc: `out` must be a channels-last CUDA float32 [N,C,Ho,Wo] tensor")
        ycs, yco = out.shape[1], out_channel_offset
    b_ptr = None
    if bias is not None:
        bias = _require(bias, "bias", 1)
        if bias.numel() != cout:
            raise RuntimeError("conv2d_nhwc: bias must have Cout elements")
        b_ptr = bias.data_ptr()
    with _device_guard(x):
        rc = _native.lib().pmb200_conv2d_nhwc(
            x.data_ptr(), filter_frag.data_ptr(), b_ptr, up_ptr, out.data_ptr(), N, H, W, cin, cout, ks, stride, pad, dil,
            1 if relu else 0, prec, 1 if transposed2x else 0, ycs, yco, rows_per_warp, _stream(x),
        )
    _native.check(rc, "conv2d_nhwc")
    return out


def warp_corr(
    ref_nhwc: Tensor, src_nhwc: Tensor, rt: Tensor, depth: Tensor, G: int, view_weights: Optional[Tensor] = None
) -> Tensor:
    """K-A.  ref [B,H,W,C], src [V,B,Hs,Ws,C], rt [V,B,12], depth [B,D,H,W].

    Without view weights: per-view similarities [V,B,G,D,H,W].  With view_weights [B,V,H,W]:
    the weighted average over views [B,G,D,H,W] (reference patchmatch.py:192-217)."""
    ref = _require(ref_nhwc, "ref_nhwc", 4)
    src = _require(src_nhwc, "src_nhwc", 5)
    rt = _require(rt, "rt", 3)
    depth = _require(depth, "depth", 4)
    B, H, W, C = ref.shape
    V, Bs, Hs, Ws, Cs = src.shape
    D = depth.shape[1]
    if Bs != B or Cs != C or rt.shape != (V, B, 12) or depth.shape != (B, D, H, W):
        raise RuntimeError("warp_corr: inconsistent shapes")
    if view_weights is not None:
        vw = _require(view_weights, "view_weights", 4)
        if vw.shape != (B, V, H, W):
            raise RuntimeError("warp_corr: view_weights must be [B,V,H,W]")
        out = torch.empty((B, G, D, H, W), dtype=torch.float32, device=ref.device)
        vw_ptr = vw.data_ptr()
    else:
        out = torch.empty((V, B, G, D, H, W), dtype=torch.float32, device=ref.device)
        vw_ptr = None
    with _device_guard(ref):
        rc = _native.lib().pmb200_warp_corr(
            ref.data_ptr(), src.data_ptr(), rt.data_ptr(), depth.data_ptr(), vw_ptr, out.data_ptr(),
            V, B, C, G, H, W, Hs, Ws, D, _stream(ref),
        )
    _native.check(rc, "warp_corr")
    return out


def _warp_args(ref_nhwc, src_nhwc, rt, depth):
    ref = _require(ref_nhwc, "ref_nhwc", 4)
    src = _require(src_nhwc, "src_nhwc", 5)
    rt = _require(rt, "rt", 3)
    depth = _require(depth, "depth", 4)
    B, H, W, C = ref.shape
    V, Bs, Hs, Ws, Cs = src.shape
    D = depth.shape[1]
    if Bs != B or Cs != C or rt.shape != (V, B, 12) or depth.shape != (B, D, H, W):
        raise RuntimeError("warp_corr: inconsistent shapes")
    return ref, src, rt, depth, (V, B, C, H, W, Hs, Ws, D)


def _score_target(xs: Optional[Tensor], B: int, D: int, H: int, W: int, device):
    if xs is None:
        out = torch.empty((B, D, H, W), dtype=torch.float32, device=device)
        return out, out.data_ptr(), 1
    if xs.shape != (B, D, H, W, 2) or not xs.is_contiguous() or xs.dtype != torch.float32 or xs.device != device:
        raise RuntimeError("xs must be a contiguous float32 [B,D,H,W,2] buffer on the same device")
    return xs, xs.data_ptr() + 4, 2  # the .y lanes


def warp_corr_score(ref_nhwc: Tensor, src_nhwc: Tensor, rt: Tensor, depth: Tensor, G: int, view_weights: Tensor,
                    head: "_native.MlpStruct", xs: Optional[Tensor] = None) -> Tensor:
    """K-A with the SimilarityNet head fused (eval mode): -> raw score [B,D,H,W]; the similarity
    tensor is never written.  `head` holds the BN-folded weights (host struct, see PointwiseHead.folded)."""
    ref, src, rt, depth, (V, B, C, H, W, Hs, Ws, D) = _warp_args(ref_nhwc, src_nhwc, rt, depth)
    vw = _require(view_weights, "view_weights", 4)
    if vw.shape != (B, V, H, W):
        raise RuntimeError("warp_corr_score: view_weights must be [B,V,H,W]")
    out, out_ptr, stride = _score_target(xs, B, D, H, W, ref.device)
    with _device_guard(ref):
        rc = _native.lib().pmb200_warp_corr_score(
            ref.data_ptr(), src.data_ptr(), rt.data_ptr(), depth.data_ptr(), vw.data_ptr(), head, out_ptr, stride,
            V, B, C, G, H, W, Hs, Ws, D, _stream(ref),
        )
    _native.check(rc, "warp_corr_score")
    return out


def warp_corr_view_weights(ref_nhwc: Tensor, src_nhwc: Tensor, rt: Tensor, depth: Tensor, G: int,
                           head: "_native.MlpStruct", keep_sims: bool = False):
    """K-A with PixelwiseNet fused (eval mode): -> pixel-wise view weights [B,V,H,W]
    (and, keep_sims, the per-view similarities [V,B,G,D,H,W] for aggregate_views_score)."""
    ref, src, rt, depth, (V, B, C, H, W, Hs, Ws, D) = _warp_args(ref_nhwc, src_nhwc, rt, depth)
    out = torch.empty((B, V, H, W), dtype=torch.float32, device=ref.device)
    sims = torch.empty((V, B, G, D, H, W), dtype=torch.float32, device=ref.device) if keep_sims else None
    with _device_guard(ref):
        rc = _native.lib().pmb200_warp_corr_view_weights(
            ref.data_ptr(), src.data_ptr(), rt.data_ptr(), depth.data_ptr(), head, out.data_ptr(),
            None if sims is None else sims.data_ptr(), V, B, C, G, H, W, Hs, Ws, D, _stream(ref),
        )
    _native.check(rc, "warp_corr_view_weights")
    return (out, sims) if keep_sims else out


def aggregate_views_score(sims: Tensor, view_weights: Tensor, head: "_native.MlpStruct", xs: Optional[Tensor] = None) -> Tensor:
    """MLP(sum_v sims[v]*w[:,v] / (1e-5 + sum_v w[:,v])) -> raw score [B,D,H,W]."""
    sims = _require(sims, "sims", 6)
    vw = _require(view_weights, "view_weights", 4)
    V, B, G, D, H, W = sims.shape
    if vw.shape != (B, V, H, W):
        raise RuntimeError("aggregate_views_score: view_weights must be [B,V,H,W]")
    out, out_ptr, stride = _score_target(xs, B, D, H, W, sims.device)
    with _device_guard(sims):
        rc = _native.lib().pmb200_aggregate_views_score(
            sims.data_ptr(), vw.data_ptr(), head, out_ptr, stride, V, B, G, D, H, W, _stream(sims)
        )
    _native.check(rc, "aggregate_views_score")
    return out


def offset_corr_weight(ref_nhwc: Tensor, offsets: Tensor, G: int, K: int, dilation: int, head: "_native.MlpStruct") -> Tensor:
    """K-A' with the FeatureWeightNet head fused (eval mode): -> feature weight [B,K,H,W]."""
    ref = _require(ref_nhwc, "ref_nhwc", 4)
    B, H, W, C = ref.shape
    off, off_cl = _offsets_arg(offsets, (B, 2 * K, H, W), "offset_corr_weight")
    out = torch.empty((B, K, H, W), dtype=torch.float32, device=ref.device)
    with _device_guard(ref):
        rc = _native.lib().pmb200_offset_corr_weight(
            ref.data_ptr(), off.data_ptr(), off_cl, head, out.data_ptr(), B, C, G, H, W, K, dilation, _stream(ref)
        )
    _native.check(rc, "offset_corr_weight")
    return out


FUSED_HEAD_SHAPES = ((64, 8), (32, 8), (16, 4))


def aggregate_views(sims: Tensor, view_weights: Tensor) -> Tensor:
    """sum_v sims[v]*w[:,v] / (1e-5 + sum_v w[:,v]) -> [B,G,D,H,W]."""
    sims = _require(sims, "sims", 6)
    vw = _require(view_weights, "view_weights", 4)
    V, B, G, D, H, W = sims.shape
    if vw.shape != (B, V, H, W):
        raise RuntimeError("aggregate_views: view_weights must be [B,V,H,W]")
    out = torch.empty((B, G, D, H, W), dtype=torch.float32, device=sims.device)
    with _device_guard(sims):
        rc = _native.lib().pmb200_aggregate_views(
            sims.data_ptr(), vw.data_ptr(), out.data_ptr(), V, B, G, D, H, W, _stream(sims)
        )
    _native.check(rc, "aggregate_views")
    return out


def offset_corr(ref_nhwc: Tensor, offsets: Tensor, G: int, K: int, dilation: int) -> Tensor:
    """K-A'.  ref [B,H,W,C], offsets [B,2K,H,W] -> [B,G,K,H,W]."""
    ref = _require(ref_nhwc, "ref_nhwc", 4)
    B, H, W, C = ref.shape
    off, off_cl = _offsets_arg(offsets, (B, 2 * K, H, W), "offset_corr")
    out = torch.empty((B, G, K, H, W), dtype=torch.float32, device=ref.device)
    with _device_guard(ref):
        rc = _native.lib().pmb200_offset_corr(
            ref.data_ptr(), off.data_ptr(), off_cl, out.data_ptr(), B, C, G, H, W, K, dilation, _stream(ref)
        )
    _native.check(rc, "offset_corr")
    return out


MODE_RANDOM, MODE_PERTURB, MODE_PASSTHROUGH = 0, 1, 2


def alloc_xs(B: int, D: int, H: int, W: int, device) -> Tensor:
    """Interleaved (normalised inverse depth, raw score) buffer [B,D,H,W,2]: K-C fills the .x lanes, the K-A score
    epilogue the .y lanes, K-B then gathers both with one 8-byte load per tap."""
    return torch.empty((B, D, H, W, 2), dtype=torch.float32, device=device)


def init_propagate(
    seed_map: Tensor,
    offsets: Optional[Tensor],
    depth_min: Tensor,
    depth_max: Tensor,
    mode: int,
    Ns: int,
    Kp: int,
    dilation: int,
    interval_scale: float,
    with_xnorm: bool = False,
    xs: Optional[Tensor] = None,
):
    """K-C.  seed_map: U[0,1) noise [B,48,H,W] (mode 0) or current depth [B,1,H,W]; -> [B,Ns+Kp,H,W]
    (and, with_xnorm, the normalised inverse depth of every hypothesis, same shape).  With `xs` (an interleaved
    (xnorm, score) buffer [B,Ns+Kp,H,W,2] from alloc_xs) the normalised inverse depth goes into its .x lanes."""
    seed = _require(seed_map, "seed_map", 4)
    B, S, H, W = seed.shape
    if S != (48 if mode == MODE_RANDOM else 1):
        raise RuntimeError("init_propagate: seed_map has the wrong number of channels")
    dmin = _require(depth_min.reshape(-1), "depth_min", 1)
    dmax = _require(depth_max.reshape(-1), "depth_max", 1)
    if dmin.numel() != B or dmax.numel() != B:
        raise RuntimeError("init_propagate: depth_min/max must have B elements")
    off_ptr, off_cl = None, 0
    if Kp > 0:
        off, off_cl = _offsets_arg(offsets, (B, 2 * Kp, H, W), "init_propagate")
        off_ptr = off.data_ptr()
    out = torch.empty((B, Ns + Kp, H, W), dtype=torch.float32, device=seed.device)
    if xs is not None:
        if xs.shape != (B, Ns + Kp, H, W, 2) or not xs.is_contiguous() or xs.dtype != torch.float32 or not _on_device(xs):
            raise RuntimeError("init_propagate: xs must be a contiguous CUDA float32 [B,Ns+Kp,H,W,2] buffer")
        xn, xn_ptr, xstride = None, xs.data_ptr(), 2
    else:
        xn = torch.empty_like(out) if with_xnorm else None
        xn_ptr, xstride = (None if xn is None else xn.data_ptr()), 1
    with _device_guard(seed):
        rc = _native.lib().pmb200_init_propagate(
            seed.data_ptr(), off_ptr, off_cl, dmin.data_ptr(), dmax.data_ptr(), out.data_ptr(), xn_ptr, xstride,
            mode, B, H, W, Ns, Kp, dilation, float(interval_scale), _stream(seed),
        )
    _native.check(rc, "init_propagate")
    if xs is not None:
        return out
    return (out, xn) if with_xnorm else out


def adaptive_eval(
    score0: Tensor,
    depth_sample: Tensor,
    offsets: Tensor,
    feature_weight: Tensor,
    depth_min: Tensor,
    depth_max: Tensor,
    dilation: int,
    interval_scale: float,
    is_inverse: bool,
    xnorm: Optional[Tensor] = None,
    xs: Optional[Tensor] = None,
):
    """K-B.  -> (depth [B,H,W], prob [B,D,H,W]).  xnorm: normalised inverse depth from init_propagate.
    With `xs` (interleaved (xnorm, score) buffer) `score0` is ignored and may be None."""
    if xs is not None:
        if xs.dim() != 5 or xs.shape[-1] != 2 or not xs.is_contiguous() or not _on_device(xs) or xs.dtype != torch.float32:
            raise RuntimeError("adaptive_eval: xs must be a contiguous CUDA float32 [B,D,H,W,2] buffer")
        score0 = xs[..., 1]  # shape carrier only; never made contiguous below
        sc = score0
    else:
        sc = _require(score0, "score0", 4)
    ds = _require(depth_sample, "depth_sample", 4)
    fw = _require(feature_weight, "feature_weight", 4)
    B, D, H, W = sc.shape
    K = fw.shape[1]
    off, off_cl = _offsets_arg(offsets, (B, 2 * K, H, W), "adaptive_eval")
    if ds.shape != sc.shape or fw.shape != (B, K, H, W):
        raise RuntimeError("adaptive_eval: inconsistent shapes")
    dmin = _require(depth_min.reshape(-1), "depth_min", 1)
    dmax = _require(depth_max.reshape(-1), "depth_max", 1)
    xn_ptr = None
    if xnorm is not None:
        xn = _require(xnorm, "xnorm", 4)
        if xn.shape != sc.shape:
            raise RuntimeError("adaptive_eval: xnorm must match depth_sample")
        xn_ptr = xn.data_ptr()
    prob = torch.empty((B, D, H, W), dtype=torch.float32, device=sc.device)
    depth = torch.empty((B, H, W), dtype=torch.float32, device=sc.device)
    with _device_guard(sc):
        rc = _native.lib().pmb200_adaptive_eval(
            None if xs is not None else sc.data_ptr(), ds.data_ptr(), xn_ptr, None if xs is None else xs.data_ptr(), off.data_ptr(), off_cl, fw.data_ptr(), dmin.data_ptr(), dmax.data_ptr(),
            prob.data_ptr(), depth.data_ptr(), B, D, H, W, K, dilation, float(interval_scale),
            1 if is_inverse else 0, _stream(sc),
        )
    _native.check(rc, "adaptive_eval")
    return depth, prob


def compose_filter_cameras(ref_K, ref_E, src_Ks, src_Es) -> Tensor:
    """[V,60] float64 (CPU): the camera matrices `pmb200_geometric_filter` takes, composed the way the reference composes
    them -- numpy float32 inverses and products (eval.py:116-139) -- and then widened to float64 without rounding."""
    import numpy as np

    f32 = lambda m: np.asarray(m.detach().cpu().numpy() if isinstance(m, torch.Tensor) else m, dtype=np.float32)
    Kr, Er = f32(ref_K), f32(ref_E)
    rows = []
    for K, E in zip(src_Ks, src_Es):
        Ks, Es = f32(K), f32(E)
        t1 = np.matmul(Es, np.linalg.inv(Er))
        t2 = np.matmul(Er, np.linalg.inv(Es))
        rows.append(np.concatenate([np.linalg.inv(Kr).ravel(), t1[:3, :4].ravel(), Ks.ravel(), np.linalg.inv(Ks).ravel(),
                                    t2[:3, :4].ravel(), Kr.ravel()]).astype(np.float64))
    return torch.from_numpy(np.stack(rows))


def geometric_filter(ref_depth: Tensor, confidence: Tensor, src_depths: Tensor, cams: Tensor, geo_pixel_thres: float = 1.0,
                     geo_depth_thres: float = 0.01, photo_thres: float = 0.8, geo_mask_thres: int = 3):
    """Geometric-consistency filtering of one reference view against V source views in one launch (reference
    eval.py:86-190, :220-256).  ref_depth / confidence [H,W], src_depths [V,Hs,Ws] CUDA float32; cams [V,60] float64 from
    compose_filter_cameras (moved to the device here).  -> (photo_mask bool [H,W], geo_mask_sum int32 [H,W],
    final_mask bool [H,W], depth_averaged float64 [H,W])."""
    ref = _require(ref_depth, "ref_depth", 2)
    conf = _require(confidence, "confidence", 2)
    src = _require(src_depths, "src_depths", 3)
    H, W = ref.shape
    V, Hs, Ws = src.shape
    if conf.shape != (H, W):
        raise RuntimeError("geometric_filter: confidence must match ref_depth")
    if cams.shape != (V, 60) or cams.dtype != torch.float64:
        raise RuntimeError("geometric_filter: cams must be a [V,60] float64 tensor (compose_filter_cameras)")
    cams = cams.to(ref.device).contiguous()
    mask_sum = torch.empty((H, W), dtype=torch.int32, device=ref.device)
    photo = torch.empty((H, W), dtype=torch.uint8, device=ref.device)
    final = torch.empty((H, W), dtype=torch.uint8, device=ref.device)
    avg = torch.empty((H, W), dtype=torch.float64, device=ref.device)
    with _device_guard(ref):
        rc = _native.lib().pmb200_geometric_filter(
            ref.data_ptr(), conf.data_ptr(), src.data_ptr(), cams.data_ptr(), V, H, W, Hs, Ws, float(geo_pixel_thres),
            float(geo_depth_thres), float(photo_thres), int(geo_mask_thres), mask_sum.data_ptr(), photo.data_ptr(),
            final.data_ptr(), avg.data_ptr(), _stream(ref),
        )
    _native.check(rc, "geometric_filter")
    return photo.bool(), mask_sum, final.bool(), avg


def compose_fusion_camera(ref_K, ref_E) -> Tensor:
    """[25] float64 (CPU): inverse(ref_intrinsics) (9) then inverse(ref_extrinsics) (16), inverted in float32 as numpy does
    on the reference's float32 camera files (eval.py:278-279) and then widened without rounding."""
    import numpy as np

    f32 = lambda m: np.asarray(m.detach().cpu().numpy() if isinstance(m, torch.Tensor) else m, dtype=np.float32)
    return torch.from_numpy(np.concatenate([np.linalg.inv(f32(ref_K)).ravel(), np.linalg.inv(f32(ref_E)).ravel()]).astype(np.float64))


def fuse_points(final_mask: Tensor, depth_averaged: Tensor, ref_img: Tensor, cam25: Tensor) -> Tensor:
    """Point-cloud half of the fusion for one reference view (reference eval.py:273-296): -> uint8 [n,15] CUDA tensor, the
    binary little-endian PLY vertex records (float32 x, y, z, uint8 r, g, b) of the surviving pixels in row-major order.
    final_mask bool / uint8 [H,W], depth_averaged float64 [H,W] (as geometric_filter returns them), ref_img float32
    [H,W,3] in [0,1].  One host synchronisation at the end (the vertex count sizes the result)."""
    if not _on_device(final_mask) or not _on_device(depth_averaged) or not _on_device(ref_img):
        raise RuntimeError("fuse_points: the B200 path needs CUDA tensors; there is no CPU fallback")
    H, W = depth_averaged.shape
    if depth_averaged.dtype != torch.float64 or tuple(final_mask.shape) != (H, W) or tuple(ref_img.shape) != (H, W, 3) or ref_img.dtype != torch.float32:
        raise RuntimeError("fuse_points: final_mask [H,W], depth_averaged float64 [H,W], ref_img float32 [H,W,3]")
    if cam25.shape != (25,) or cam25.dtype != torch.float64:
        raise RuntimeError("fuse_points: cam25 must be a [25] float64 tensor (compose_fusion_camera)")
    mask = final_mask.to(torch.uint8).contiguous()
    depth = depth_averaged.contiguous()
    img = ref_img.contiguous()
    cam = cam25.to(depth.device).contiguous()
    body = torch.empty((H * W, 15), dtype=torch.uint8, device=depth.device)
    count = torch.zeros(1, dtype=torch.int32, device=depth.device)
    scratch = torch.empty(((H * W + 255) // 256,), dtype=torch.int32, device=depth.device)
    with _device_guard(depth):
        rc = _native.lib().pmb200_fuse_points(mask.data_ptr(), depth.data_ptr(), img.data_ptr(), cam.data_ptr(), H, W, body.data_ptr(),
                                              count.data_ptr(), scratch.data_ptr(), _stream(depth))
    _native.check(rc, "fuse_points")
    return body[: int(count.item())]


def split_ply_body(body: Tensor):
    """uint8 [n,15] vertex records -> (vertices float32 [n,3], colours uint8 [n,3]) on the CPU."""
    b = body.detach().cpu().contiguous()
    return b[:, :12].contiguous().view(torch.float32).view(-1, 3), b[:, 12:].contiguous()
