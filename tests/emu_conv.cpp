// emu_conv.cpp -- the channels-last tensor-core conv family (pm_conv.cu) compiled for the HOST on top of tests/warp_emu.h:
// the kernel AND the launch planning of pmb200_conv2d_nhwc run as they are (the CUDA runtime calls of the planner are
// stubbed in warp_emu.h, mma.sync / cp.async / cvt.rna are restated under PM_EMU).  TEST INFRASTRUCTURE, see emu_kernels.cpp.
#define PM_EMU 1
#include "warp_emu.h"

static thread_local char g_conv_err[256] = "";
extern "C" int pmb200_internal_fail(int code, const char *msg) {
    snprintf(g_conv_err, sizeof g_conv_err, "%s", msg);
    return code;
}
extern "C" const char *emu_conv_last_error(void) { return g_conv_err; }

// the real entry points, under emulation-only names (libpmb200.so exports the product ones)
#define pmb200_conv2d_nhwc emu_conv2d_nhwc
#define pmb200_conv2d_filter_floats emu_conv2d_filter_floats
#include "../patchmatchnet_b200/csrc/pm_conv.cu"

// K-S, the fused conv0 -> conv1 stem of FeatureNet (pm_stem.cu): plain fp32 FFMA, nothing to restate under PM_EMU
#define pmb200_conv_stem emu_conv_stem
#define pmb200_refine_low emu_refine_low
#include "../patchmatchnet_b200/csrc/pm_stem.cu"

// K-R: the full-resolution half of Refinement (pm_refine.cu)
#define pmb200_refine_full emu_refine_full
#include "../patchmatchnet_b200/csrc/pm_refine.cu"
