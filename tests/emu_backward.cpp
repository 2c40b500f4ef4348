// emu_backward.cpp -- the backward kernels (pm_backward.cu) compiled for the HOST on top of tests/warp_emu.h.
// TEST INFRASTRUCTURE, see emu_kernels.cpp.  Each function does what the launcher of the same name in pm_backward.cu does
// (parameter block, zero fill of the scatter targets, grid) and runs the kernel body under the emulator.
#define PM_EMU 1
#include "warp_emu.h"

#include "../patchmatchnet_b200/csrc/pm_backward.cu"

extern "C" {

int emu_warp_corr_backward(const float *ref_nhwc, const float *src_nhwc, const float *rt, const float *depth, const float *vw,
                           const float *grad_out, float *d_ref_nhwc, float *d_src_nhwc, int V, int B, int C, int G, int H, int W,
                           int Hs, int Ws, int D) {
    WarpCorrBwdParams p;
    p.ref = ref_nhwc; p.src = src_nhwc; p.rt = rt; p.depth = depth; p.vw = vw; p.gout = grad_out;
    p.dref = d_ref_nhwc; p.dsrc = d_src_nhwc;
    p.V = V; p.B = B; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws; p.D = D;
    p.sx = (W > 1) ? (float)(Ws - 1) / (float)(W - 1) : 1.0f;
    p.sy = (H > 1) ? (float)(Hs - 1) / (float)(H - 1) : 1.0f;
    memset(d_src_nhwc, 0, (size_t)V * B * Hs * Ws * C * sizeof(float));
    const int HW = H * W;
#define EMU_WB(CC, GG)                                                                                   \
    do {                                                                                                 \
        const int pix_per_block = 4 * BwdMap<CC, GG>::PPW;                                               \
        emu::launch(dim3((HW + pix_per_block - 1) / pix_per_block, B), dim3(128), 0, [&] { warp_corr_backward_kernel<CC, GG>(p); }); \
        return 0;                                                                                        \
    } while (0)
    if (C == 64 && G == 8) EMU_WB(64, 8);
    if (C == 32 && G == 8) EMU_WB(32, 8);
    if (C == 16 && G == 4) EMU_WB(16, 4);
#undef EMU_WB
    return -2;
}

int emu_aggregate_views_backward(const float *sims, const float *vw, const float *grad_out, float *d_sims, float *d_vw, int V, int B,
                                 int G, int D, int H, int W) {
    const int HW = H * W;
    emu::launch(dim3((HW + 127) / 128, B), dim3(128), 0, [&] { aggregate_views_backward_kernel(sims, vw, grad_out, d_sims, d_vw, V, B, G * D, HW); });
    return 0;
}

int emu_offset_corr_backward(const float *ref_nhwc, const float *offsets, const float *grad_out, float *d_offsets, int B, int C, int G,
                             int H, int W, int K, int dilation) {
    OffsetCorrBwdParams p;
    p.ref = ref_nhwc; p.offsets = offsets; p.gout = grad_out; p.doff = d_offsets;
    p.B = B; p.H = H; p.W = W; p.K = K; p.dilation = dilation;
    const int HW = H * W;
#define EMU_OB(CC, GG)                                                                                   \
    do {                                                                                                 \
        const int pix_per_block = 4 * BwdMap<CC, GG>::PPW;                                               \
        emu::launch(dim3((HW + pix_per_block - 1) / pix_per_block, B), dim3(128), 0, [&] { offset_corr_backward_kernel<CC, GG>(p); }); \
        return 0;                                                                                        \
    } while (0)
    if (C == 64 && G == 8) EMU_OB(64, 8);
    if (C == 32 && G == 8) EMU_OB(32, 8);
    if (C == 16 && G == 4) EMU_OB(16, 4);
#undef EMU_OB
    return -2;
}

int emu_init_propagate_backward(const float *seed_map, const float *offsets, const float *depth_min, const float *depth_max,
                                const float *grad_out, float *d_offsets, int mode, int B, int H, int W, int Ns, int Kp, int dilation,
                                float interval_scale) {
    PropBwdParams p;
    p.seed = seed_map; p.offsets = offsets; p.dmin = depth_min; p.dmax = depth_max; p.gout = grad_out; p.doff = d_offsets;
    p.mode = mode; p.B = B; p.H = H; p.W = W; p.Ns = Ns; p.Kp = Kp; p.dilation = dilation; p.interval_scale = interval_scale;
    emu::launch(dim3((H * W + 127) / 128, B), dim3(128), 0, [&] { init_propagate_backward_kernel(p); });
    return 0;
}

int emu_adaptive_eval_backward(const float *score0, const float *depth_sample, const float *xnorm, const float *offsets,
                               const float *feature_weight, const float *depth_min, const float *depth_max, const float *prob,
                               const float *grad_depth, const float *grad_prob, float *d_score0, float *d_depth_sample,
                               float *d_offsets, float *d_feature_weight, int B, int D, int H, int W, int K, int dilation,
                               float interval_scale, int is_inverse) {
    EvalBwdParams p;
    p.score0 = score0; p.depth = depth_sample; p.xnorm = xnorm; p.offsets = offsets; p.fw = feature_weight;
    p.dmin = depth_min; p.dmax = depth_max; p.prob = prob; p.gdepth = grad_depth; p.gprob = grad_prob;
    p.dscore0 = d_score0; p.dhyp = d_depth_sample; p.doff = d_offsets; p.dfw = d_feature_weight;
    p.B = B; p.D = D; p.H = H; p.W = W; p.K = K; p.dilation = dilation; p.is_inverse = is_inverse;
    p.interval_scale = interval_scale;
    memset(d_score0, 0, (size_t)B * D * H * W * sizeof(float));
    emu::launch(dim3((H * W + 63) / 64, B), dim3(64), 0, [&] { adaptive_eval_backward_kernel(p); });
    return 0;
}

}  // extern "C"
