// emu_kernels.cpp -- the repo's warp-level kernels compiled for the HOST on top of tests/warp_emu.h.
// TEST INFRASTRUCTURE: built by tests/conftest.py (g++, -DPM_EMU) into tests/_emu_kernels.so and driven by
// tests/test_emulated_kernels.py; never loaded by the product package.  The functions below do what the launchers in
// pm_kernels.cu do (fill the parameter block, size the grid) and then run the kernel body under the emulator.
#define PM_EMU 1
#include "warp_emu.h"

#ifndef PM_EMU_KERNELS
#define PM_EMU_KERNELS "../patchmatchnet_b200/csrc/pm_kernels.cu"
#endif
#include PM_EMU_KERNELS

namespace {

template <int C, int G, int EPI, int DC, int PIPE>
int run_wc3(const WarpCorrParams &p, const MlpParams &m, float *sims_out) {
    if constexpr ((LaneMap<C, G>::PPW * DC) % 32 != 0) {
        return -2;
    } else {
        const int HW = p.H * p.W;
        constexpr int pix_per_block = kWarps2 * LaneMap<C, G>::PPW;
        dim3 grid((HW + pix_per_block - 1) / pix_per_block, (p.D + DC - 1) / DC, p.B);
        emu::launch(grid, dim3(kWarps2 * 32), 0, [&] { warp_corr3_kernel<C, G, EPI, DC, PIPE, 1>(p, m, sims_out); });
        return 0;
    }
}

template <int C, int G, int EPI>
int run_wc3_cfg(const WarpCorrParams &p, const MlpParams &m, float *sims_out, int dc, int pipe) {
#define EMU_TRY(DD, PP) \
    if (dc == DD && pipe == PP) return run_wc3<C, G, EPI, DD, PP>(p, m, sims_out);
    EMU_TRY(4, 0) EMU_TRY(4, 1) EMU_TRY(8, 0) EMU_TRY(8, 1) EMU_TRY(16, 0) EMU_TRY(16, 1)
#undef EMU_TRY
    return -2;
}

template <int C, int G>
int run_wc3_epi(const WarpCorrParams &p, const MlpParams &m, float *sims_out, int epi, int dc, int pipe) {
    switch (epi) {
        case kEpiSims: return run_wc3_cfg<C, G, kEpiSims>(p, m, sims_out, dc, pipe);
        case kEpiAgg: return run_wc3_cfg<C, G, kEpiAgg>(p, m, sims_out, dc, pipe);
        case kEpiScore: return run_wc3_cfg<C, G, kEpiScore>(p, m, sims_out, dc, pipe);
        case kEpiViewW: return run_wc3_cfg<C, G, kEpiViewW>(p, m, sims_out, dc, pipe);
    }
    return -2;
}

// K-A, fourth generation: persistent producer / consumer pipeline (pm_warpcorr4.cuh).  The TMA engine is emulated by
// immediate copies that complete their mbarrier transactions; the producer warp and the consumer warps are fibers like any
// other thread, an mbarrier wait yields until the phase flips.
template <int C, int G, int EPI, int NW>
int run_wc4(const WarpCorrParams &p, const MlpParams &m, float *sims_out, int cap, int grid, int stages) {
    using L = wc4::Layout<C, G, NW>;
    wc4::Params4 q;
    q.p = p;
    q.ntx = (p.W + wc4::kTW - 1) / wc4::kTW;
    q.nty = (p.H + NW - 1) / NW;
    q.nd = (p.D + wc4::kDItem - 1) / wc4::kDItem;
    q.nitems = q.ntx * q.nty * q.nd * p.B;
    q.cap = cap;
    q.stages = stages < 2 ? 2 : (stages > wc4::kMaxStages ? wc4::kMaxStages : stages);
    const wc4::TensorMap tm{p.ref, p.B, p.H, p.W, C, NW};
    const size_t smem = (size_t)L::fixed_bytes + (size_t)q.stages * cap * C * 4;
    if (grid < 1 || grid > q.nitems) grid = q.nitems;
    emu::launch(dim3(grid), dim3((NW + 1) * 32), smem, [&] { wc4::warp_corr4_kernel<C, G, EPI, NW, 1>(q, m, sims_out, tm); });
    return 0;
}

template <int C, int G>
int run_wc4_epi(const WarpCorrParams &p, const MlpParams &m, float *sims_out, int epi, int nw, int cap, int grid, int stages) {
#define EMU_TRY4(EE, NN) \
    if (epi == EE && nw == NN) return run_wc4<C, G, EE, NN>(p, m, sims_out, cap, grid, stages);
    EMU_TRY4(kEpiSims, 4) EMU_TRY4(kEpiAgg, 4) EMU_TRY4(kEpiScore, 4) EMU_TRY4(kEpiViewW, 4)
    EMU_TRY4(kEpiSims, 8) EMU_TRY4(kEpiAgg, 8) EMU_TRY4(kEpiScore, 8) EMU_TRY4(kEpiViewW, 8)
#undef EMU_TRY4
    return -2;
}

}  // namespace

extern "C" {

// K-A, fourth generation.  Same epilogue numbering as emu_warp_corr3; nw = consumer warps (tile rows), cap = texels per
// window slot (small values force the global-memory fallback of the gather), grid = persistent CTAs (fewer than items
// makes every CTA walk several items through the rings), stages = depth of the window ring (2..4).
int emu_warp_corr4(const float *ref_nhwc, const float *src_nhwc, const float *rt, const float *depth, const float *vw,
                   const pmb200_mlp *head, float *out, float *sims_out, int ostride, int V, int B, int C, int G, int H, int W,
                   int Hs, int Ws, int D, int epi, int nw, int cap, int grid, int stages) {
    WarpCorrParams p;
    p.ref = ref_nhwc; p.src = src_nhwc; p.rt = rt; p.depth = depth; p.vw = vw; p.out = out;
    p.V = V; p.B = B; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws; p.D = D;
    p.sx = (W > 1) ? (float)(Ws - 1) / (float)(W - 1) : 1.0f;
    p.sy = (H > 1) ? (float)(Hs - 1) / (float)(H - 1) : 1.0f;
    p.ostride = ostride < 1 ? 1 : ostride;
    MlpParams m{};
    if (head) m = to_device_layout(head);
    if (C == 64 && G == 8) return run_wc4_epi<64, 8>(p, m, sims_out, epi, nw, cap, grid, stages);
    if (C == 32 && G == 8) return run_wc4_epi<32, 8>(p, m, sims_out, epi, nw, cap, grid, stages);
    if (C == 16 && G == 4) return run_wc4_epi<16, 4>(p, m, sims_out, epi, nw, cap, grid, stages);
    return -2;
}

// K-A, third generation.  epi: 0 per-view sims [V,B,G,D,H,W], 1 weighted average [B,G,D,H,W], 2 score [B,D,H,W]*ostride,
// 3 view weights [B,V,H,W] (out must be zero-filled; sims_out optional).  Returns -2 for a combination that is not built.
int emu_warp_corr3(const float *ref_nhwc, const float *src_nhwc, const float *rt, const float *depth, const float *vw,
                   const pmb200_mlp *head, float *out, float *sims_out, int ostride, int V, int B, int C, int G, int H, int W,
                   int Hs, int Ws, int D, int epi, int dc, int pipe) {
    WarpCorrParams p;
    p.ref = ref_nhwc; p.src = src_nhwc; p.rt = rt; p.depth = depth; p.vw = vw; p.out = out;
    p.V = V; p.B = B; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws; p.D = D;
    p.sx = (W > 1) ? (float)(Ws - 1) / (float)(W - 1) : 1.0f;
    p.sy = (H > 1) ? (float)(Hs - 1) / (float)(H - 1) : 1.0f;
    p.ostride = ostride < 1 ? 1 : ostride;
    MlpParams m{};
    if (head) m = to_device_layout(head);
    if (C == 64 && G == 8) return run_wc3_epi<64, 8>(p, m, sims_out, epi, dc, pipe);
    if (C == 32 && G == 8) return run_wc3_epi<32, 8>(p, m, sims_out, epi, dc, pipe);
    if (C == 16 && G == 4) return run_wc3_epi<16, 4>(p, m, sims_out, epi, dc, pipe);
    return -2;
}

// K-B.  Same arguments as pmb200_adaptive_eval plus the block shape (TP pixels x DY hypothesis lanes).
int emu_adaptive_eval(const float *score0, const float *depth_sample, const float *xnorm, const float *xnorm_score,
                      const float *offsets, int offsets_channels_last, const float *feature_weight, const float *depth_min,
                      const float *depth_max, float *prob_out, float *depth_out, int B, int D, int H, int W, int K, int dilation,
                      float interval_scale, int is_inverse, int TP, int DY) {
    EvalParams p;
    p.score0 = score0; p.depth = depth_sample; p.xnorm = xnorm; p.offsets = offsets; p.fw = feature_weight;
    p.xs = reinterpret_cast<const float2 *>(xnorm_score);
    p.off_nhwc = offsets_channels_last ? 1 : 0;
    p.dmin = depth_min; p.dmax = depth_max; p.prob = prob_out; p.depth_out = depth_out;
    p.B = B; p.D = D; p.H = H; p.W = W; p.K = K; p.dilation = dilation; p.is_inverse = is_inverse;
    p.interval_scale = interval_scale;
    const int HW = H * W;
    if (DY > D) DY = D;
    const size_t smem = (size_t)K * TP * (sizeof(float4) + sizeof(int) + sizeof(float)) + 2 * (size_t)D * TP * sizeof(float);
    dim3 grid((HW + TP - 1) / TP, B);
    if (K == 9) emu::launch(grid, dim3(TP, DY), smem, [&] { adaptive_eval_kernel<9>(p); });
    else if (K == 17) emu::launch(grid, dim3(TP, DY), smem, [&] { adaptive_eval_kernel<17>(p); });
    else return -2;
    return 0;
}

// K-C.  Same arguments as pmb200_init_propagate; the kernel is picked by hypothesis count exactly as the launcher does.
int emu_init_propagate(const float *seed_map, const float *offsets, int offsets_channels_last, const float *depth_min,
                       const float *depth_max, float *out, float *xnorm_out, int xnorm_stride, int mode, int B, int H, int W, int Ns,
                       int Kp, int dilation, float interval_scale) {
    PropParams p;
    p.seed = seed_map; p.offsets = offsets; p.dmin = depth_min; p.dmax = depth_max; p.out = out; p.xnorm = xnorm_out;
    p.xstride = xnorm_stride < 1 ? 1 : xnorm_stride;
    p.off_nhwc = offsets_channels_last ? 1 : 0;
    p.mode = mode; p.B = B; p.H = H; p.W = W; p.Ns = Ns; p.Kp = Kp; p.dilation = dilation;
    p.interval_scale = interval_scale;
    const int HW = H * W, D = Ns + Kp;
    dim3 grid((HW + 127) / 128, B);
    auto warp_grid = [&](int lanes_per_pixel) {
        const int pix_per_block = 4 * (32 / lanes_per_pixel);
        return dim3((HW + pix_per_block - 1) / pix_per_block, B);
    };
    if (Kp == 0 && D <= 64) emu::launch(grid, dim3(128), 0, [&] { init_only_kernel(p); });
    else if (D <= 8) emu::launch(warp_grid(8), dim3(128), 0, [&] { init_propagate_kernel<8>(p); });
    else if (D <= 16) emu::launch(warp_grid(16), dim3(128), 0, [&] { init_propagate_kernel<16>(p); });
    else if (D <= 32) emu::launch(warp_grid(32), dim3(128), 0, [&] { init_propagate_kernel<32>(p); });
    else if (D <= 64) emu::launch(warp_grid(32), dim3(128), 0, [&] { init_propagate_kernel<64>(p); });
    else emu::launch(grid, dim3(128), 0, [&] { init_propagate_generic_kernel(p); });
    return 0;
}

// K-A'.  head == NULL: correlations [B,G,K,H,W]; head != NULL: feature weights [B,K,H,W] (FeatureWeightNet MLP + sigmoid).
int emu_offset_corr(const float *ref_nhwc, const float *offsets, int offsets_channels_last, const pmb200_mlp *head, float *out,
                    int B, int C, int G, int H, int W, int K, int dilation) {
    OffsetCorrParams p;
    p.ref = ref_nhwc; p.offsets = offsets; p.out = out;
    p.B = B; p.H = H; p.W = W; p.K = K; p.dilation = dilation; p.off_nhwc = offsets_channels_last ? 1 : 0;
    MlpParams m{};
    if (head) m = to_device_layout(head);
    const int HW = H * W;
    const int nchunk = (K + kChunk - 1) / kChunk;
#define EMU_OC(CC, GG)                                                                                                      \
    do {                                                                                                                    \
        dim3 grid((HW + kWarpsPerBlock * LaneMap<CC, GG>::PPW - 1) / (kWarpsPerBlock * LaneMap<CC, GG>::PPW), nchunk, B);   \
        if (head) emu::launch(grid, dim3(kWarpsPerBlock * 32), 0, [&] { offset_corr_kernel<CC, GG, true>(p, m); });         \
        else emu::launch(grid, dim3(kWarpsPerBlock * 32), 0, [&] { offset_corr_kernel<CC, GG, false>(p, m); });             \
        return 0;                                                                                                           \
    } while (0)
    if (C == 64 && G == 8) EMU_OC(64, 8);
    if (C == 32 && G == 8) EMU_OC(32, 8);
    if (C == 16 && G == 4) EMU_OC(16, 4);
#undef EMU_OC
    return -2;
}

// ---- helper kernels ---------------------------------------------------------------------------------------------------
int emu_relative_projection(const float *ref_proj, long long ref_batch_stride, const float *const *src_projs, long long src_batch_stride,
                            int V, int B, float *rt_out) {
    ProjParams p;
    p.ref = ref_proj;
    for (int v = 0; v < V; ++v) p.src.p[v] = src_projs[v];
    p.out = rt_out; p.ref_stride = ref_batch_stride; p.src_stride = src_batch_stride; p.V = V; p.B = B;
    emu::launch(dim3((V * B + 63) / 64), dim3(64), 0, [&] { relative_projection_kernel(p); });
    return 0;
}

int emu_pack_nhwc(const float *const *maps, int n, int B, int C, int H, int W, float *out_nhwc) {
    PackParams p;
    for (int i = 0; i < n; ++i) p.maps.p[i] = maps[i];
    p.out = out_nhwc; p.n = n; p.B = B; p.C = C; p.HW = H * W;
    emu::launch(dim3((p.HW + 31) / 32, (C + 31) / 32, n * B), dim3(32, 8), 0, [&] { pack_nhwc_kernel(p); });
    return 0;
}

int emu_photometric_confidence(const float *prob, float *out, int B, int D, int h, int w, int H_out, int W_out) {
    const size_t total = (size_t)B * H_out * W_out;
    emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [&] { photometric_confidence_kernel(prob, out, B, D, h, w, H_out, W_out); });
    return 0;
}

int emu_upsample2x_add_nhwc(const float *x, const float *y, const float *bias, float *out, int N, int h, int w, int C) {
    const size_t total = (size_t)N * 2 * h * 2 * w * (C / 4);
    emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [&] {
        upsample2x_add_nhwc_kernel(reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(y), reinterpret_cast<const float4 *>(bias),
                                   reinterpret_cast<float4 *>(out), N, h, w, C / 4);
    });
    return 0;
}

int emu_aggregate_views(const float *sims, const float *vw, float *out, int V, int B, int G, int D, int H, int W) {
    const size_t total = (size_t)B * G * D * H * W;
    size_t blocks = (total + 255) / 256;
    if (blocks > 7) blocks = 7;  // fewer CTAs than elements: exercises the grid-stride loop
    emu::launch(dim3((unsigned)blocks), dim3(256), 0, [&] { aggregate_views_kernel(sims, vw, out, V, B, G * D, H * W); });
    return 0;
}

int emu_aggregate_views_score(const float *sims, const float *vw, const pmb200_mlp *head, float *score, int stride, int V, int B, int G,
                              int D, int H, int W) {
    const MlpParams m = to_device_layout(head);
    const size_t total = (size_t)B * D * H * W;
    const unsigned blocks = (unsigned)((total + 127) / 128);
    const int os = stride < 1 ? 1 : stride;
    if (G == 8) emu::launch(dim3(blocks), dim3(128), 0, [&] { aggregate_score_kernel<8>(sims, vw, score, m, V, B, D, H * W, os); });
    else if (G == 4) emu::launch(dim3(blocks), dim3(128), 0, [&] { aggregate_score_kernel<4>(sims, vw, score, m, V, B, D, H * W, os); });
    else return -2;
    return 0;
}

// generic (any C % G == 0) slow-path K-A: one thread per (batch, hypothesis, pixel)
int emu_warp_corr_generic(const float *ref_nhwc, const float *src_nhwc, const float *rt, const float *depth, const float *vw, float *out,
                          int V, int B, int C, int G, int H, int W, int Hs, int Ws, int D) {
    WarpCorrParams p;
    p.ref = ref_nhwc; p.src = src_nhwc; p.rt = rt; p.depth = depth; p.vw = vw; p.out = out;
    p.V = V; p.B = B; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws; p.D = D;
    p.sx = (W > 1) ? (float)(Ws - 1) / (float)(W - 1) : 1.0f;
    p.sy = (H > 1) ? (float)(Hs - 1) / (float)(H - 1) : 1.0f;
    const size_t total = (size_t)B * D * H * W;
    emu::launch(dim3((unsigned)((total + 127) / 128)), dim3(128), 0, [&] { warp_corr_generic_kernel(p, C, G); });
    return 0;
}

}  // extern "C"
