// pm_conv.cu -- sm_100a channels-last 2-D convolution family for the small learned convs either side of the
// PatchMatch hot path: the offset convs ON the path (propa_conv / eval_conv, reference models/patchmatch.py:288-311),
// the feature pyramid that produces the path's inputs (FeatureNet, reference models/net.py:9-70) and the refinement
// head that consumes its output (Refinement, reference models/net.py:73-122).  SURVEY.md 8(f) rows f1 / f3.
//
// Why hand-written: these layers have 1..64 channels.  The library's implicit-GEMM kernels are tiled for hundreds of
// channels (256x64x8 / 128x128x8 tiles) and ran the 23 convs of one 640x512 forward at ~28 TFLOP/s and far below HBM
// speed (profiles/r1_run12_launches.md: 50 % of the device time of a forward).  At these widths a conv is a
// memory-bound streaming op with a small dense contraction inside, so the design is:
//
//   * activations channels-last (NHWC) end to end -- the layout the fused PatchMatch kernels gather from;
//   * one CTA = 4 warps = an output tile of 16 columns x (4*MT) rows.  The input halo tile (all input channels) is
//     staged ONCE in shared memory with a pixel stride chosen so that the tensor-core fragment loads are bank-conflict
//     free for the layer's stride; zero padding, image borders and the zero-stuffing of a transposed conv are
//     resolved during staging, so the inner loop has no bounds checks;
//   * the contraction runs on the tensor cores as implicit GEMM, M = 16 output pixels of one row, N = 8 output
//     channels, K = 8 input channels of one filter tap (mma.sync.m16n8k8 TF32, fp32 accumulate).  The filter is
//     pre-arranged ON THE HOST in fragment order ([tap][k-slice][n-tile][lane] -> (b0,b1)), so a warp fetches a B
//     fragment with one coalesced 256-byte read-only load that stays L1/L2-resident for every CTA of the layer:
//     no shared memory for weights, no per-CTA weight staging;
//   * precision follows the library's contract for this op: `precision == 1` rounds both operands to TF32
//     (what cuDNN does under torch.backends.cudnn.allow_tf32 = True, torch's default), `precision == 3` is the
//     error-compensated 3xTF32 split (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi), fp32-accurate, used by the parity tests;
//   * bias + ReLU fused in the epilogue; the output may be written into a channel slice of a wider buffer
//     (concat fusion) and a transposed stride-2 conv is the same kernel over a virtually zero-stuffed input.
//
// tcgen05/TMEM is deliberately not used here: with N <= 64 and K <= 64 per tap the operands are far below the
// 128xNx8 UMMA tile economy, and the layers are bound by activation traffic, not by MMA issue.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/patchmatch_b200.h"

extern "C" int pmb200_internal_fail(int code, const char *msg);  // pm_kernels.cu: sets pmb200_last_error()

namespace {

struct ConvParams {
    const float *x;      // [N,H,W,Cin]
    const float2 *wf;    // fragment-ordered filter [taps][KCIN/8][NT][32] (b0,b1)
    const float *bias;   // [Cout] or null
    float *y;            // [N,Ho,Wo,ycs] written at channel offset yco
    int N, H, W, Cin, Ho, Wo, Cout;
    int KS, S, pad, dil, relu;
    int ps;              // shared-memory pixel stride (floats)
    int rw, rh;          // staged input region (pixels)
    int ycs, yco;        // output channel stride / offset
    int stuff;           // 1: the input is virtually zero-stuffed x2 (transposed stride-2 conv)
    int tiles_x, tiles_y;
};

__device__ __forceinline__ uint32_t to_tf32(float f) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// KCIN: input channels rounded up to 8/16/32/64 (zero padded while staging); NT: 8-wide output-channel tiles;
// MT: 16-pixel output rows per warp; PREC: 1 = TF32, 3 = 3xTF32.
template <int KCIN, int NT, int MT, int PREC>
__global__ void __launch_bounds__(128) conv_nhwc_mma_kernel(const ConvParams p) {
    extern __shared__ __align__(16) float s_in[];
    constexpr int KK = KCIN / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int n = blockIdx.y;
    const int ox0 = tx * 16, oy0 = ty * (4 * MT);
    const int ix0 = ox0 * p.S - p.pad, iy0 = oy0 * p.S - p.pad;

    // ---- stage the input halo tile: [rh][rw][ps], channels >= Cin and out-of-image pixels are zero ----
    {
        const int npix = p.rh * p.rw;
        const bool vec4 = (p.Cin % 4) == 0;
        const int Hv = p.stuff ? 2 * p.H : p.H, Wv = p.stuff ? 2 * p.W : p.W;  // virtual (zero-stuffed) extent
        constexpr int V = KCIN / 4;  // 4-channel vectors per pixel
        for (int idx = threadIdx.x; idx < npix * V; idx += 128) {
            const int pix = idx / V, v = idx - pix * V;
            const int ry = pix / p.rw, rx = pix - ry * p.rw;
            int iy = iy0 + ry, ix = ix0 + rx;
            bool inside = iy >= 0 && iy < Hv && ix >= 0 && ix < Wv;
            if (p.stuff) {
                inside = inside && ((iy | ix) & 1) == 0;
                iy >>= 1;
                ix >>= 1;
            }
            const int ch = v * 4;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inside && ch < p.Cin) {
                const float *src = p.x + (((size_t)n * p.H + iy) * p.W + ix) * p.Cin + ch;
                if (vec4) {
                    val = __ldg(reinterpret_cast<const float4 *>(src));
                } else {
                    val.x = __ldg(src);
                    if (ch + 1 < p.Cin) val.y = __ldg(src + 1);
                    if (ch + 2 < p.Cin) val.z = __ldg(src + 2);
                    if (ch + 3 < p.Cin) val.w = __ldg(src + 3);
                }
            }
            if (PREC == 1) {
                val.x = __uint_as_float(to_tf32(val.x));
                val.y = __uint_as_float(to_tf32(val.y));
                val.z = __uint_as_float(to_tf32(val.z));
                val.w = __uint_as_float(to_tf32(val.w));
            }
            float2 *dst = reinterpret_cast<float2 *>(s_in + (size_t)pix * p.ps + ch);  // ps is even: 8-byte aligned
            dst[0] = make_float2(val.x, val.y);
            dst[1] = make_float2(val.z, val.w);
        }
    }
    __syncthreads();

    float acc[MT][NT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[m][j][i] = 0.f;

    const float2 *wl = p.wf + lane;
    const int arow = g * p.S * p.ps + t;  // lane part of the A-fragment address
    const int astep8 = 8 * p.S * p.ps;    // 8 output pixels further right
    for (int ky = 0; ky < p.KS; ++ky) {
        for (int kx = 0; kx < p.KS; ++kx) {
            const int tap = ky * p.KS + kx;
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) {
                float2 b[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) b[j] = __ldg(wl + ((size_t)(tap * KK + kk) * NT + j) * 32);
                uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    bh[j][0] = to_tf32(b[j].x);
                    bh[j][1] = to_tf32(b[j].y);
                    if (PREC == 3) {
                        bl[j][0] = to_tf32(b[j].x - __uint_as_float(bh[j][0]));
                        bl[j][1] = to_tf32(b[j].y - __uint_as_float(bh[j][1]));
                    }
                }
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const int orow = warp * MT + m;
                    const float *ap = s_in + ((orow * p.S + ky * p.dil) * p.rw + kx * p.dil) * p.ps + kk * 8 + arow;
                    const float a0 = ap[0], a1 = ap[astep8], a2 = ap[4], a3 = ap[astep8 + 4];
                    uint32_t ah[4];
                    if (PREC == 1) {  // already rounded while staging
                        ah[0] = __float_as_uint(a0); ah[1] = __float_as_uint(a1);
                        ah[2] = __float_as_uint(a2); ah[3] = __float_as_uint(a3);
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_tf32(acc[m][j], ah, bh[j][0], bh[j][1]);
                    } else {
                        uint32_t al[4];
                        ah[0] = to_tf32(a0); ah[1] = to_tf32(a1); ah[2] = to_tf32(a2); ah[3] = to_tf32(a3);
                        al[0] = to_tf32(a0 - __uint_as_float(ah[0]));
                        al[1] = to_tf32(a1 - __uint_as_float(ah[1]));
                        al[2] = to_tf32(a2 - __uint_as_float(ah[2]));
                        al[3] = to_tf32(a3 - __uint_as_float(ah[3]));
#pragma unroll
                        for (int j = 0; j < NT; ++j) {  // small terms first
                            mma_tf32(acc[m][j], al, bh[j][0], bh[j][1]);
                            mma_tf32(acc[m][j], ah, bl[j][0], bl[j][1]);
                            mma_tf32(acc[m][j], ah, bh[j][0], bh[j][1]);
                        }
                    }
                }
            }
        }
    }

    // ---- epilogue: bias, ReLU, channels-last store (c0,c1 -> pixel g; c2,c3 -> pixel g+8; channels 2t, 2t+1) ----
    const bool vec2 = ((p.ycs | p.yco) & 1) == 0;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const int oy = oy0 + warp * MT + m;
        if (oy >= p.Ho) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = j * 8 + 2 * t;
            if (co >= p.Cout) continue;
            const bool pair = co + 1 < p.Cout;
            const float bi0 = p.bias ? __ldg(p.bias + co) : 0.f;
            const float bi1 = (p.bias && pair) ? __ldg(p.bias + co + 1) : 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ox = ox0 + g + 8 * h;
                if (ox >= p.Wo) continue;
                float v0 = acc[m][j][2 * h] + bi0, v1 = acc[m][j][2 * h + 1] + bi1;
                if (p.relu) {
                    v0 = fmaxf(v0, 0.f);
                    v1 = fmaxf(v1, 0.f);
                }
                float *dst = p.y + (((size_t)n * p.Ho + oy) * p.Wo + ox) * p.ycs + p.yco + co;
                if (pair && vec2) {
                    *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
                } else {
                    dst[0] = v0;
                    if (pair) dst[1] = v1;
                }
            }
        }
    }
}

int conv_fail(int code, const char *msg) { return pmb200_internal_fail(code, msg); }

// Shared-memory pixel stride (floats) >= kcin such that the 8 pixel rows x 4 channel columns of an A fragment fall
// into 32 distinct banks: (stride_in_pixels * ps) mod 32 must be an odd multiple of 4; ps even for 8-byte stores.
int pixel_stride(int kcin, int S) {
    for (int ps = kcin;; ps += 2) {
        const int r = (S * ps) % 32;
        if (r == 4 || r == 12 || r == 20 || r == 28) return ps;
    }
}

template <int KCIN, int NT, int MT, int PREC>
int launch_conv(const ConvParams &p, size_t smem, cudaStream_t st) {
    auto kern = conv_nhwc_mma_kernel<KCIN, NT, MT, PREC>;
    if (smem > 48 * 1024) {
        // opt in to > 48 KB of dynamic shared memory once per (instantiation, device); the high-water mark is only ever
        // raised, so concurrent callers (one host thread per GPU) at worst repeat an idempotent call
        static int granted[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || granted[dev] < (int)smem) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                char msg[200];
                snprintf(msg, sizeof(msg), "conv2d_nhwc: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
                return pmb200_internal_fail((int)e, msg);
            }
            if (dev >= 0 && dev < 64) granted[dev] = (int)smem;
        }
    }
    dim3 grid(p.tiles_x * p.tiles_y, p.N);
    kern<<<grid, 128, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        char msg[200];
        snprintf(msg, sizeof(msg), "conv2d_nhwc launch: %s", cudaGetErrorString(e));
        return pmb200_internal_fail((int)e, msg);
    }
    return 0;
}

template <int KCIN, int NT, int MT>
int launch_prec(const ConvParams &p, size_t smem, int prec, cudaStream_t st) {
    return prec == 1 ? launch_conv<KCIN, NT, MT, 1>(p, smem, st) : launch_conv<KCIN, NT, MT, 3>(p, smem, st);
}

template <int KCIN, int NT>
int launch_mt(const ConvParams &p, size_t smem, int mt, int prec, cudaStream_t st) {
    if constexpr (NT <= 2) {
        if (mt == 4) return launch_prec<KCIN, NT, 4>(p, smem, prec, st);
    }
    if constexpr (NT <= 4) {
        if (mt >= 2) return launch_prec<KCIN, NT, 2>(p, smem, prec, st);
    }
    return launch_prec<KCIN, NT, 1>(p, smem, prec, st);
}

template <int KCIN>
int launch_nt(const ConvParams &p, size_t smem, int nt, int mt, int prec, cudaStream_t st) {
    switch (nt) {
        case 1: return launch_mt<KCIN, 1>(p, smem, mt, prec, st);
        case 2: return launch_mt<KCIN, 2>(p, smem, mt, prec, st);
        case 3: return launch_mt<KCIN, 3>(p, smem, mt, prec, st);
        case 4: return launch_mt<KCIN, 4>(p, smem, mt, prec, st);
        default: return launch_mt<KCIN, 8>(p, smem, mt, prec, st);
    }
}

inline int round_kcin(int cin) { return cin <= 8 ? 8 : (cin <= 16 ? 16 : (cin <= 32 ? 32 : 64)); }
inline int round_nt(int cout) {
    const int nt = (cout + 7) / 8;
    return nt <= 4 ? nt : 8;
}

// rows per warp actually instantiated for this (nt, requested mt)
inline int effective_mt(int nt, int mt) {
    if (nt <= 2 && mt == 4) return 4;
    if (nt <= 4 && mt >= 2) return 2;
    return 1;
}

}  // namespace

extern "C" {

int pmb200_conv2d_filter_floats(int Cin, int Cout, int KS) {
    if (Cin < 1 || Cin > 64 || Cout < 1 || Cout > 64 || KS < 1) return PMB200_EINVAL;
    return KS * KS * (round_kcin(Cin) / 8) * round_nt(Cout) * 64;
}

int pmb200_conv2d_nhwc(const float *x, const float *filter_frag, const float *bias, float *y, int N, int H, int W,
                       int Cin, int Cout, int KS, int stride, int pad, int dil, int relu, int precision,
                       int transposed2x, int y_channel_stride, int y_channel_offset, int rows_per_warp, void *stream) {
    if (!x || !filter_frag || !y) return conv_fail(PMB200_EINVAL, "conv2d_nhwc: null pointer");
    if (N < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1 || Cin > 64 || Cout > 64)
        return conv_fail(PMB200_EINVAL, "conv2d_nhwc: sizes out of range (1 <= Cin, Cout <= 64)");
    if (KS < 1 || KS > 7 || stride < 1 || stride > 2 || dil < 1 || pad < 0)
        return conv_fail(PMB200_EINVAL, "conv2d_nhwc: kernel size 1..7, stride 1..2, dilation >= 1");
    if (precision != 1 && precision != 3) return conv_fail(PMB200_EINVAL, "conv2d_nhwc: precision must be 1 (TF32) or 3 (3xTF32)");
    if (transposed2x && stride != 1) return conv_fail(PMB200_EINVAL, "conv2d_nhwc: the zero-stuffed (transposed) form runs at stride 1");
    if ((Cin % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u)) return conv_fail(PMB200_EINVAL, "conv2d_nhwc: x must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(filter_frag) & 7u) return conv_fail(PMB200_EINVAL, "conv2d_nhwc: filter must be 8-byte aligned");
    const int Hv = transposed2x ? 2 * H : H, Wv = transposed2x ? 2 * W : W;
    const int Ho = (Hv + 2 * pad - dil * (KS - 1) - 1) / stride + 1;
    const int Wo = (Wv + 2 * pad - dil * (KS - 1) - 1) / stride + 1;
    if (Ho < 1 || Wo < 1) return conv_fail(PMB200_EINVAL, "conv2d_nhwc: empty output");
    const int ycs = y_channel_stride > 0 ? y_channel_stride : Cout;
    if (y_channel_offset < 0 || y_channel_offset + Cout > ycs) return conv_fail(PMB200_EINVAL, "conv2d_nhwc: channel slice outside the output");
    if (((ycs | y_channel_offset) & 1) == 0 && (reinterpret_cast<uintptr_t>(y) & 7u))
        return conv_fail(PMB200_EINVAL, "conv2d_nhwc: y must be 8-byte aligned");

    const int kcin = round_kcin(Cin), nt = round_nt(Cout);
    ConvParams p;
    p.x = x; p.wf = reinterpret_cast<const float2 *>(filter_frag); p.bias = bias; p.y = y;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
    p.KS = KS; p.S = stride; p.pad = pad; p.dil = dil; p.relu = relu ? 1 : 0;
    p.ps = pixel_stride(kcin, stride);
    p.ycs = ycs; p.yco = y_channel_offset; p.stuff = transposed2x ? 1 : 0;
    p.tiles_x = (Wo + 15) / 16;
    p.rw = 15 * stride + dil * (KS - 1) + 1;

    // rows per warp: as many as the accumulator budget allows while the grid still fills the 148 SMs a few times
    // over and the halo tile fits in shared memory
    int mt = rows_per_warp > 0 ? rows_per_warp : 4;
    for (;;) {
        mt = effective_mt(nt, mt);
        const int rows = 4 * mt;
        p.tiles_y = (Ho + rows - 1) / rows;
        p.rh = (rows - 1) * stride + dil * (KS - 1) + 1;
        const size_t smem = (size_t)p.rh * p.rw * p.ps * sizeof(float);
        const long ctas = (long)p.tiles_x * p.tiles_y * N;
        const bool fits = smem <= 200 * 1024;
        const bool enough = rows_per_warp > 0 || ctas >= 4 * 148 || mt == 1;
        const bool roomy = rows_per_warp > 0 || smem <= 56 * 1024 || mt == 1;  // >= 4 CTAs per SM
        if (fits && enough && roomy) break;
        if (mt == 1) {
            if (!fits) return conv_fail(PMB200_EINVAL, "conv2d_nhwc: halo tile does not fit in shared memory (dilation too large)");
            break;
        }
        mt /= 2;
    }
    const size_t smem = (size_t)p.rh * p.rw * p.ps * sizeof(float);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (kcin) {
        case 8: return launch_nt<8>(p, smem, nt, mt, precision, st);
        case 16: return launch_nt<16>(p, smem, nt, mt, precision, st);
        case 32: return launch_nt<32>(p, smem, nt, mt, precision, st);
        default: return launch_nt<64>(p, smem, nt, mt, precision, st);
    }
}

}  // extern "C"
