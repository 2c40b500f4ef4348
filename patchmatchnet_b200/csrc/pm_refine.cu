// pm_refine.cu -- K-R: the full-resolution half of Refinement in ONE launch, exact fp32.
//
//   up   = relu(bn(deconv(low)))                  ConvTranspose2d(8, 8, 3, stride 2, pad 1, output_padding 1)   models/net.py:85-86, :112
//   c0   = relu(bn(conv0(img)))                   3 -> 8, 3x3, pad 1                                            :82, :103
//   c3   = relu(bn(conv3(cat(up, c0))))           16 -> 8, 3x3, pad 1                                           :88, :114
//   res  = conv_res(c3)                           8 -> 1, 3x3, pad 1, no bias                                   :89, :115
//   out  = (nearest_up2x(d) + res) * span + lo,   d = (depth_half - lo) / span                                   :104-106, :117-120
//
// Before: six launches of the conv family (every one with <= 16 channels, i.e. mostly MMA padding, and the 3xTF32 split on
// top) plus five ATen elementwise / upsample / layout kernels, with the 16-channel full-resolution map (21 MB at 640x512)
// written and read back.  Here one CTA owns a 32 x 8 output tile and walks the chain through shared memory with halo
// recompute: `both` (16 channels) on the tile + 2, conv3 on the tile + 1, res on the tile; 2,040 folded weights ride in the
// kernel parameter block (8.2 KB; CUDA >= 12.1 allows 32 KB on sm_70+), every multiply-add is a plain fp32 FFMA.
//   * the transposed conv is evaluated in its zero-stuffed form by output parity: an even row uses filter row 1 only, an odd row
//     rows 0 and 2 (same for columns), so a pixel costs 1, 2 or 4 taps instead of 9, and the threads of a loop share one parity
//     pattern (no divergence): x pairs (even, odd) per thread, even rows in one loop, odd rows in another;
//   * intermediate maps live as float4 planes [channel quad][pixel] (consecutive lanes, consecutive 16-byte slots);
//   * the conv3 buffer aliases the image / low-resolution staging area (dead by then): 39 KB of static shared memory.
// FFMA per output pixel: res 72 + conv3 1152 x 1.33 + conv0 216 x 1.69 + deconv 144 x 1.69 = 2,210 (halo recompute included);
// at 640x512 that is 0.72 G FFMA = 20 us at 100 % fp32 issue.
#if !defined(PM_EMU)  // host emulation build (tests/warp_emu.h) brings its own CUDA vocabulary
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/patchmatch_b200.h"

extern "C" int pmb200_internal_fail(int code, const char *msg);  // pm_kernels.cu: sets pmb200_last_error()
#if !defined(PM_EMU)
extern "C" int pmb200_internal_launch_status(const char *what);
#endif

namespace {

constexpr int kRTW = 32, kRTH = 8, kRT = 128;          // output tile, threads
constexpr int kBW = kRTW + 4, kBH = kRTH + 4;          // `both` region (conv3's halo of conv3's halo): 36 x 12
constexpr int kCW = kRTW + 2, kCH = kRTH + 2;          // conv3 region: 34 x 10
constexpr int kIW2 = kRTW + 6, kIH2 = kRTH + 6;        // image halo for conv0 on the `both` region: 38 x 14
constexpr int kLW = kRTW / 2 + 4, kLH = kRTH / 2 + 4;  // low-resolution halo for the transposed conv: 20 x 8
constexpr int kBPix = kBW * kBH, kCPix = kCW * kCH;    // 432, 340
static_assert(kBPix <= 4 * kRT && kCPix <= 3 * kRT, "loop trip counts below assume these");

struct RefineParams {
    const float *low;   // [N,h,w,8]
    const float *img;   // [N,3,H,W]
    const float *dh;    // [N,1,h,w]
    const float *dmin;  // [N]
    const float *dmax;  // [N]
    float *out;         // [N,1,H,W]
    int N, H, W, tiles_x, tiles_y;
    float wd[9 * 64];   // transposed conv in zero-stuffed (flipped) form: [(ky*3+kx)][ci][co]
    float bd[8];
    float w0[27 * 8];   // [(ci*9 + ky*3 + kx)][co]
    float b0[8];
    float w3[9 * 16 * 8];  // [((ky*3+kx)*16 + ci)][co]
    float b3[8];
    float wr[9 * 8];    // [(ky*3+kx)][ci]
};

__device__ __forceinline__ void fma8(float (&a)[8], const float4 lo, const float4 hi, const float *w) {  // a[co] += v[ci] * w[ci*8+co]
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int ci = 0; ci < 8; ++ci)
#pragma unroll
        for (int co = 0; co < 8; ++co) a[co] = fmaf(v[ci], w[ci * 8 + co], a[co]);
}

__global__ void __launch_bounds__(kRT) refine_full_kernel(const __grid_constant__ RefineParams p) {
    // shared memory: [both: 4 planes x 432 float4][stage: image halo + low halo, later the conv3 planes]
    __shared__ float4 both[4][kBPix];
    constexpr int kStageFloats = 3 * kIH2 * kIW2 + 2 * kLH * kLW * 4;
    static_assert(kStageFloats >= 2 * kCPix * 4, "the conv3 planes must fit in the staging area they alias");
    __shared__ __align__(16) float stage[kStageFloats];
    float (*img)[kIH2][kIW2] = reinterpret_cast<float (*)[kIH2][kIW2]>(stage);
    float4 (*lowh)[kLH * kLW] = reinterpret_cast<float4 (*)[kLH * kLW]>(stage + 3 * kIH2 * kIW2);
    float4 (*c3)[kCPix] = reinterpret_cast<float4 (*)[kCPix]>(stage);

    const int tile = blockIdx.x;
    const int per_img = p.tiles_x * p.tiles_y;
    const int n = tile / per_img, tt = tile - n * per_img;
    const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
    const int y0 = ty * kRTH, x0 = tx * kRTW;  // both even
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = p.H >> 1, w = p.W >> 1;

    // ---- stage the image halo (rows y0-3 .., cols x0-3 ..; one warp per row) and the low-resolution halo (rows y0/2-2 ..)
    {
        const float *xin = p.img + (size_t)n * 3 * p.H * p.W;
        constexpr int kRows = 3 * kIH2, kRowsPerWarp = (kRows + 3) / 4;
        float v0[kRowsPerWarp], v1[kRowsPerWarp];
        const int gx0 = x0 - 3 + lane, gx1 = gx0 + 32;
        const bool in0 = gx0 >= 0 && gx0 < p.W, in1 = lane < kIW2 - 32 && gx1 < p.W;
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
            const int row = warp + 4 * k;
            const int ci = row / kIH2, r = row - ci * kIH2;
            const int gy = y0 - 3 + r;
            const bool rin = row < kRows && gy >= 0 && gy < p.H;
            const float *src = xin + ((size_t)ci * p.H + (rin ? gy : 0)) * p.W;
            v0[k] = (rin && in0) ? src[gx0] : 0.0f;
            v1[k] = (rin && in1) ? src[gx1] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
            const int row = warp + 4 * k;
            if (row < kRows) {
                const int ci = row / kIH2, r = row - ci * kIH2;
                img[ci][r][lane] = v0[k];
                if (lane < kIW2 - 32) img[ci][r][32 + lane] = v1[k];
            }
        }
        const int ly0 = (y0 >> 1) - 2, lx0 = (x0 >> 1) - 2;
        const float4 *lin = reinterpret_cast<const float4 *>(p.low) + (size_t)n * h * w * 2;
        for (int i = tid; i < 2 * kLH * kLW; i += kRT) {
            const int half = i & 1, pos = i >> 1;
            const int r = pos / kLW, c = pos - r * kLW;
            const int ly = ly0 + r, lx = lx0 + c;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (ly >= 0 && ly < h && lx >= 0 && lx < w) v = lin[((size_t)ly * w + lx) * 2 + half];
            lowh[half][pos] = v;
        }
    }
    __syncthreads();

    // ---- both[0..7] = relu(transposed conv + folded BN) on the 36 x 12 region, by output parity.  Region row r <-> y = y0-2+r
    // (same parity as r), column pair pc <-> x = x0-2+2pc (even) and x+1 (odd); low-halo coordinates: row r/2+1 (even y) or
    // (r-1)/2+1 and +1 (odd y), columns pc+1 and pc+2.
    auto store_up = [&](int r, int c, const float (&a)[8]) {
        const int gy = y0 - 2 + r, gx = x0 - 2 + c;
        const bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
        float o[8];
#pragma unroll
        for (int co = 0; co < 8; ++co) o[co] = inside ? fmaxf(a[co], 0.0f) : 0.0f;
        both[0][r * kBW + c] = make_float4(o[0], o[1], o[2], o[3]);
        both[1][r * kBW + c] = make_float4(o[4], o[5], o[6], o[7]);
    };
    for (int i = tid; i < (kBH / 2) * (kBW / 2); i += kRT) {  // even rows: filter row 1 only
        const int rr = i / (kBW / 2), pc = i - rr * (kBW / 2);
        const int r = 2 * rr, lp = (rr + 1) * kLW + pc + 1;
        const float4 a_lo = lowh[0][lp], a_hi = lowh[1][lp], b_lo = lowh[0][lp + 1], b_hi = lowh[1][lp + 1];
        float e[8], o[8];
#pragma unroll
        for (int co = 0; co < 8; ++co) e[co] = o[co] = p.bd[co];
        fma8(e, a_lo, a_hi, p.wd + (1 * 3 + 1) * 64);  // even x: filter column 1
        fma8(o, a_lo, a_hi, p.wd + (1 * 3 + 0) * 64);  // odd x: columns 0 and 2
        fma8(o, b_lo, b_hi, p.wd + (1 * 3 + 2) * 64);
        store_up(r, 2 * pc, e);
        store_up(r, 2 * pc + 1, o);
    }
    for (int i = tid; i < (kBH / 2) * (kBW / 2); i += kRT) {  // odd rows: filter rows 0 and 2
        const int rr = i / (kBW / 2), pc = i - rr * (kBW / 2);
        const int r = 2 * rr + 1, lp = (rr + 1) * kLW + pc + 1;
        const float4 a_lo = lowh[0][lp], a_hi = lowh[1][lp], b_lo = lowh[0][lp + 1], b_hi = lowh[1][lp + 1];
        const float4 c_lo = lowh[0][lp + kLW], c_hi = lowh[1][lp + kLW], d_lo = lowh[0][lp + kLW + 1], d_hi = lowh[1][lp + kLW + 1];
        float e[8], o[8];
#pragma unroll
        for (int co = 0; co < 8; ++co) e[co] = o[co] = p.bd[co];
        fma8(e, a_lo, a_hi, p.wd + (0 * 3 + 1) * 64);
        fma8(e, c_lo, c_hi, p.wd + (2 * 3 + 1) * 64);
        fma8(o, a_lo, a_hi, p.wd + (0 * 3 + 0) * 64);
        fma8(o, b_lo, b_hi, p.wd + (0 * 3 + 2) * 64);
        fma8(o, c_lo, c_hi, p.wd + (2 * 3 + 0) * 64);
        fma8(o, d_lo, d_hi, p.wd + (2 * 3 + 2) * 64);
        store_up(r, 2 * pc, e);
        store_up(r, 2 * pc + 1, o);
    }

    // ---- both[8..15] = relu(conv0(img) + folded BN) on the same region
    for (int pos = tid; pos < kBPix; pos += kRT) {
        const int r = pos / kBW, c = pos - r * kBW;
        const int gy = y0 - 2 + r, gx = x0 - 2 + c;
        float a[8];
#pragma unroll
        for (int co = 0; co < 8; ++co) a[co] = p.b0[co];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float v = img[ci][r + ky][c + kx];
#pragma unroll
                    for (int co = 0; co < 8; ++co) a[co] = fmaf(v, p.w0[(ci * 9 + ky * 3 + kx) * 8 + co], a[co]);
                }
        const bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
#pragma unroll
        for (int co = 0; co < 8; ++co) a[co] = inside ? fmaxf(a[co], 0.0f) : 0.0f;
        both[2][pos] = make_float4(a[0], a[1], a[2], a[3]);
        both[3][pos] = make_float4(a[4], a[5], a[6], a[7]);
    }
    __syncthreads();  // `both` complete; the staging area is dead from here on

    // ---- c3 = relu(conv3(both) + folded BN) on the 34 x 10 region; three positions per thread share the weight loads
    {
        float acc[3][8];
        int pos[3], bp[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            pos[q] = tid + q * kRT;
            const int pc = pos[q] < kCPix ? pos[q] : 0;  // out-of-range slots compute a valid position and are not stored
            const int r = pc / kCW, c = pc - r * kCW;
            bp[q] = r * kBW + c;
#pragma unroll
            for (int co = 0; co < 8; ++co) acc[q][co] = p.b3[co];
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int b = bp[q] + ky * kBW + kx;
                    fma8(acc[q], both[0][b], both[1][b], p.w3 + ((ky * 3 + kx) * 16 + 0) * 8);
                    fma8(acc[q], both[2][b], both[3][b], p.w3 + ((ky * 3 + kx) * 16 + 8) * 8);
                }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (pos[q] < kCPix) {
                const int r = pos[q] / kCW, c = pos[q] - r * kCW;
                const int gy = y0 - 1 + r, gx = x0 - 1 + c;
                const bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
                float o[8];
#pragma unroll
                for (int co = 0; co < 8; ++co) o[co] = inside ? fmaxf(acc[q][co], 0.0f) : 0.0f;
                c3[0][pos[q]] = make_float4(o[0], o[1], o[2], o[3]);
                c3[1][pos[q]] = make_float4(o[4], o[5], o[6], o[7]);
            }
    }
    __syncthreads();

    // ---- res + the residual tail: two output pixels per thread (rows oy, oy + 4)
    const float lo = p.dmin[n], span = p.dmax[n] - lo;
    const int ox = lane, oy = warp;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int ry = oy + 4 * q;
        float r = 0.0f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int cp = (ry + ky) * kCW + ox + kx;
                const float4 a = c3[0][cp], b = c3[1][cp];
                const float *wr = p.wr + (ky * 3 + kx) * 8;
                r = fmaf(a.x, wr[0], r); r = fmaf(a.y, wr[1], r); r = fmaf(a.z, wr[2], r); r = fmaf(a.w, wr[3], r);
                r = fmaf(b.x, wr[4], r); r = fmaf(b.y, wr[5], r); r = fmaf(b.z, wr[6], r); r = fmaf(b.w, wr[7], r);
            }
        const int gy = y0 + ry, gx = x0 + ox;
        if (gy < p.H && gx < p.W) {
            const float d = (p.dh[((size_t)n * h + (gy >> 1)) * w + (gx >> 1)] - lo) / span;  // nearest x2 of the normalised map
            p.out[((size_t)n * p.H + gy) * p.W + gx] = (d + r) * span + lo;
        }
    }
}

}  // namespace

extern "C" {

// The full-resolution half of Refinement (reference models/net.py:103, :112-120) in one launch; see the file header.
//   low_nhwc    device [N,h,w,8]: relu(conv2(relu(conv1(d))))  (pmb200_refine_low)      img_nchw  device [N,3,2h,2w] contiguous
//   depth_half  device [N,1,h,w];  depth_min / depth_max device [N];  depth_out device [N,1,2h,2w]
//   host_wd     HOST [8 in][8 out][3][3] (ConvTranspose2d layout, BatchNorm folded over the OUTPUT channel), host_bd [8]
//   host_w0     HOST [8][3][3][3], host_b0 [8];  host_w3 HOST [8][16][3][3] (input channels: 8 upsampled, then 8 image), host_b3 [8]
//   host_wr     HOST [1][8][3][3] (the residual conv has neither bias nor BatchNorm)
int pmb200_refine_full(const float *low_nhwc, const float *img_nchw, const float *depth_half, const float *depth_min, const float *depth_max,
                       const float *host_wd, const float *host_bd, const float *host_w0, const float *host_b0, const float *host_w3,
                       const float *host_b3, const float *host_wr, float *depth_out, int N, int H, int W, void *stream) {
    if (!low_nhwc || !img_nchw || !depth_half || !depth_min || !depth_max || !host_wd || !host_bd || !host_w0 || !host_b0 || !host_w3 || !host_b3 ||
        !host_wr || !depth_out)
        return pmb200_internal_fail(PMB200_EINVAL, "refine_full: null pointer");
    if (N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1)) return pmb200_internal_fail(PMB200_EINVAL, "refine_full: H and W must be even (twice the low-resolution map)");
    if (reinterpret_cast<uintptr_t>(low_nhwc) & 15u) return pmb200_internal_fail(PMB200_EINVAL, "refine_full: low_nhwc must be 16-byte aligned");
    static RefineParams zero_init;  // zero-initialised template (the struct is too large to brace-initialise on the stack cheaply)
    RefineParams p = zero_init;
    p.low = low_nhwc; p.img = img_nchw; p.dh = depth_half; p.dmin = depth_min; p.dmax = depth_max; p.out = depth_out;
    p.N = N; p.H = H; p.W = W;
    p.tiles_x = (W + kRTW - 1) / kRTW; p.tiles_y = (H + kRTH - 1) / kRTH;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * N;
    if (tiles > 0x7fffffffLL) return pmb200_internal_fail(PMB200_EINVAL, "refine_full: too many tiles");
    for (int co = 0; co < 8; ++co) {
        p.bd[co] = host_bd[co]; p.b0[co] = host_b0[co]; p.b3[co] = host_b3[co];
        for (int ci = 0; ci < 8; ++ci)
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx)  // zero-stuffed form: the kernel of the equivalent stride-1 conv is the flipped one
                    p.wd[((ky * 3 + kx) * 8 + ci) * 8 + co] = host_wd[((ci * 8 + co) * 3 + (2 - ky)) * 3 + (2 - kx)];
        for (int ci = 0; ci < 3; ++ci)
            for (int t = 0; t < 9; ++t) p.w0[(ci * 9 + t) * 8 + co] = host_w0[(co * 3 + ci) * 9 + t];
        for (int ci = 0; ci < 16; ++ci)
            for (int t = 0; t < 9; ++t) p.w3[(t * 16 + ci) * 8 + co] = host_w3[(co * 16 + ci) * 9 + t];
    }
    for (int ci = 0; ci < 8; ++ci)
        for (int t = 0; t < 9; ++t) p.wr[t * 8 + ci] = host_wr[ci * 9 + t];
#if defined(PM_EMU)
    (void)stream;
    emu::launch(dim3((unsigned)tiles), dim3(kRT), 0, [&] { refine_full_kernel(p); });
    return 0;
#else
    refine_full_kernel<<<(unsigned)tiles, kRT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return pmb200_internal_launch_status("refine_full");
#endif
}

}  // extern "C"
