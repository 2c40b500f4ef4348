// pm_geo.cu -- geometric-consistency filtering of one reference depth map against V source depth maps in ONE launch
// (reference eval.py:86-190 `reproject_with_depth` / `check_geometric_consistency` and the
// per-view accumulation of `filter_depth`, eval.py:226-256).
//
// The reference runs this on one CPU thread in numpy + cv2.remap, one source view at a time, materialising ~20 H x W
// float64 temporaries per (reference, source) pair.  Here one thread owns one reference pixel and walks the source views:
// project with the pixel's depth, sample the source depth map, project back, threshold, accumulate -- the same warp /
// gather pattern as K-A with a depth map in place of the features.  No intermediate ever reaches memory: the launch
// reads (1 + V) depth maps + the confidence map and writes the three masks and the averaged depth.
//
// Arithmetic follows the reference's dtypes step by step so that the thresholded masks agree with it: camera matrices
// arrive already composed the way numpy composes them (float32 inverses / products, see ops.geometric_filter), the
// projections run in float64, the map coordinates are rounded to float32 before sampling, and the bilinear sample is
// cv2.remap's published algorithm -- coordinates rounded to 1/32 pixel (nearest-even), float32 weight table, taps outside
// the image contribute the constant border 0, float32 multiply and add WITHOUT contraction.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/patchmatch_b200.h"

extern "C" int pmb200_internal_fail(int code, const char *msg);
extern "C" int pmb200_internal_launch_status(const char *what);

namespace {

constexpr int kCamDoubles = 60;  // Kref^-1 (9) | T_ref->src (3x4) | Ksrc (9) | Ksrc^-1 (9) | T_src->ref (3x4) | Kref (9)

struct GeoParams {
    const float *ref_depth, *confidence, *src_depths;
    const double *cams;
    int V, H, W, Hs, Ws;
    double pixel_thres;
    float depth_thres, photo_thres;
    int mask_thres;
    int *mask_sum;
    unsigned char *photo_mask, *final_mask;
    double *depth_avg;
};

__device__ __forceinline__ void mat3(const double *m, double x, double y, double z, double &ox, double &oy, double &oz) {
    ox = m[0] * x + m[1] * y + m[2] * z;
    oy = m[3] * x + m[4] * y + m[5] * z;
    oz = m[6] * x + m[7] * y + m[8] * z;
}

__device__ __forceinline__ void mat34(const double *m, double x, double y, double z, double &ox, double &oy, double &oz) {
    ox = m[0] * x + m[1] * y + m[2] * z + m[3];
    oy = m[4] * x + m[5] * y + m[6] * z + m[7];
    oz = m[8] * x + m[9] * y + m[10] * z + m[11];
}

// cvRound(v * 32) as OpenCV computes it on the float32 product: nearest-even; non-finite / out-of-int-range -> INT_MIN
__device__ __forceinline__ int fixed_coord(float v) {
    const float s = rintf(__fmul_rn(v, 32.0f));
    if (!(fabsf(s) < 2147483648.0f)) return INT32_MIN;  // also catches NaN
    return (int)s;
}

// cv2.remap(src, x, y, INTER_LINEAR), float32 single channel, BORDER_CONSTANT(0)
__device__ __forceinline__ float remap_linear(const float *__restrict__ src, int rows, int cols, float x, float y) {
    const int sx = fixed_coord(x), sy = fixed_coord(y);
    const float fx = (float)(sx & 31) * 0.03125f, fy = (float)(sy & 31) * 0.03125f;
    const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
    const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
    const float w0 = __fmul_rn(gy, gx), w1 = __fmul_rn(gy, fx), w2 = __fmul_rn(fy, gx), w3 = __fmul_rn(fy, fx);
    const bool x0 = ix >= 0 && ix < cols, x1 = ix + 1 >= 0 && ix + 1 < cols;
    const bool y0 = iy >= 0 && iy < rows, y1 = iy + 1 >= 0 && iy + 1 < rows;
    const float t00 = (x0 && y0) ? __ldg(src + (size_t)iy * cols + ix) : 0.0f;
    const float t01 = (x1 && y0) ? __ldg(src + (size_t)iy * cols + ix + 1) : 0.0f;
    const float t10 = (x0 && y1) ? __ldg(src + (size_t)(iy + 1) * cols + ix) : 0.0f;
    const float t11 = (x1 && y1) ? __ldg(src + (size_t)(iy + 1) * cols + ix + 1) : 0.0f;
    float out = __fmul_rn(t00, w0);
    out = __fadd_rn(out, __fmul_rn(t01, w1));
    out = __fadd_rn(out, __fmul_rn(t10, w2));
    out = __fadd_rn(out, __fmul_rn(t11, w3));
    return out;
}

__global__ void __launch_bounds__(256) geometric_filter_kernel(const GeoParams p) {
    extern __shared__ double s_cam[];  // [V][60]
    for (int i = threadIdx.x; i < p.V * kCamDoubles; i += blockDim.x) s_cam[i] = p.cams[i];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.H * p.W) return;
    const int x = n % p.W, y = n / p.W;
    const float dref = __ldg(p.ref_depth + n);
    const double dx = (double)x, dy = (double)y, dd = (double)dref;
    float sum = 0.0f;  // Python's sum(): 0 + a_0 + a_1 + ... in float32
    int cnt = 0;
    for (int v = 0; v < p.V; ++v) {
        const double *c = s_cam + v * kCamDoubles;
        double rx, ry, rz, sxw, syw, szw, kx, ky, kz;
        mat3(c, dx * dd, dy * dd, dd, rx, ry, rz);                  // eval.py:116-117
        mat34(c + 9, rx, ry, rz, sxw, syw, szw);                    // :119-120
        mat3(c + 21, sxw, syw, szw, kx, ky, kz);                    // :122
        const double xs = kx / kz, ys = ky / kz;                    // :123
        const float sampled = remap_linear(p.src_depths + (size_t)v * p.Hs * p.Ws, p.Hs, p.Ws, (float)xs, (float)ys);  // :126-128
        const double sd = (double)sampled;
        double bx, by, bz, qx, qy, qz;
        mat3(c + 30, xs * sd, ys * sd, sd, bx, by, bz);             // :132-133
        mat34(c + 39, bx, by, bz, qx, qy, qz);                      // :135-136
        const float drep = (float)qz;                               // :138
        mat3(c + 51, qx, qy, qz, kx, ky, kz);                       // :139
        const float x2 = (float)(kx / kz), y2 = (float)(ky / kz);   // :140-142
        const double ex = (double)x2 - dx, ey = (double)y2 - dy;
        const double dist = sqrt(ex * ex + ey * ey);                // :180
        const float rel = __fdiv_rn(fabsf(__fsub_rn(drep, dref)), dref);  // :183-184 (float32)
        const bool ok = dist < p.pixel_thres && rel < p.depth_thres;      // :187 (NaN compares false)
        sum = __fadd_rn(sum, ok ? drep : 0.0f);                     // :188, :249
        cnt += ok ? 1 : 0;                                          // :248
    }
    const bool photo = __ldg(p.confidence + n) > p.photo_thres;     // :220
    p.mask_sum[n] = cnt;
    p.photo_mask[n] = photo ? 1 : 0;
    p.final_mask[n] = (photo && cnt >= p.mask_thres) ? 1 : 0;       // :254-255
    p.depth_avg[n] = (double)__fadd_rn(sum, dref) / (double)(cnt + 1);  // :252 (float32 / int32 -> float64 in numpy)
}

}  // namespace

extern "C" int pmb200_geometric_filter(const float *ref_depth, const float *confidence, const float *src_depths,
                                       const double *cams, int V, int H, int W, int Hs, int Ws, double geo_pixel_thres,
                                       float geo_depth_thres, float photo_thres, int geo_mask_thres, int *geo_mask_sum_out,
                                       unsigned char *photo_mask_out, unsigned char *final_mask_out, double *depth_avg_out,
                                       void *stream) {
    if (!ref_depth || !confidence || !src_depths || !cams || !geo_mask_sum_out || !photo_mask_out || !final_mask_out || !depth_avg_out)
        return pmb200_internal_fail(PMB200_EINVAL, "geometric_filter: null pointer");
    if (V < 1 || V > 64 || H < 1 || W < 1 || Hs < 1 || Ws < 1 || (long long)H * W >= (1ll << 31) || (long long)Hs * Ws >= (1ll << 31))
        return pmb200_internal_fail(PMB200_EINVAL, "geometric_filter: bad size (1 <= V <= 64)");
    GeoParams p;
    p.ref_depth = ref_depth; p.confidence = confidence; p.src_depths = src_depths; p.cams = cams;
    p.V = V; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws;
    p.pixel_thres = geo_pixel_thres; p.depth_thres = geo_depth_thres; p.photo_thres = photo_thres; p.mask_thres = geo_mask_thres;
    p.mask_sum = geo_mask_sum_out; p.photo_mask = photo_mask_out; p.final_mask = final_mask_out; p.depth_avg = depth_avg_out;
    const int threads = 256;
    const int blocks = (int)(((long long)H * W + threads - 1) / threads);
    geometric_filter_kernel<<<blocks, threads, (size_t)V * kCamDoubles * sizeof(double), reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return pmb200_internal_launch_status("geometric_filter");
}
