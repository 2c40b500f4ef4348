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
// The per-pixel arithmetic (dtype by dtype the reference's) lives in pm_geo_math.cuh, shared with the host-side formula
// tests.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/patchmatch_b200.h"
#include "pm_geo_math.cuh"

extern "C" int pmb200_internal_fail(int code, const char *msg);
extern "C" int pmb200_internal_launch_status(const char *what);

namespace {

using pmgeo::kCamDoubles;

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

__global__ void __launch_bounds__(256) geometric_filter_kernel(const GeoParams p) {
    extern __shared__ double s_cam[];  // [V][60]
    for (int i = threadIdx.x; i < p.V * kCamDoubles; i += blockDim.x) s_cam[i] = p.cams[i];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.H * p.W) return;
    const int x = n % p.W, y = n / p.W;
    const pmgeo::PixelResult r = pmgeo::filter_pixel(s_cam, p.src_depths, p.V, p.Hs, p.Ws, x, y, __ldg(p.ref_depth + n),
                                                     __ldg(p.confidence + n), p.pixel_thres, p.depth_thres, p.photo_thres,
                                                     p.mask_thres);
    p.mask_sum[n] = r.count;
    p.photo_mask[n] = r.photo ? 1 : 0;
    p.final_mask[n] = r.final ? 1 : 0;
    p.depth_avg[n] = r.depth_avg;
}

// ---- fusion half (reference eval.py:273-296): masked back-projection + colour gather + ORDER-PRESERVING compaction ----------
// Three small launches: per-block survivor counts, one-block exclusive scan of the counts, write.  A survivor's position is
// scan[block] + (survivors before it in the block), so the records come out in row-major pixel order exactly as the
// reference's boolean indexing produces them; each record is the 15-byte PLY vertex (float32 x, y, z, uint8 r, g, b).
constexpr int kFuseBlock = 256;

__global__ void __launch_bounds__(kFuseBlock) fuse_count_kernel(const unsigned char *mask, int n, int *counts) {
    const int i = blockIdx.x * kFuseBlock + threadIdx.x;
    const int c = __syncthreads_count(i < n && mask[i] != 0);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) fuse_scan_kernel(int *counts, int nblocks, int *total) {  // in place: exclusive scan
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? counts[i] : 0;
        int x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, off);
            if ((threadIdx.x & 31) >= off) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
            for (int off = 1; off < 32; off <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, off);
                if (threadIdx.x >= off) w += y;
            }
            s_warp[threadIdx.x] = w;  // inclusive over warps
        }
        __syncthreads();
        const int before_warp = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int carry = s_carry;
        if (i < nblocks) counts[i] = carry + before_warp + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + before_warp + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(kFuseBlock) fuse_write_kernel(const unsigned char *mask, const double *depth_avg, const float *rgb,
                                                                const double *cam, int W, int n, const int *offsets,
                                                                unsigned char *body) {
    __shared__ int s_warp[kFuseBlock / 32];
    __shared__ double s_cam[pmgeo::kFuseCamDoubles];
    if (threadIdx.x < pmgeo::kFuseCamDoubles) s_cam[threadIdx.x] = cam[threadIdx.x];
    const int i = blockIdx.x * kFuseBlock + threadIdx.x;
    const bool keep = i < n && mask[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (!keep) return;
    const int rank = offsets[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
    pmgeo::fuse_point(s_cam, i % W, i / W, depth_avg[i], rgb + (size_t)i * 3, body + (size_t)rank * pmgeo::kPlyVertexBytes);
}

}  // namespace

extern "C" int pmb200_fuse_points(const unsigned char *final_mask, const double *depth_avg, const float *ref_img_hwc,
                                  const double *cam25, int H, int W, unsigned char *ply_body_out, int *count_out,
                                  int *block_scratch, void *stream) {
    if (!final_mask || !depth_avg || !ref_img_hwc || !cam25 || !ply_body_out || !count_out || !block_scratch)
        return pmb200_internal_fail(PMB200_EINVAL, "fuse_points: null pointer");
    if (H < 1 || W < 1 || (long long)H * W >= (1ll << 31)) return pmb200_internal_fail(PMB200_EINVAL, "fuse_points: bad size");
    const int n = H * W, nblocks = (n + kFuseBlock - 1) / kFuseBlock;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    fuse_count_kernel<<<nblocks, kFuseBlock, 0, st>>>(final_mask, n, block_scratch);
    fuse_scan_kernel<<<1, 1024, 0, st>>>(block_scratch, nblocks, count_out);
    fuse_write_kernel<<<nblocks, kFuseBlock, 0, st>>>(final_mask, depth_avg, ref_img_hwc, cam25, W, n, block_scratch, ply_body_out);
    return pmb200_internal_launch_status("fuse_points");
}

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
