// pm_conv5.cu -- K-D5: channels-last 2-D convolution as an implicit GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA), fp32-accurate through the error-compensated
// 3xTF32 split.  Serves the FLOP-bound layers of the cascade in the fp32-accurate (timed, parity) mode: FeatureNet conv2..10
// (reference models/net.py:9-70), the stage-2/3 offset convs (models/patchmatch.py:288-311), Refinement's 16->8 conv
// (models/net.py:73-122).  The memory-bound layers and everything with a fused epilogue stay on pm_conv.cu (mma.sync).
//
// GEMM view: D[128 output pixels, Cout] = sum over filter taps of A_tap[128, Cin] . W_tap[Cin, Cout].
//   * A_tap is ONE TMA box per tap: tensor map over the channels-last input {Cin, W, H, N}, box {<=32 channels, 16*s, 8*s, 1}
//     with element strides {1, s, s, 1} (s = conv stride), placed at (x0*s + kx*dil - pad, y0*s + ky*dil - pad): the
//     padding border and the map edge come back as zeros from the copy engine, the 128 pixels of the 16 x 8 output tile
//     land as 128 rows of a K-major operand in the shared-memory swizzle (32 / 64 / 128 B = row bytes) tcgen05 reads.
//   * W_tap (hi and lo parts, Cout padded to a multiple of 16) is pre-packed by the host IN that shared-memory image
//     (ops.pack_conv_filter_tc5) and fetched with one bulk copy per tap.
//   * 3xTF32: four "worker" warps split every landed A tile in place into A_hi (13 low mantissa bits cleared -- what the
//     tensor core would read anyway) and A_lo = A - A_hi (exact), then the MMA warp issues per 8-channel slice
//     A_lo.W_hi, A_hi.W_lo, A_hi.W_hi into the same fp32 TMEM accumulator: >= 21 bits of every product.
//   * Epilogue: tcgen05.ld (32 lanes x 32 bit, 16 columns at a time) -> bias, ReLU -> channels-last stores (channel slice
//     of a wider tensor allowed).
// Roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = 3xTF32 split,
// warps 6..9 = epilogue.  Two accumulators in TMEM: the epilogue of tile i drains one while the MMAs of tile i+1 fill the
// other (run 3 measured the single-accumulator version, whose split warps also ran the epilogue, at 1.5 us per filter tap).
// Persistent CTAs walk tiles blockIdx.x, +gridDim.x, ...; an S-deep ring of (A, A_lo, W_hi, W_lo) stages decouples the roles;
// every mbarrier wait is bounded (a pipeline bug traps instead of hanging the GPU).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/patchmatch_b200.h"

extern "C" int pmb200_internal_fail(int code, const char *msg);
extern "C" int pmb200_internal_launch_status(const char *what);

namespace {

constexpr int kTW = 16, kTH = 8;  // output tile: 128 pixels = the M of one UMMA
constexpr int kThreads = 320;

struct Conv5Params {
    const float *wpack;  // [tap][kblock][w_hi rows Npad | w_lo rows Npad][RB bytes], swizzled image
    const float *bias;
    float *y;
    int N, H, W, Ho, Wo, Cin, Cout, Npad, KS, S, pad, dil, relu, ycs, yco;
    int tiles_x, tiles_y, total_tiles;
    int RB;        // row bytes of one K block: min(Cin, 32) * 4
    int KB;        // K blocks: Cin / 32 (1 when Cin <= 32)
    int stages;
    int a_bytes;   // 128 * RB * KB
    int w_bytes;   // Npad * RB * KB (one of hi / lo)
    int stage_bytes;
    uint32_t idesc, idesc2;  // N = Npad / N = 2 Npad
    uint32_t layout_type;  // UMMA smem-descriptor swizzle code: 2 = 128B, 4 = 64B, 6 = 32B
    uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin)
        if (mbar_try_wait(b, parity)) return;
    __trap();
}
__device__ __forceinline__ void tma_box_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<unsigned long long>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {  // arrives on `bar` once every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Shared-memory matrix descriptor of a K-major operand tile in the 32/64/128-byte swizzle (rows of RB bytes, 8-row groups
// 8*RB apart): bits 0-13 start address >> 4, 16-29 leading byte offset (unused when swizzled: 1), 32-45 stride byte offset
// (8*RB) >> 4, 46 version (Blackwell), 61-63 swizzle code (2 = 128B, 4 = 64B, 6 = 32B).  The kernels build it as two words.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kThreads) conv5_kernel(const Conv5Params p, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ unsigned char smem_raw[];
    // the swizzled operand tiles need 1024-byte alignment: align the dynamic segment by hand (the launch adds the slack)
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // [0, 256): barriers + TMEM address; stages start at 1024
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);  // [stages]  TMA landed
    uint64_t *split = full + 8;                            // [stages]  A_hi / A_lo written
    uint64_t *empty = split + 8;                           // [stages]  MMAs that read the stage have completed
    uint64_t *acc_full = empty + 8;                        // [2] tile accumulated in TMEM buffer a
    uint64_t *acc_empty = acc_full + 2;                    // [2] the epilogue has read buffer a
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    unsigned char *stage0 = smem + 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 4);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: Npad fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int T = p.KS * p.KS;
    const int per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0) {
            int s = 0;
            uint32_t par = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int n = tile / per_img, tt = tile - n * per_img;
                const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
                const int x_in0 = tx * kTW * p.S - p.pad, y_in0 = ty * kTH * p.S - p.pad;
                for (int t = 0; t < T; ++t) {
                    mbar_wait(&empty[s], par ^ 1u);
                    unsigned char *st = stage0 + (size_t)s * p.stage_bytes;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(p.a_bytes + 2 * p.w_bytes));
                    const int ky = t / p.KS, kx = t - ky * p.KS;
                    for (int kb = 0; kb < p.KB; ++kb)
                        tma_box_4d(st + (size_t)kb * 128 * p.RB, &xmap, kb * 32, x_in0 + kx * p.dil, y_in0 + ky * p.dil, n, &full[s]);
                    bulk_g2s(st + 2 * (size_t)p.a_bytes, reinterpret_cast<const unsigned char *>(p.wpack) + (size_t)t * 2 * p.w_bytes,
                             (uint32_t)(2 * p.w_bytes), &full[s]);
                    if (++s == p.stages) { s = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------- MMA issuer -------------------------------------------------
        // Whole warp on warp-uniform values, tcgen05 instructions from the elected lane (see conv5h_kernel); 2 MMAs per K
        // slice: a_hi x [w_hi | w_lo] (N = 2 Npad, columns [0, 2 Npad)) and a_lo x w_hi (N = Npad, columns [0, Npad)).
        const bool leader = elect_one();
        int s = 0, it = 0;
        uint32_t par = 0;
        const int per_row = p.RB / 32;  // 8-channel slices per operand row
        // descriptor words: lo = start address >> 4 | LBO field 1, hi = SBO (8 rows) >> 4 | version | swizzle code << 29
        const uint32_t hiword = ((8u * (uint32_t)p.RB) >> 4) | (1u << 14) | (p.layout_type << 29);
        const uint32_t stage_u32 = smem_u32(stage0);
        const uint32_t a_kb16 = (128u * (uint32_t)p.RB) >> 4, w_kb16 = (2u * (uint32_t)p.Npad * (uint32_t)p.RB) >> 4;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;                                  // accumulator buffer of this tile
            const uint32_t tacc = tmem + (uint32_t)(ab * 2 * p.Npad);
            mbar_wait(&acc_empty[ab], (((uint32_t)it >> 1) & 1u) ^ 1u);  // the epilogue of tile it-2 has drained this buffer
            tc_fence_after();
            uint32_t first = 0;
            for (int t = 0; t < T; ++t) {
                mbar_wait(&split[s], par);
                tc_fence_after();
                const uint32_t a_hi = ((stage_u32 + (uint32_t)(s * p.stage_bytes)) >> 4) | (1u << 16);
                const uint32_t a_lo = a_hi + ((uint32_t)p.a_bytes >> 4), w0 = a_lo + ((uint32_t)p.a_bytes >> 4);
                for (int kb = 0; kb < p.KB; ++kb) {
                    for (int kin = 0; kin < per_row; ++kin) {
                        const uint32_t ao = (uint32_t)kb * a_kb16 + 2u * kin, wo = (uint32_t)kb * w_kb16 + 2u * kin;  // 32 bytes per slice
                        const uint64_t dw = ((uint64_t)hiword << 32) | (w0 + wo);
                        if (leader) {
                            tc_mma_tf32(tacc, ((uint64_t)hiword << 32) | (a_hi + ao), dw, p.idesc2, first);
                            tc_mma_tf32(tacc, ((uint64_t)hiword << 32) | (a_lo + ao), dw, p.idesc, 1);
                        }
                        first = 1;
                    }
                }
                if (leader) {
                    tc_commit(&empty[s]);                        // stage free once these MMAs have read it
                    if (t == T - 1) tc_commit(&acc_full[ab]);    // accumulator complete
                }
                if (++s == p.stages) { s = 0; par ^= 1u; }
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------ 3xTF32 split ------------------------------------------------
        const int wt = threadIdx.x - 64;     // 0..127: operand row handled by this thread
        int s = 0;
        uint32_t par = 0;
        const int chunks = p.RB / 16;        // 16-byte chunks per operand row of one K block
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            for (int t = 0; t < T; ++t) {
                mbar_wait(&full[s], par);
                unsigned char *a = stage0 + (size_t)s * p.stage_bytes;
                for (int kb = 0; kb < p.KB; ++kb) {
                    unsigned char *rowp = a + (size_t)kb * 128 * p.RB + (size_t)wt * p.RB;
                    for (int j = 0; j < chunks; ++j) {
                        const int c = (j + wt) & (chunks - 1);  // rotated: the 8 lanes of a quarter-warp hit different banks
                        float4 *q = reinterpret_cast<float4 *>(rowp + c * 16);
                        const float4 v = *q;
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
                        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
                        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
                        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
                        *q = h;
                        *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(q) + p.a_bytes) = l;
                    }
                }
                proxy_fence_async();  // generic-proxy writes above -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
                if (++s == p.stages) { s = 0; par ^= 1u; }
            }
        }
    } else {
        // -------------------------------------------------- epilogue --------------------------------------------------
        const int quad = warp & 3;           // TMEM lane quadrant this warp may read: lanes 32*quad .. 32*quad+31
        const int row = quad * 32 + lane;    // accumulator row = pixel of the tile
        const bool vec4 = ((p.ycs | p.yco) & 3) == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            mbar_wait(&acc_full[ab], ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            const int n = tile / per_img, tt = tile - n * per_img;
            const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
            const int oy = ty * kTH + row / kTW, ox = tx * kTW + row % kTW;
            const bool inside = oy < p.Ho && ox < p.Wo;
            float *dst = p.y + (((size_t)n * p.Ho + (inside ? oy : 0)) * p.Wo + (inside ? ox : 0)) * p.ycs + p.yco;
            for (int c0 = 0; c0 < p.Npad; c0 += 16) {
                float v[16], u[16];
                const uint32_t tcol = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * 2 * p.Npad + c0);
                tmem_ld16(tcol, v);                        // a_hi.w_hi + a_lo.w_hi
                tmem_ld16(tcol + (uint32_t)p.Npad, u);     // a_hi.w_lo
                if (c0 + 16 >= p.Npad) {  // last read of the accumulator: hand it back before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[ab]);
                }
                if (!inside) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int co = c0 + i;
                    float o = (v[i] + u[i]) + ((p.bias && co < p.Cout) ? __ldg(p.bias + co) : 0.0f);
                    v[i] = p.relu ? fmaxf(o, 0.0f) : o;
                }
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const int co = c0 + i;
                    if (co + 3 < p.Cout && vec4) {
                        *reinterpret_cast<float4 *>(dst + co) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (co + e < p.Cout) dst[co + e] = v[i + e];
                    }
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// K-D5h: the stride-1 form.  conv5_kernel above re-reads (and re-splits) the input once per filter tap; for stride 1 every
// tap's A operand is just a SHIFTED view of one halo tile, if that tile is stored so that a shift keeps the tensor core's
// canonical layout intact.  The no-swizzle K-major layout does: its core matrix is 8 rows x 16 bytes stored contiguously,
// so with the halo tile kept as channel-chunk PLANES [Cin/4][rows][cols][4 floats] (16 bytes per pixel and plane)
//   * 8 consecutive pixels of one plane ARE a core matrix (128 contiguous bytes),
//   * the next 8-row group of the operand = the next output row = + (halo cols * 16) bytes      (descriptor SBO),
//   * the second half of an 8-channel K slice = the next plane = + (halo rows * cols * 16) bytes (descriptor LBO),
//   * filter tap (ky, kx) = start address + ((ky*dil) * cols + kx*dil) * 16 bytes -- always 16-byte aligned, no swizzle phase.
// The halo tile arrives as ONE dense TMA box {Cin, cols, rows, 1} (rows of Cin*4 bytes, zero fill for the padding) in a staging
// buffer; the four split warps, which touch every element anyway, write A_hi / A_lo TRANSPOSED into the planes (chunk order
// rotated by pixel: conflict-free reads and writes).  (Run 6 measured the first version, whose 5-D tensor map {4, W, H, Cin/4, N}
// made the TMA engine gather the planes itself in 16-byte pieces: correct, but 55 us on conv6/7 -- the engine is slow at that
// granularity.)  Output tile: 8 wide x 16 tall.  Load and split happen ONCE per tile; the filter taps stream through a small
// ring.  Roles and accumulator double buffering as conv5_kernel.
//
// MMA economy (tools/umma_rate.cu, profiles/r2_umma_rate.json): one tcgen05.mma with M = 128 costs max(49, N/2) cycles whatever
// its kind -- below N = 128 the instruction, not the arithmetic, is the unit of cost, and the first version of this kernel
// (3 MMAs of N = Npad per K slice, ~23 issue-thread instructions each) ran at 146 cycles per MMA, bound by the issuing thread.
// So (a) the filter tap is stored as ONE operand [chunk][w_hi rows | w_lo rows][4 floats] (ops.pack_conv_filter_tc5h):
// a_hi x [w_hi | w_lo] is one MMA of N = 2 Npad into columns [0, 2 Npad), a_lo x w_hi one of N = Npad into [0, Npad) -- 2 MMAs
// per K slice instead of 3, the epilogue adds the two column halves; (b) the issue loop carries descriptors as (lo, hi) words
// and only adds 16-byte offsets to the low word.
// ----------------------------------------------------------------------------------------------------------------------
constexpr int kHTW = 8, kHTH = 16;

struct Conv5hParams {
    const float *wpack;  // [tap][Cin/4 chunks][hi rows Npad | lo rows Npad][4]
    const float *bias;
    float *y;
    int N, H, W, Ho, Wo, Cin, Cout, Npad, KS, pad, dil, relu, ycs, yco;
    int tiles_x, tiles_y, total_tiles;
    int hcols, hrows;     // halo tile: 8 + dil*(KS-1) columns, 16 + dil*(KS-1) rows
    int plane_bytes;      // hrows * hcols * 16
    int a_bytes;          // plane_bytes * Cin/4 (one of hi / lo) = what one TMA box delivers into a staging buffer
    int a_stride;         // a_bytes rounded up to 128: distance between staging buffers, hi -> lo, plane sets (2 * a_stride)
    int hbufs;            // plane sets (hi, lo): 2 when they fit (split of tile i+1 under the MMAs of tile i), else 1
    int sbufs;            // staging buffers: 2 when they fit (TMA of tile i+1 under the split of tile i), else 1
    int w_bytes;          // Npad * Cin * 4 (one of hi / lo)
    int wstages;
    int w_resident;       // 1: wstages == KS*KS, every tap loaded once per CTA; 0: taps stream through the ring per tile
    uint32_t idesc, idesc2, tmem_cols;  // idesc: N = Npad, idesc2: N = 2 Npad
    long long *trace;     // debugging aid (pmb200_debug_conv5h_trace): clock64 stamps of CTA 0's roles, [role][tile 0..15][event 0..3]
};

__device__ __forceinline__ void stamp(const Conv5hParams &p, int role, int it, int ev) {
    if (p.trace && blockIdx.x == 0 && it < 16) p.trace[(role * 16 + it) * 4 + ev] = clock64();
}

// Operands of this kernel are no-swizzle K-major: core matrices of 8 rows x 16 bytes; descriptor LBO = distance of the
// K-adjacent core matrix, SBO = distance of the next 8-row group, version bit 46, layout type 0.
// the MMAs of one filter tap: per 8-channel slice a_hi x [w_hi | w_lo] (N = 2 Npad) and a_lo x w_hi (N = Npad)
template <int KSL>
__device__ __forceinline__ void issue_tap(bool leader, uint32_t tacc, uint32_t ah, uint32_t al, uint32_t w, uint32_t a_kstep, uint32_t w_kstep,
                                          uint32_t a_hiword, uint32_t w_hiword, uint32_t idesc, uint32_t idesc2, uint32_t first) {
    if (!leader) return;
#pragma unroll
    for (int k = 0; k < KSL; ++k) {
        const uint64_t dw = ((uint64_t)w_hiword << 32) | (w + k * w_kstep);
        tc_mma_tf32(tacc, ((uint64_t)a_hiword << 32) | (ah + k * a_kstep), dw, idesc2, k == 0 ? first : 1u);
        tc_mma_tf32(tacc, ((uint64_t)a_hiword << 32) | (al + k * a_kstep), dw, idesc, 1);
    }
}

__global__ void __launch_bounds__(kThreads) conv5h_kernel(const Conv5hParams p, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t *h_full = reinterpret_cast<uint64_t *>(smem);  // [2] halo tile landed in a staging buffer
    uint64_t *s_free = h_full + 2;                           // [2] the split has read the staging buffer
    uint64_t *h_split = s_free + 2;                          // [2] A_hi / A_lo planes written
    uint64_t *h_empty = h_split + 2;                         // [2] every MMA of the tile has read the plane set
    uint64_t *w_full = h_empty + 2;                          // [25] filter tap landed
    uint64_t *w_empty = w_full + 25;                         // [25] its MMAs have completed
    uint64_t *acc_full = w_empty + 25;                       // [2]
    uint64_t *acc_empty = acc_full + 2;                      // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    unsigned char *stage0 = smem + 640;                                     // [sbufs][a_stride]  dense [rows][cols][Cin]
    unsigned char *halo0 = stage0 + (size_t)p.sbufs * p.a_stride;           // [hbufs][hi, lo][a_stride]  planes
    unsigned char *wring = halo0 + (size_t)p.hbufs * 2 * p.a_stride;        // [wstages][hi, lo][w_bytes]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int a = 0; a < 2; ++a) {
            mbar_init(&h_full[a], 1);
            mbar_init(&s_free[a], 4);
            mbar_init(&h_split[a], 4);
            mbar_init(&h_empty[a], 1);
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        for (int s = 0; s < p.wstages; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int T = p.KS * p.KS;
    const int per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0) {
            auto load_halo = [&](int tile, int use) {  // `use` = how many halo tiles this CTA has requested before
                const int sb = use % p.sbufs;
                const int n = tile / per_img, tt = tile - n * per_img;
                const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
                mbar_wait(&s_free[sb], ((uint32_t)(use / p.sbufs) & 1u) ^ 1u);
                stamp(p, 0, use, 0);
                mbar_arrive_expect_tx(&h_full[sb], (uint32_t)p.a_bytes);
                tma_box_4d(stage0 + (size_t)sb * p.a_stride, &xmap, 0, tx * kHTW - p.pad, ty * kHTH - p.pad, n, &h_full[sb]);
            };
            auto load_tap = [&](int t, int s) {
                mbar_arrive_expect_tx(&w_full[s], (uint32_t)(2 * p.w_bytes));
                bulk_g2s(wring + (size_t)s * 2 * p.w_bytes, reinterpret_cast<const unsigned char *>(p.wpack) + (size_t)t * 2 * p.w_bytes,
                         (uint32_t)(2 * p.w_bytes), &w_full[s]);
            };
            if ((int)blockIdx.x < p.total_tiles) load_halo(blockIdx.x, 0);
            if (p.w_resident)
                for (int t = 0; t < T; ++t) load_tap(t, t);  // the whole filter stays in shared memory: loaded once per CTA
            int it = 0, s = 0;
            uint32_t wpar = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                // the NEXT tile's halo first: its latency (and its split) hide under this tile's MMAs
                if (tile + (int)gridDim.x < p.total_tiles) load_halo(tile + gridDim.x, it + 1);
                if (p.w_resident) continue;
                for (int t = 0; t < T; ++t) {
                    mbar_wait(&w_empty[s], wpar ^ 1u);
                    load_tap(t, s);
                    if (++s == p.wstages) { s = 0; wpar ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------- MMA issuer -------------------------------------------------
        // The WHOLE warp runs the loops, on warp-uniform values only (kernel parameters, blockIdx, loop counters): the
        // compiler then keeps descriptors and addresses in uniform registers and the tcgen05 instructions issue back to back
        // from the elected lane.  (Run under `if (lane == 0)` every MMA was wrapped in an R2UR / ELECT / BRA.U.ANY sequence:
        // 108 cycles per MMA measured by the role trace, twice the tensor pipe's 49.)
        const bool leader = elect_one();
        int it = 0, s = 0;
        uint32_t wpar = 0;
        const int kslices = p.Cin / 8;
        // descriptors as (lo, hi) words: lo = start address >> 4 | LBO >> 4 << 16, hi = SBO >> 4 | version; offsets are added
        // to lo in 16-byte units (shared-memory addresses stay below 2^18, so the 14-bit field never carries)
        const uint32_t a_hiword = ((uint32_t)p.hcols) | (1u << 14);           // SBO = hcols * 16 bytes
        const uint32_t w_hiword = (128u >> 4) | (1u << 14);                   // SBO = 128 bytes
        const uint32_t a_lbo16 = (uint32_t)p.plane_bytes >> 4, w_lbo16 = 2u * (uint32_t)p.Npad;  // next chunk: plane / 2 Npad rows
        const uint32_t a_kstep = 2u * a_lbo16, w_kstep = 2u * w_lbo16;        // one K slice = two chunks
        const uint32_t halo_u32 = smem_u32(halo0), ring_u32 = smem_u32(wring);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int hb = it % p.hbufs, ab = it & 1;
            const uint32_t tacc = tmem + (uint32_t)(ab * 2 * p.Npad);
            mbar_wait(&acc_empty[ab], (((uint32_t)it >> 1) & 1u) ^ 1u);
            if (lane == 0) stamp(p, 1, it, 0);
            mbar_wait(&h_split[hb], (uint32_t)(it / p.hbufs) & 1u);
            if (lane == 0) stamp(p, 1, it, 1);
            tc_fence_after();
            const uint32_t ahi_lo = ((halo_u32 + (uint32_t)(hb * 2 * p.a_stride)) >> 4) | (a_lbo16 << 16);
            const uint32_t alo_lo = ahi_lo + ((uint32_t)p.a_stride >> 4);
            uint32_t first = 0;  // 0 for the very first MMA of the tile (overwrites the accumulator)
            for (int ky = 0; ky < p.KS; ++ky) {
                for (int kx = 0; kx < p.KS; ++kx) {
                    if (!p.w_resident || it == 0) {  // a resident filter is waited for once, during the first tile
                        mbar_wait(&w_full[s], wpar);
                        tc_fence_after();
                    }
                    const uint32_t shift = (uint32_t)((ky * p.dil) * p.hcols + kx * p.dil);
                    const uint32_t ah = ahi_lo + shift, al = alo_lo + shift;
                    const uint32_t w = ((ring_u32 + (uint32_t)(s * 2 * p.w_bytes)) >> 4) | (w_lbo16 << 16);
                    switch (kslices) {  // unrolled: the per-slice descriptor steps become immediates of uniform adds
                    case 1: issue_tap<1>(leader, tacc, ah, al, w, a_kstep, w_kstep, a_hiword, w_hiword, p.idesc, p.idesc2, first); break;
                    case 2: issue_tap<2>(leader, tacc, ah, al, w, a_kstep, w_kstep, a_hiword, w_hiword, p.idesc, p.idesc2, first); break;
                    case 4: issue_tap<4>(leader, tacc, ah, al, w, a_kstep, w_kstep, a_hiword, w_hiword, p.idesc, p.idesc2, first); break;
                    default: issue_tap<8>(leader, tacc, ah, al, w, a_kstep, w_kstep, a_hiword, w_hiword, p.idesc, p.idesc2, first); break;
                    }
                    first = 1;
                    if (p.w_resident) {
                        if (++s == p.wstages) s = 0;  // phase 0 of every w_full completed once and stays complete
                    } else {
                        if (leader) tc_commit(&w_empty[s]);
                        if (++s == p.wstages) { s = 0; wpar ^= 1u; }
                    }
                }
            }
            if (leader) {
                tc_commit(&h_empty[hb]);   // plane set free once every MMA of the tile has read it
                tc_commit(&acc_full[ab]);
            }
            if (lane == 0) stamp(p, 1, it, 2);
        }
    } else if (warp < 6) {
        // ------------------------------------------------ 3xTF32 split ------------------------------------------------
        const int wt = threadIdx.x - 64;  // 0..127
        int it = 0;
        const int P = p.hcols * p.hrows, NC = p.Cin / 4;  // pixels of the halo tile, 16-byte channel chunks per pixel
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int sb = it % p.sbufs, hb = it % p.hbufs;
            mbar_wait(&h_full[sb], (uint32_t)(it / p.sbufs) & 1u);            // dense halo tile landed
            if (wt == 0) stamp(p, 2, it, 0);
            mbar_wait(&h_empty[hb], ((uint32_t)(it / p.hbufs) & 1u) ^ 1u);    // the MMAs that read this plane set are done
            if (wt == 0) stamp(p, 2, it, 1);
            const float4 *src = reinterpret_cast<const float4 *>(stage0 + (size_t)sb * p.a_stride);
            float4 *ahi = reinterpret_cast<float4 *>(halo0 + (size_t)hb * 2 * p.a_stride);
            float4 *alo = reinterpret_cast<float4 *>(halo0 + (size_t)hb * 2 * p.a_stride + p.a_stride);
            for (int px = wt; px < P; px += 128) {
                for (int j = 0; j < NC; ++j) {
                    const int c = (j + px) & (NC - 1);  // rotated by pixel: the 8 lanes of a quarter-warp read 8 bank groups
                    const float4 v = src[px * NC + c];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
                    ahi[c * P + px] = h;  // plane c: 16 bytes per pixel; consecutive lanes, consecutive pixels (P is even)
                    alo[c * P + px] = l;
                }
            }
            proxy_fence_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&h_split[hb]);
                mbar_arrive(&s_free[sb]);
            }
            if (wt == 0) stamp(p, 2, it, 2);
        }
    } else {
        // -------------------------------------------------- epilogue --------------------------------------------------
        const int quad = warp & 3;
        const int row = quad * 32 + lane;  // accumulator row m -> output pixel (y0 + m / 8, x0 + m % 8)
        const bool vec4 = ((p.ycs | p.yco) & 3) == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            mbar_wait(&acc_full[ab], ((uint32_t)it >> 1) & 1u);
            if (threadIdx.x == 192) stamp(p, 3, it, 0);
            tc_fence_after();
            const int n = tile / per_img, tt = tile - n * per_img;
            const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
            const int oy = ty * kHTH + row / kHTW, ox = tx * kHTW + row % kHTW;
            const bool inside = oy < p.Ho && ox < p.Wo;
            float *dst = p.y + (((size_t)n * p.Ho + (inside ? oy : 0)) * p.Wo + (inside ? ox : 0)) * p.ycs + p.yco;
            for (int c0 = 0; c0 < p.Npad; c0 += 16) {
                float v[16], u[16];
                const uint32_t tcol = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * 2 * p.Npad + c0);
                tmem_ld16(tcol, v);                        // a_hi.w_hi + a_lo.w_hi
                tmem_ld16(tcol + (uint32_t)p.Npad, u);     // a_hi.w_lo
                if (c0 + 16 >= p.Npad) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[ab]);
                }
                if (!inside) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int co = c0 + i;
                    float o = (v[i] + u[i]) + ((p.bias && co < p.Cout) ? __ldg(p.bias + co) : 0.0f);
                    v[i] = p.relu ? fmaxf(o, 0.0f) : o;
                }
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const int co = c0 + i;
                    if (co + 3 < p.Cout && vec4) {
                        *reinterpret_cast<float4 *>(dst + co) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (co + e < p.Cout) dst[co + e] = v[i + e];
                    }
                }
            }
            if (threadIdx.x == 192) stamp(p, 3, it, 1);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

long long *g_conv5h_trace = nullptr;

PFN_cuTensorMapEncodeTiled_v12000 encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    }();
    return fn;
}

// Opt a kernel in to the device's maximum dynamic shared memory, once per (kernel, device).  Always the SAME value, so
// concurrent host threads (one per GPU, or several streams of one GPU) can only repeat an idempotent call -- a per-thread or
// per-size high-water mark let one thread lower the attribute under another thread's larger launch.
template <typename Kern>
bool optin_max_smem(Kern kern, std::atomic<unsigned long long> &done, int dev, int smem_optin) {
    if (dev >= 0 && dev < 64 && ((done.load(std::memory_order_relaxed) >> dev) & 1ull)) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_relaxed);
    return true;
}

int npad_of(int cout) { return (cout + 15) / 16 * 16; }

}  // namespace

extern "C" {

// 1 when pmb200_conv2d_tc5 takes this layer
int pmb200_conv2d_tc5_supported(int Cin, int Cout, int KS, int stride) {
    const bool cin_ok = Cin == 8 || Cin == 16 || Cin == 32 || Cin == 64;
    return (cin_ok && Cout >= 1 && Cout <= 64 && (KS == 1 || KS == 3 || KS == 5) && (stride == 1 || stride == 2)) ? 1 : 0;
}

// floats of the packed filter: [tap][Cin/32 blocks][hi rows Npad | lo rows Npad][min(Cin,32)] (ops.pack_conv_filter_tc5)
int pmb200_conv2d_tc5_filter_floats(int Cin, int Cout, int KS) {
    if (!pmb200_conv2d_tc5_supported(Cin, Cout, KS, 1)) return -1;
    return KS * KS * 2 * npad_of(Cout) * Cin;
}

int pmb200_conv2d_tc5(const float *x_nhwc, const float *filter_tc5, const float *bias, float *y_nhwc, int N, int H, int W, int Cin,
                      int Cout, int KS, int stride, int pad, int dil, int relu, int y_channel_stride, int y_channel_offset,
                      void *stream) {
    if (!x_nhwc || !filter_tc5 || !y_nhwc) return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5: null pointer");
    if (!pmb200_conv2d_tc5_supported(Cin, Cout, KS, stride))
        return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5: Cin in {8,16,32,64}, Cout <= 64, KS in {1,3,5}, stride 1 or 2");
    if (N < 1 || H < 1 || W < 1 || pad < 0 || dil < 1) return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5: bad size");
    if ((reinterpret_cast<uintptr_t>(x_nhwc) & 15u) || (reinterpret_cast<uintptr_t>(filter_tc5) & 15u))
        return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5: input and filter must be 16-byte aligned");
    const int Ho = (H + 2 * pad - dil * (KS - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dil * (KS - 1) - 1) / stride + 1;
    if (Ho < 1 || Wo < 1) return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5: empty output");
    if (y_channel_stride < Cout + y_channel_offset || y_channel_offset < 0)
        return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5: output channel slice out of range");
    auto enc = encoder();
    if (!enc) return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5: cuTensorMapEncodeTiled unavailable");

    Conv5Params p;
    p.wpack = filter_tc5; p.bias = bias; p.y = y_nhwc;
    p.N = N; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.Cin = Cin; p.Cout = Cout; p.Npad = npad_of(Cout);
    p.KS = KS; p.S = stride; p.pad = pad; p.dil = dil; p.relu = relu; p.ycs = y_channel_stride; p.yco = y_channel_offset;
    p.tiles_x = (Wo + kTW - 1) / kTW; p.tiles_y = (Ho + kTH - 1) / kTH;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * N;
    if (tiles > 0x7fffffffLL) return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5: too many tiles");
    p.total_tiles = (int)tiles;
    const int cblk = Cin < 32 ? Cin : 32;
    p.RB = cblk * 4;
    p.KB = Cin / cblk;
    p.a_bytes = 128 * p.RB * p.KB;
    p.w_bytes = p.Npad * p.RB * p.KB;
    p.stage_bytes = 2 * p.a_bytes + 2 * p.w_bytes;  // multiples of 1024 except tiny layers: keep every stage 1024-aligned
    p.stage_bytes = (p.stage_bytes + 1023) / 1024 * 1024;
    p.layout_type = p.RB == 128 ? 2u : (p.RB == 64 ? 4u : 6u);
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);  // F32 += TF32 . TF32, K-major both
    p.idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * p.Npad >> 3) << 17) | ((128u >> 4) << 24);
    p.tmem_cols = 4 * p.Npad <= 64 ? 64u : (4 * p.Npad <= 128 ? 128u : 256u);  // two accumulators of 2 Npad fp32 columns

    int dev = 0, sms = 0, smem_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // two CTAs per SM when the ring allows at least 3 stages each, else one CTA with as many stages as fit (<= 8)
    int ctas = 2;
    int stages = ((smem_optin + 1024) / 2 - 1024 - 2048) / p.stage_bytes;
    if (stages < 3) {
        ctas = 1;
        stages = (smem_optin - 2048) / p.stage_bytes;
    }
    if (stages > 8) stages = 8;
    if (stages < 2) return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5: a ring of two stages does not fit in shared memory");
    p.stages = stages;
    const int smem = 1024 + stages * p.stage_bytes + 1024;  // + slack for the 1024-byte alignment of the dynamic segment
    static std::atomic<unsigned long long> optin_done{0};
    if (!optin_max_smem(conv5_kernel, optin_done, dev, smem_optin))
        return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5: shared-memory opt-in failed");
    CUtensorMap xmap;
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
    const cuuint32_t box[4] = {(cuuint32_t)cblk, (cuuint32_t)(kTW * stride), (cuuint32_t)(kTH * stride), 1u};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
    const CUtensorMapSwizzle sw = p.RB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.RB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    if (enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x_nhwc), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5: tensor map rejected (alignment / size)");
    long long grid = (long long)sms * ctas;
    if (grid > tiles) grid = tiles;
    conv5_kernel<<<(unsigned)grid, kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p, xmap);
    return pmb200_internal_launch_status("conv2d_tc5");
}

// Debugging aid: device buffer of 4 roles x 16 tiles x 4 events (256 int64) that the next K-D5h launches stamp with clock64
// for CTA 0 (producer / MMA issuer / split / epilogue); nullptr switches it off.  Not part of the reference-facing surface.
int pmb200_debug_conv5h_trace(long long *device_buffer_256) {
    g_conv5h_trace = device_buffer_256;
    return 0;
}

// Stride-1 form (K-D5h): one halo tile per 8 x 16 output tile, filter taps as shifted views.  Filter image:
// [tap][Cin/4 chunks][hi rows Npad | lo rows Npad][4 floats] (ops.pack_conv_filter_tc5h), same float count as the per-tap form.
int pmb200_conv2d_tc5h(const float *x_nhwc, const float *filter_tc5h, const float *bias, float *y_nhwc, int N, int H, int W, int Cin,
                       int Cout, int KS, int pad, int dil, int relu, int y_channel_stride, int y_channel_offset, void *stream) {
    if (!x_nhwc || !filter_tc5h || !y_nhwc) return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5h: null pointer");
    if (!pmb200_conv2d_tc5_supported(Cin, Cout, KS, 1))
        return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5h: Cin in {8,16,32,64}, Cout <= 64, KS in {1,3,5}");
    if (N < 1 || H < 1 || W < 1 || pad < 0 || dil < 1) return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5h: bad size");
    if ((reinterpret_cast<uintptr_t>(x_nhwc) & 15u) || (reinterpret_cast<uintptr_t>(filter_tc5h) & 15u))
        return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5h: input and filter must be 16-byte aligned");
    const int Ho = H + 2 * pad - dil * (KS - 1), Wo = W + 2 * pad - dil * (KS - 1);
    if (Ho < 1 || Wo < 1) return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5h: empty output");
    if (y_channel_stride < Cout + y_channel_offset || y_channel_offset < 0)
        return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5h: output channel slice out of range");
    auto enc = encoder();
    if (!enc) return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5h: cuTensorMapEncodeTiled unavailable");
    Conv5hParams p;
    p.wpack = filter_tc5h; p.bias = bias; p.y = y_nhwc;
    p.N = N; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.Cin = Cin; p.Cout = Cout; p.Npad = npad_of(Cout);
    p.KS = KS; p.pad = pad; p.dil = dil; p.relu = relu; p.ycs = y_channel_stride; p.yco = y_channel_offset;
    p.trace = g_conv5h_trace;
    p.tiles_x = (Wo + kHTW - 1) / kHTW; p.tiles_y = (Ho + kHTH - 1) / kHTH;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * N;
    if (tiles > 0x7fffffffLL) return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5h: too many tiles");
    p.total_tiles = (int)tiles;
    p.hcols = kHTW + dil * (KS - 1); p.hrows = kHTH + dil * (KS - 1);
    if (p.hcols > 256 || p.hrows > 256) return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5h: dilation too large for one TMA box");
    p.plane_bytes = p.hrows * p.hcols * 16;
    p.a_bytes = p.plane_bytes * (Cin / 4);
    p.w_bytes = p.Npad * Cin * 4;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
    p.idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * p.Npad >> 3) << 17) | ((128u >> 4) << 24);
    p.tmem_cols = 4 * p.Npad <= 64 ? 64u : (4 * p.Npad <= 128 ? 128u : 256u);  // two accumulators of 2 Npad fp32 columns
    if ((unsigned)p.plane_bytes >= (1u << 18) || p.hcols * 16 >= (1 << 18))
        return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5h: halo tile exceeds the descriptor's 14-bit offsets");
    int dev = 0, sms = 0, smem_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    p.a_stride = (p.a_bytes + 127) / 128 * 128;
    // Shared memory: barriers (640 B), staging buffers, plane sets, filter stages, 128 B alignment slack.  Preference order:
    // whole filter resident with both double buffers; resident with one plane set; else a ring as deep as fits (>= 2 stages
    // wanted) with the buffers reduced in the order plane sets, staging.  Two CTAs per SM when everything fits twice.
    const int T = KS * KS;
    auto fixed_of = [&](int sb, int hb) { return 640 + sb * p.a_stride + hb * 2 * p.a_stride + 128; };
    p.sbufs = 2; p.hbufs = 2; p.w_resident = 0;
    int wst = 0;
    if (fixed_of(2, 2) + T * 2 * p.w_bytes <= smem_optin) { p.w_resident = 1; wst = T; }
    else if (fixed_of(2, 1) + T * 2 * p.w_bytes <= smem_optin) { p.hbufs = 1; p.w_resident = 1; wst = T; }
    else {
        const int want_w = T < 4 ? T : 4;
        if (fixed_of(2, 2) + want_w * 2 * p.w_bytes > smem_optin) p.hbufs = 1;
        if (fixed_of(2, p.hbufs) + want_w * 2 * p.w_bytes > smem_optin) p.sbufs = 1;
        wst = (smem_optin - fixed_of(p.sbufs, p.hbufs)) / (2 * p.w_bytes);
        if (wst > 8) wst = 8;
        if (wst > T) wst = T;
        if (wst < 1) return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5h: halo tile + filter ring do not fit in shared memory");
    }
    const int fixed = fixed_of(p.sbufs, p.hbufs);
    int ctas = 1;
    if (2 * (fixed + wst * 2 * p.w_bytes + 1024) <= smem_optin + 1024) ctas = 2;
    p.wstages = wst;
    const int smem = fixed + wst * 2 * p.w_bytes;
    static std::atomic<unsigned long long> optin_done{0};
    if (!optin_max_smem(conv5h_kernel, optin_done, dev, smem_optin))
        return pmb200_internal_fail(PMB200_EUNSUPPORTED, "conv2d_tc5h: shared-memory opt-in failed");
    CUtensorMap xmap;
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
    const cuuint32_t box[4] = {(cuuint32_t)Cin, (cuuint32_t)p.hcols, (cuuint32_t)p.hrows, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    if (enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x_nhwc), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return pmb200_internal_fail(PMB200_EINVAL, "conv2d_tc5h: tensor map rejected (alignment / size)");
    long long grid = (long long)sms * ctas;
    if (grid > tiles) grid = tiles;
    conv5h_kernel<<<(unsigned)grid, kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p, xmap);
    return pmb200_internal_launch_status("conv2d_tc5h");
}

}  // extern "C"
