// Micro-benchmark: cost of one tcgen05.mma (kind::tf32, M = 128, K = 8) as a function of N, of the number of accumulators the
// stream alternates between, and of the operand layout (no swizzle / 128-byte swizzle), issued back to back by one thread.
// Answers the question the K-D5 / K-D5h convs raise: is a stream of small-N MMAs bound by the tensor pipe (work ~ N) or by a
// fixed per-instruction cost?  Operands are zeros in shared memory (the arithmetic result is irrelevant).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate tools/umma_rate.cu && /tmp/umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int KIND>  // 0: tf32, 1: f16 (bf16 operands)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(
                         tmem_d),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
                         tmem_d),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                     : "memory");
}
__device__ __forceinline__ uint64_t desc_plain(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((1024 >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int KIND>
__global__ void __launch_bounds__(128) rate_kernel(int N, int M, int nacc, int reps, int swz, long long *out, int a_sbo = 128, int a_lbo = 2048,
                                                   int a_shift = 0) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 64);
    unsigned char *A = smem + 1024, *B = A + 32768;
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(A)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t kind_bits = KIND == 0 ? ((2u << 7) | (2u << 10)) : ((1u << 7) | (1u << 10));  // tf32 / bf16 operands
    const uint32_t idesc = (1u << 4) | kind_bits | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (threadIdx.x == 0) {
        // plain: core matrices 8 rows x 16 B; K-adjacent core matrix 128*16 B away (A) / N*16 B away (B); next 8 rows 128 B away
        const uint64_t da = swz ? desc_sw128(smem_u32(A)) : desc_plain(smem_u32(A) + (uint32_t)a_shift, (uint32_t)a_lbo, (uint32_t)a_sbo);
        const uint64_t db = swz ? desc_sw128(smem_u32(B)) : desc_plain(smem_u32(B), (uint32_t)N * 16, 128);
        uint32_t par = 0;
        for (int warm = 0; warm < 2; ++warm) {
            const long long t0 = clock64();
            const uint32_t d0 = tmem, d1 = tmem + (uint32_t)((nacc - 1) * N);  // the issue loop carries no address arithmetic
#pragma unroll 1
            for (int r = 0; r < reps; r += 8) {
                tc_mma<KIND>(d0, da, db, idesc, 1); tc_mma<KIND>(d1, da, db, idesc, 1);
                tc_mma<KIND>(d0, da, db, idesc, 1); tc_mma<KIND>(d1, da, db, idesc, 1);
                tc_mma<KIND>(d0, da, db, idesc, 1); tc_mma<KIND>(d1, da, db, idesc, 1);
                tc_mma<KIND>(d0, da, db, idesc, 1); tc_mma<KIND>(d1, da, db, idesc, 1);
            }
            tc_commit(bar);
            const long long t1 = clock64();
            while (!mbar_try_wait(bar, par)) {}
            par ^= 1u;
            const long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long *out;
    cudaMalloc(&out, 16);
    const int smem = 1024 + 32768 + 65536 + 1024;
    cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 2048;
    printf("{\"reps\": %d, \"rows\": [\n", reps);
    bool first = true;
    for (int kind = 0; kind < 2; ++kind)
        for (int swz = 0; swz < 2; ++swz)
            for (int grid : {1, 148})
                for (int N : {16, 32, 64, 128, 256})
                  for (int M : {64, 128})
                    for (int nacc : {1, 2}) {
                        if (nacc * N > 512) continue;
                        if (kind == 0)
                            rate_kernel<0><<<grid, 128, smem>>>(N, M, nacc, reps, swz, out);
                        else
                            rate_kernel<1><<<grid, 128, smem>>>(N, M, nacc, reps, swz, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        long long h[2] = {0, 0};
                        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                        printf("%s {\"kind\": \"%s\", \"swizzle128\": %d, \"ctas\": %d, \"M\": %d, \"N\": %d, \"accumulators\": %d, \"issue_cycles_per_mma\": %.1f, "
                               "\"complete_cycles_per_mma\": %.1f, \"err\": \"%s\"}",
                               first ? "" : ",\n", kind == 0 ? "tf32 K8" : "bf16 K16", swz, grid, M, N, nacc, (double)h[0] / reps, (double)h[1] / reps,
                               e == cudaSuccess ? "" : cudaGetErrorString(e));
                        first = false;
                        if (e != cudaSuccess) { printf("]}\n"); return 1; }
                    }
    // the K-D5h A operand: core matrices (8 pixels x 16 B) of a halo tile -- row groups hcols*16 = 160 B apart (not 128 B
    // aligned), planes 2880 B apart, tap shifts in 16 B steps
    for (int N : {32, 64, 128})
        for (int geo = 0; geo < 5; ++geo) {
            const int sbo = geo == 0 ? 128 : (geo == 3 ? 256 : 160), lbo = geo == 0 ? 2048 : (geo == 3 ? 4096 : 2880), shift = geo == 2 ? 176 : (geo == 4 ? 64 : 0);
            rate_kernel<0><<<148, 128, smem>>>(N, 128, 1, reps, 0, out, sbo, lbo, shift);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2] = {0, 0};
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf(",\n {\"kind\": \"tf32 K8\", \"a_operand\": {\"sbo\": %d, \"lbo\": %d, \"start_shift\": %d}, \"ctas\": 148, \"M\": 128, \"N\": %d, "
                   "\"complete_cycles_per_mma\": %.1f, \"err\": \"%s\"}", sbo, lbo, shift, N, (double)h[1] / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    printf("]}\n");
    return 0;
}
