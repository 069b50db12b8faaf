// Micro-benchmark: issue rate of tcgen05.mma.cta_group::2 (M = 256 over the CTA pair, N = 256) from shared-memory operands,
// kind::f16 (bf16, K = 16) against kind::f8f6f4 (e4m3, K = 32), with nothing else going on -- no TMA, no epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ab/mma_rate tools/ab/mma_rate.cu && tools/ab/mma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
    uint64_t d = (uint64_t)((a & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int stages, int random_data, const uint8_t* src, int copy_stream, int chain) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < stages * 32768 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        // finite, moderate values in either format: every byte in 0x30..0x4F with a random sign bit
        reinterpret_cast<uint32_t*>(smem)[i] = random_data ? ((h & 0x9F9F9F9Fu & 0x8F8F8F8Fu) | 0x30303030u | (h & 0x0F0F0F0Fu)) : 0u;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    __shared__ uint64_t cbar[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&cbar[s])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (copy_stream && threadIdx.x == 32) {
        // what the TMA producer does to the shared-memory port: 32 KB per stage streamed in from L2, `stages` copies in flight
        const uint8_t* from = src + (size_t)blockIdx.x * 262144;
        const uint32_t n_copies = (uint32_t)iters;          // one stage per four MMAs, as in the tower
        for (uint32_t it = 0; it < n_copies + (uint32_t)stages; ++it) {
            const int s = it % stages;
            if (it >= (uint32_t)stages) {
                uint32_t ok;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&cbar[s])), "r"(((it / stages) - 1) & 1) : "memory");
                } while (!ok);
            }
            if (it >= n_copies) continue;                   // drain: every copy has landed before the kernel ends
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[s])), "r"(32768u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + s * 32768)), "l"(from + (it % 8) * 32768), "r"(32768u), "r"(smem_u32(&cbar[s])) : "memory");
        }
    }
    if (rank == 0 && threadIdx.x == 0) {
        // chain 0: the accumulator changes every four instructions; 1: every instruction adds to ONE accumulator (what a tower
        // item does: one dependent chain); 2: N = 128 halves, two interleaved chains on the two column halves of one accumulator
        const uint32_t n = chain == 2 ? 128u : 256u;
        const uint32_t idesc = KIND == 0 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24))
                                         : ((1u << 4) | ((n >> 3) << 17) | ((256u >> 4) << 24));
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t a = smem_u32(smem + (i % stages) * 32768);
            const uint64_t da = umma_desc(a), db = umma_desc(a + 16384);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (chain == 2) {
                    mma<KIND>(tmem, da + 2 * kk, db + 2 * kk, idesc, 1u);
                    mma<KIND>(tmem + 128, da + 2 * kk, db + 2 * kk + (8192 >> 4), idesc, 1u);     // the other 64 rows of this CTA's B half
                } else mma<KIND>(tmem + (chain == 1 ? 0 : (i & 1) * 256), da + 2 * kk, db + 2 * kk, idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
        out[blockIdx.x / 2] = clock64() - t0;
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int KIND>
void run(const char* name, int grid, int stages, int random_data, int copy_stream, int chain) {
    static uint8_t* src = nullptr;
    if (!src) { cudaMalloc(&src, (size_t)148 * 262144); cudaMemset(src, 0x3c, (size_t)148 * 262144); }
    long long* d;
    cudaMalloc(&d, 128 * sizeof(long long));
    const int iters = 4096, smem = stages * 32768 + 2048;
    cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = getenv("MMA_SUSTAIN") ? 600 : 2;          // 600 launches of ~1.2 ms back to back: long enough for the power cap to act
    for (int rep = 0; rep < reps; ++rep) {
        rate_kernel<KIND><<<grid, 128, smem>>>(d, iters, stages, random_data, src, copy_stream, chain);
        if (rep + 1 < reps && reps > 2) continue;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    long long h[128];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid / 2; ++i) avg += (double)h[i];
    avg /= grid / 2;
    printf("%-28s %s grid %3d, %d stages: %.1f cycles per K step (256x256xK over the pair), %.0f FLOP/cycle/SM\n", name, chain == 0 ? "accumulator changes every 4," : (chain == 1 ? "one accumulator throughout," : "two N=128 chains interleaved,"), grid, stages,
           avg / (iters * 4.0), (KIND == 0 ? 16.0 : 32.0) * 256 * 256 * 2 / 2 / (avg / (iters * 4.0)));
    cudaFree(d);
}

int main() {
    for (int chain : {1}) {
        run<0>("kind::f16 (bf16, K=16)", 148, 6, 1, 1, chain);
        run<1>("kind::f8f6f4 (e4m3, K=32)", 148, 6, 1, 1, chain);
    }
    return 0;
}
