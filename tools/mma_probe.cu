// mma_probe.cu -- microbenchmark: cycles per tcgen05.mma.kind::tf32 for different smem layouts / shapes.
// Data is garbage; only timing matters.   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_probe mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long mkdesc(unsigned saddr, unsigned lbo, unsigned sbo, unsigned layout) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fffu);
    d |= (unsigned long long)((lbo >> 4) & 0x3fffu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;
    d |= (unsigned long long)layout << 61;
    return d;
}
struct P { int M, N, layout, a_lbo, a_sbo, b_lbo, b_sbo, iters, ntiles, kadv, a_step, b_step, unroll; };
__global__ void __launch_bounds__(128, 1) probe(P p, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned long long bar;
    __shared__ unsigned tmem_ptr;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((float*)sm)[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_ptr;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(p.N >> 3) << 17) | ((unsigned)(p.M >> 4) << 24);
        const unsigned long long a0 = mkdesc(smem_u32(sm), p.a_lbo, p.a_sbo, p.layout);
        const unsigned long long b0 = mkdesc(smem_u32(sm + 96 * 1024), p.b_lbo, p.b_sbo, p.layout);
        t0 = clock64();
        if (p.unroll) {
            const unsigned np = (unsigned)p.N;
            for (int it = 0; it < p.iters * p.ntiles / 8; ++it) {
                const unsigned long long ad = a0 + (unsigned)((it & 7) * p.a_step * 4);
                const unsigned long long bd = b0 + (unsigned)((it & 7) * p.b_step);
                const unsigned long long astep = (unsigned)p.a_step;
                asm volatile(
                    "{\n\t.reg .pred pe;\n\t.reg .b64 a1, a2, a3;\n\t.reg .b32 d1, d2, d3;\n\t"
                    "elect.sync _|pe, 0xffffffff;\n\t"
                    "add.u64 a1, %1, %5;\n\tadd.u64 a2, a1, %5;\n\tadd.u64 a3, a2, %5;\n\t"
                    "add.u32 d1, %0, %4;\n\tadd.u32 d2, d1, %4;\n\tadd.u32 d3, d2, %4;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [d1], a1, %2, %3, 1;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [d2], a2, %2, %3, 1;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [d3], a3, %2, %3, 1;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [d1], a1, %2, %3, 1;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [d2], a2, %2, %3, 1;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [d3], a3, %2, %3, 1;\n\t}"
                    ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(p.ntiles >= 4 ? np : 0u), "l"(astep)
                    : "memory");
            }
        } else
        for (int it = 0; it < p.iters; ++it) {
            for (int t = 0; t < p.ntiles; ++t) {
                const unsigned long long ad = a0 + (unsigned)((it & 3) * p.kadv + t * p.a_step + (it & 7) * p.a_step * 4);
                const unsigned long long bd = b0 + (unsigned)((it & 3) * p.kadv + (it & 7) * p.b_step);
                asm volatile(
                    "{\n\t.reg .pred pe;\n\t"
                    "elect.sync _|pe, 0xffffffff;\n\t"
                    "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;\n\t}" ::"r"(tmem + t * p.N), "l"(ad), "l"(bd), "r"(idesc)
                    : "memory");
            }
        }
        asm volatile(
            "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        }
        t1 = clock64();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { const char* name; P p; } cases[] = {
        // M, N, layout, a_lbo, a_sbo, b_lbo, b_sbo, iters, ntiles, kadv, a_step, b_step (16B units), unroll
        {"loop    M128 N64", {128, 64, 0, 9792, 544, 1024, 128, 512, 4, 0, 8, 64, 0}},
        {"unroll8 M128 N32", {128, 32, 0, 9792, 544, 512, 128, 512, 4, 0, 8, 64, 1}},
        {"unroll8 M128 N64", {128, 64, 0, 9792, 544, 1024, 128, 512, 4, 0, 8, 64, 1}},
        {"unroll8 M128 N128", {128, 128, 0, 9792, 544, 2048, 128, 512, 4, 0, 8, 64, 1}},
        {"unroll8 M128 N256 (2 tiles)", {128, 256, 0, 9792, 544, 4096, 128, 512, 2, 0, 8, 64, 1}},
        {"unroll8 M128 N64 same accumulator", {128, 64, 0, 9792, 544, 1024, 128, 512, 1, 0, 8, 64, 1}},
        {"unroll8 M64 N64", {64, 64, 0, 9792, 544, 1024, 128, 512, 4, 0, 8, 64, 1}},
        {"unroll8 M64 N256 (2 tiles)", {64, 256, 0, 9792, 544, 4096, 128, 512, 2, 0, 8, 64, 1}},
    };
    for (auto& c : cases) {
        probe<<<148, 128, 200 * 1024>>>(c.p, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const double n = (double)c.p.iters * c.p.ntiles;
        printf("%-62s %s  %8.1f cyc/MMA  (%.0f%% of tf32 peak)\n", c.name, e == cudaSuccess ? "ok " : cudaGetErrorString(e), cyc / n,
               100.0 * (c.p.M * (double)c.p.N * 8 / 2048.0) / (cyc / n));
    }
    return 0;
}
