// umma_peak.cu -- the int8 tensor-pipe peak of THIS GPU as a pure tcgen05.mma.kind::i8 issue loop ("measure a pure-UMMA i8
// loop as own peak"): the denominator of the matcher's roofline (bench.py reads the committed result, profiles/r2_int8_peak.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_peak.bin tools/umma_peak.cu && tools/umma_peak.bin [seconds]
// One CTA per SM (148), one elected thread issues MMAs back to back into two alternating TMEM accumulators; operands are whatever
// bytes sit in shared memory (throughput does not depend on the values).  Variants: A from shared memory (SS) / from tensor memory
// (TS, what the matcher uses), N = 256 / N = 96 (the matcher's tile).  Prints TOP/s for a short burst and for a seconds-long run
// (clocks settle under the power cap), as JSON on the last line.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {      // K-major SWIZZLE_128B, 8-row groups 1024 B apart
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k_peak(int iters, unsigned long long* sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = base;                       // A: 128 rows x 128 B (4 k-steps of 32)
    uint8_t* sb = base + 128 * 128;           // B: 256 rows x 128 B
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x01FF01FFu * (i | 1);
    const uint32_t bar_s = smem_u32(&bar);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint64_t da = desc_sw128(smem_u32(sa)), db = desc_sw128(smem_u32(sb));
        constexpr uint32_t NACC = (TS && N > 224) ? 1u : 2u;   // 512 TMEM columns: accumulators + (TS) 32 columns of A
        const uint32_t ta = tm + NACC * N;    // TS: the A operand lives in tensor memory behind the accumulators (128 x 128 B = 32 columns)
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tm + ((uint32_t)it & (NACC - 1u)) * N;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (TS)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(d), "r"(ta + 8u * k), "l"(db + 2u * k), "r"(IDESC), "r"(k) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(da + 2u * k), "l"(db + 2u * k), "r"(IDESC), "r"(k) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_s), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    if (sink && threadIdx.x == 0 && iters < 0) sink[blockIdx.x] = bar;
}

template <int N, bool TS>
double run(int iters, double seconds, double* sustained)
{
    const int smem = (128 + 256) * 128 + 2048;
    cudaFuncSetAttribute(k_peak<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double ops = 148.0 * iters * 4.0 * 2.0 * 128.0 * N * 32.0;
    k_peak<N, TS><<<148, 128, smem>>>(iters, nullptr);
    k_peak<N, TS><<<148, 128, smem>>>(iters, nullptr);
    double best = 0;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k_peak<N, TS><<<148, 128, smem>>>(iters, nullptr);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ops / (ms * 1e-3) / 1e12 > best ? ops / (ms * 1e-3) / 1e12 : best;
    }
    if (sustained) {
        float ms1; cudaEventRecord(e0); k_peak<N, TS><<<148, 128, smem>>>(iters, nullptr); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
        const int reps = (int)(seconds * 1e3 / ms1) + 1;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) k_peak<N, TS><<<148, 128, smem>>>(iters, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        *sustained = ops * reps / (ms * 1e-3) / 1e12;
    }
    return best;
}

int main(int argc, char** argv)
{
    const double seconds = argc > 1 ? atof(argv[1]) : 3.0;
    const int iters = 20000;
    double s_ss256 = 0, s_ts96 = 0;
    const double ss256 = run<256, false>(iters, seconds, &s_ss256);
    const double ts256 = run<256, true>(iters, 0, nullptr);
    const double ss96 = run<96, false>(iters, 0, nullptr);
    const double ts96 = run<96, true>(iters, seconds, &s_ts96);
    printf("tcgen05.mma.cta_group::1.kind::i8 M128 K32, 148 CTAs, 80 000 MMAs each, best of 5 (burst) / %.0f s back to back (sustained)\n", seconds);
    printf("  A in smem, N=256: %.0f TOP/s burst, %.0f sustained\n  A in TMEM, N=256: %.0f TOP/s burst\n  A in smem, N=96:  %.0f TOP/s burst\n"
           "  A in TMEM, N=96:  %.0f TOP/s burst, %.0f sustained\n", ss256, s_ss256, ts256, ss96, ts96, s_ts96);
    printf("{\"int8_tops_burst\": %.1f, \"int8_tops_sustained\": %.1f, \"int8_tops_burst_ts_n256\": %.1f, \"int8_tops_burst_ss_n96\": %.1f, "
           "\"int8_tops_burst_ts_n96\": %.1f, \"int8_tops_sustained_ts_n96\": %.1f, \"how\": \"tools/umma_peak.cu: pure tcgen05.mma.kind::i8 issue loop, cta_group::1, M128 K32, 148 CTAs\"}\n",
           ss256, s_ss256, ts256, ss96, ts96, s_ts96);
    return 0;
}
