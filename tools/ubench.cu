// ubench.cu -- issue-rate microbenchmarks for the instruction classes the extraction kernels lean on (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
// Prints warp-instructions per clock per SM sub-partition (SMSP) for each class, 16 warps per SMSP resident.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 256
#define UNR 16

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, long long* cyc, uint32_t seed)
{
    uint32_t a[UNR], b = seed * 3 + threadIdx.x, c = seed * 7 + 1;
    unsigned long long A[UNR], B2 = ((unsigned long long)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f), C2 = B2 + 5;
#pragma unroll
    for (int i = 0; i < UNR; ++i) { a[i] = threadIdx.x + i * seed; A[i] = B2 + i * (unsigned long long)seed; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < UNR; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B2), "l"(C2));
            if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B2));
            if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 8) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 9) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+r"(a[i]) : "r"(c));
            if (OP == 10) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 11) a[i] = __vimax3_u16x2(a[i], b, c);
            if (OP == 12) a[i] = __vmaxu2(a[i], b);
            if (OP == 13) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B2));
            if (OP == 14) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(i + 8) % UNR]) : "r"(b), "r"(c)); }
            if (OP == 15) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B2), "l"(C2)); asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < UNR; ++i) s += a[i] + (uint32_t)A[i] + (uint32_t)(A[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter)
{
    uint32_t* out; long long* cyc;
    const int blocks = 148 * 8;
    cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&cyc, blocks * 8);
    k<OP><<<blocks, 256>>>(out, cyc, 1); k<OP><<<blocks, 256>>>(out, cyc, 2);
    cudaDeviceSynchronize();
    static long long h[148 * 8];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
    // per SMSP: 8 blocks x 8 warps / 4 = 16 warps (<= 64 registers / thread), each issuing ITER * UNR * per_iter instructions
    const double inst = 16.0 * ITER * UNR * per_iter;
    printf("%-28s %8.0f cycles  %.3f warp-inst/clk/SMSP\n", name, avg, inst / avg);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA (3 reg)", 1);
    run<9>("FFMA (imm)", 1);
    run<10>("FADD", 1);
    run<1>("FFMA2", 1);
    run<2>("FADD2", 1);
    run<13>("FMUL2", 1);
    run<3>("IMAD", 1);
    run<8>("DP4A", 1);
    run<4>("PRMT", 1);
    run<6>("LOP3", 1);
    run<7>("IADD", 1);
    run<11>("VIMNMX3.U16x2", 1);
    run<12>("VIMNMX.U16x2", 1);
    run<14>("FFMA + LOP3 (pair)", 2);
    run<15>("FFMA2 + PRMT (pair)", 2);
    return 0;
}
