// ubench2.cu -- calibrated issue-rate microbenchmarks (B200, sm_100a): which pipe an instruction class occupies and how
// two classes share the issue slots.  One 1024-thread CTA per SM (8 warps per SM sub-partition), 8 independent chains per
// thread, clock64() around the loop; prints cycles per warp-instruction per SMSP (1.0 = one issue slot per instruction,
// 2.0 = a half-rate pipe) for each class alone and for pairs (pair ~ max of the two = different pipes, ~ sum = same pipe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench2 tools/ubench2.cu && /tmp/ubench2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITER 512
#define UNR 8

enum { FFMA, IMAD, IMADI, LOP3, PRMT, IADD3, SHF, VMNMX, VMNMX3, HMNMX2, HSET2, HADD2, HFMA2, VABSD4, VIADD16, FMNMX, IMNMX, POPC, IDP4A, LEA, ISETPSEL, VOTE, FLO, NOPS };
const char* const kNames[] = {"FFMA", "IMAD", "IMAD.imm", "LOP3", "PRMT", "IADD3", "SHF", "VIMNMX.U16x2", "VIMNMX3.U16x2", "HMNMX2", "HSET2", "HADD2", "HFMA2",
                              "VABSDIFF4", "VIADD.16x2", "FMNMX", "IMNMX", "POPC", "IDP4A", "LEA", "ISETP+SEL", "VOTE", "FLO"};

template <int OP>
__device__ __forceinline__ void op(uint32_t& a, uint32_t b, uint32_t c)
{
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == IMADI) asm volatile("mad.lo.u32 %0, %0, 0xFFFF0001, %1;" : "+r"(a) : "r"(c));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));
    if (OP == IADD3) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c));
    if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == VMNMX) a = __vmaxu2(a, b);
    if (OP == VMNMX3) a = __vimax3_u16x2(a, b, c);
    if (OP == HMNMX2) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == HSET2) asm volatile("set.gt.u32.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == VABSD4) a = __vabsdiffu4(a, b);
    if (OP == VIADD16) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == IMNMX) asm volatile("max.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == POPC) asm volatile("{ .reg .u32 t; popc.b32 t, %0; xor.b32 %0, t, %1; }" : "+r"(a) : "r"(b));
    if (OP == IDP4A) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == LEA) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(a) : "r"(b));
    if (OP == ISETPSEL) asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(a) : "r"(b), "r"(c));
    if (OP == VOTE) asm volatile("{ .reg .pred p; .reg .u32 t; setp.gt.u32 p, %0, %1; vote.sync.ballot.b32 t, p, 0xffffffff; xor.b32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c));
    if (OP == FLO) asm volatile("{ .reg .u32 t; bfind.u32 t, %0; xor.b32 %0, t, %1; }" : "+r"(a) : "r"(b));
}

template <int OP1, int OP2>
__global__ void __launch_bounds__(1024) k(uint32_t* out, long long* cyc, uint32_t seed)
{
    uint32_t a[UNR], d[UNR], b = seed * 3 + threadIdx.x, c = seed * 7 + 1;
#pragma unroll
    for (int i = 0; i < UNR; ++i) { a[i] = threadIdx.x + i * seed; d[i] = a[i] * 5 + 1; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < UNR; ++i) {
            op<OP1>(a[i], b, c);
            if (OP2 != NOPS) op<OP2>(d[i], c, b);
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < UNR; ++i) s += a[i] + d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP1, int OP2>
double run()
{
    uint32_t* out; long long* cyc;
    const int blocks = 148;
    cudaMalloc(&out, blocks * 1024 * 4); cudaMalloc(&cyc, blocks * 8);
    k<OP1, OP2><<<blocks, 1024>>>(out, cyc, 1); k<OP1, OP2><<<blocks, 1024>>>(out, cyc, 2);
    cudaDeviceSynchronize();
    static long long h[148];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
    cudaFree(out); cudaFree(cyc);
    return avg / (8.0 * ITER * UNR);          // cycles per loop slot (one OP1 [+ one OP2]) per SMSP
}

template <int OP> void single() { printf("%-16s alone %.2f   +FFMA %.2f   +LOP3 %.2f   (cycles per slot per SMSP; PTX ops that expand to 2 SASS count as one slot)\n", kNames[OP], run<OP, NOPS>(), run<OP, FFMA>(), run<OP, LOP3>()); }

int main()
{
    single<FFMA>(); single<IMAD>(); single<IMADI>(); single<LOP3>(); single<PRMT>(); single<IADD3>(); single<SHF>(); single<VMNMX>(); single<VMNMX3>();
    single<HMNMX2>(); single<HSET2>(); single<HADD2>(); single<HFMA2>(); single<VABSD4>(); single<VIADD16>(); single<FMNMX>(); single<IMNMX>();
    single<POPC>(); single<IDP4A>(); single<LEA>(); single<ISETPSEL>(); single<VOTE>(); single<FLO>();
    printf("VIMNMX+HMNMX2 %.2f  VIMNMX+IMAD %.2f  HMNMX2+HSET2 %.2f  HMNMX2+IMAD %.2f  PRMT+IMAD %.2f  HSET2+HADD2 %.2f  VABSD4+VIMNMX %.2f\n",
           run<VMNMX, HMNMX2>(), run<VMNMX, IMAD>(), run<HMNMX2, HSET2>(), run<HMNMX2, IMAD>(), run<PRMT, IMAD>(), run<HSET2, HADD2>(), run<VABSD4, VMNMX>());
    return 0;
}
