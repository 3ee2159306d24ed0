// tma_probe.cu -- checks the TMA behaviour the FAST kernel relies on (B200, sm_100a): a 3-D u8 tensor map (x, y, frame),
// box 144 x 24 x 1 fetched by cp.async.bulk.tensor at arbitrary rows / frames and 16-byte aligned columns (an x coordinate that is
// not a multiple of 16 bytes raises 'illegal instruction' -- measured), negative / out-of-range coordinates included, into shared
// memory; completion through an mbarrier, out-of-bounds bytes zero-filled.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe.bin tools/tma_probe.cu && tools/tma_probe.bin
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int BW = 144, BH = 24;

__global__ void k(const __grid_constant__ CUtensorMap tm, int x0, int y0, int z0, uint8_t* out)
{
    __shared__ __align__(128) uint8_t tile[BW * BH];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bar), tile_s = (unsigned)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(BW * BH) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(tile_s), "l"(&tm), "r"(x0), "r"(y0), "r"(z0), "r"(bar_s) : "memory");
    }
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_s), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}

int main()
{
    const int W = 640, H = 480, P = 640, B = 3;
    const size_t frame = (size_t)P * H + 256;
    std::vector<uint8_t> h(frame * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, BW * BH);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B}, strides[2] = {(cuuint64_t)P, (cuuint64_t)frame};
    const cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    CUresult r = ((enc_t)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int cases[][3] = {{0, 0, 0}, {16, 8, 1}, {32, 27, 2}, {-16, -4, 1}, {592, 470, 2}, {112, 77, 0}, {496, 3, 1}};
    int bad = 0;
    std::vector<uint8_t> t(BW * BH);
    for (auto& cs : cases) {
        k<<<1, 128>>>(tm, cs[0], cs[1], cs[2], o);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(t.data(), o, t.size(), cudaMemcpyDeviceToHost);
        int mism = 0;
        for (int y = 0; y < BH; ++y)
            for (int x = 0; x < BW; ++x) {
                const int gx = cs[0] + x, gy = cs[1] + y;
                const uint8_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[frame * cs[2] + (size_t)gy * P + gx] : 0;
                mism += t[y * BW + x] != want;
            }
        printf("box at (%d, %d, %d): %d mismatches\n", cs[0], cs[1], cs[2], mism);
        bad += mism;
    }
    printf(bad ? "TMA PROBE FAILED\n" : "TMA PROBE OK\n");
    return bad != 0;
}
