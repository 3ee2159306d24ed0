// orbx_pyr.cuh -- A.2 INTER_LINEAR_EXACT pyramid level l from level l-1 (cv::resize inside cv::ORB::detectAndCompute,
// src/frontend.cpp:153), sm_100a, warp-private source tiles fed by TMA.
//
//   dst(x, y) = (h0 * (256 - cy) + h1 * cy + 32768) >> 16,   h = p[i0] * (256 - cx) + p[i0 + 1] * cx     (8.8 taps, host tables)
//
// Unit of work: one WARP owns a column tile of 128 output columns (lane = 4 adjacent columns) and walks a few strips of PT_RH
// output rows down it.  The source rows a strip needs arrive as ONE cp.async.bulk.tensor box (TMA; the box dimensions are fitted
// per level on the host and baked into the level's tensor map), double-buffered: the box of the next strip lands while the
// current one is computed.  No block barrier, no per-thread copy instructions.
//   horizontal pass  the lane's 4 outputs read at most 7 consecutive source bytes: three aligned 32-bit shared loads and two
//                    funnel shifts put them at the bottom of a 64-bit window; per output ONE PRMT gathers (p0, p0, p1, p1) and ONE
//                    IDP.4A multiplies by the tap weights split into bytes (c0 = a + b, c1 = c + d, each <= 128): exact.
//   vertical pass    the horizontal pass of a source row is taken once and reused by the next output row (ping-pong registers);
//                    two IMADs per pixel (h <= 255 * 256, so byte 2 of the 24-bit sum IS the rounded pixel) and three PRMTs per quad.
// ALU pipe (SHF, PRMT) and FMA pipe (IDP, IMAD) carry about the same number of instructions.
// Levels whose scale factor makes a lane's 4 outputs span more than 7 source bytes, or whose boxes would not fit, keep the
// register-window kernel k_pyr_down (orbx_kernels.cuh).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "orbx_geom.h"

namespace orbx {

constexpr int PT_RH = 16;                // output rows per strip
constexpr int PT_CW = 128;               // output columns per column tile
constexpr int PT_NWARP = 4;              // warps per CTA (independent units)

struct PyrMaps { CUtensorMap m[ORBX_LEVELS_MAX]; };   // m[l]: SOURCE level l-1 as (x, y, frame), box pt_bw x pt_bh x 1 of destination level l

__host__ __device__ inline int pt_buf_bytes(int bw, int bh) { return (bw * bh + 16 + 127) / 128 * 128; }   // + 16: a lane's third word may lie past the last row
__host__ __device__ inline int pt_warp_bytes(int bw, int bh) { return 2 * pt_buf_bytes(bw, bh) + 128; }   // two boxes + two mbarriers

// (min 6 CTAs per SM stated: ptxas then spends 70 instead of 56 registers per thread and the seven levels take 0.163 instead of 0.175 ms)
__global__ void __launch_bounds__(PT_NWARP * 32, 6) k_pyr_tma(const __grid_constant__ Geom g, const __grid_constant__ CUtensorMap map, int l, int f0,
                                                           uint8_t* __restrict__ pyr, const uint32_t* __restrict__ tabs, int* __restrict__ status)
{
    extern __shared__ __align__(128) uint8_t pt_smem[];
    const LevelGeom& D = g.L[l];
    const LevelGeom& S = g.L[l - 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int task = blockIdx.x * PT_NWARP + wid;
    if (task >= D.pt_ntask) return;                          // warps are independent: no block-level synchronisation anywhere
    const int cx = task % D.pt_ncx, sg = task / D.pt_ncx;
    const int nstrips = (D.h + PT_RH - 1) / PT_RH;
    const int strip0 = sg * D.pt_k, strip1 = min(strip0 + D.pt_k, nstrips);
    const int f = blockIdx.y;
    const int bw = D.pt_bw, bh = D.pt_bh;
    const unsigned bufb = (unsigned)pt_buf_bytes(bw, bh);
    const unsigned base_s = (((unsigned)__cvta_generic_to_shared(pt_smem) + 127u) & ~127u) + (unsigned)wid * (unsigned)pt_warp_bytes(bw, bh);
    const unsigned bar_s = base_s + 2u * bufb;
    const uint32_t* xtab = tabs + D.xtab;
    const uint32_t* ytab = tabs + D.ytab;
    const int x = cx * PT_CW + 4 * lane;
    const int X0 = (int)(__ldg(xtab + cx * PT_CW) & 0xffffu) & ~15;      // first source column of the warp's boxes (TMA: 16-byte granularity)

    // ---- per-lane constants of the horizontal pass
    unsigned sel[4], wgt[4], a_off, sh;
    {
        const bool live = x < D.w;
        const int b0 = live ? (int)(__ldg(xtab + x) & 0xffffu) : X0;     // first source byte the lane needs
        const int rel = b0 - X0;
        a_off = (unsigned)(rel & ~3);
        sh = (unsigned)(rel & 3) * 8u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned s = 0x1100u, w = 0u;                                // columns past the level's width: weights 0 -> pixel 0 (row padding stays zero)
            if (x + k < D.w) {
                const uint32_t t = __ldg(xtab + x + k);
                const unsigned off = (t & 0xffffu) - (unsigned)b0;       // 0 .. 6
                const unsigned c1 = t >> 16, c0 = 256u - c1;
                const unsigned c0a = min(c0, 128u), c1a = min(c1, 128u);
                s = off | (off << 4) | ((off + 1u) << 8) | ((off + 1u) << 12);
                w = c0a | ((c0 - c0a) << 8) | (c1a << 16) | ((c1 - c1a) << 24);
            }
            sel[k] = s; wgt[k] = w;
        }
    }
    if (lane == 0) {
        mbar_init(bar_s, 1);
        mbar_init(bar_s + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto request = [&](int strip, unsigned b) {
        if (lane == 0) {
            const int s_first = (int)(__ldg(ytab + strip * PT_RH) & 0xffffu);
            mbar_expect_tx(bar_s + 8u * b, (unsigned)(bw * bh));
            tma_load_tile_3d(base_s + b * bufb, &map, X0, s_first, f0 + f, bar_s + 8u * b);
        }
    };
    // horizontal pass of the box row whose (lane-adjusted) shared address is ad
    auto hpass = [&](unsigned ad, unsigned* h) {
        const unsigned w0 = lds_u32(ad), w1 = lds_u32(ad + 4u), w2 = lds_u32(ad + 8u);
        const unsigned lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __dp4a(__byte_perm(lo, hi, sel[k]), wgt[k], 0u);
    };

    request(strip0, 0u);
    unsigned par0 = 0u, par1 = 0u;
    const int pitch = D.pitch, sh1 = S.h - 1;
    const bool store = x < pitch;
    uint8_t* const dst0 = pyr + (size_t)(f0 + f) * g.pyr_frame + D.img_off + x;   // (pyr is the whole buffer: TMA coordinates are absolute frames)
#pragma unroll 1
    for (int strip = strip0; strip < strip1; ++strip) {
        const unsigned b = (unsigned)(strip - strip0) & 1u;
        if (strip + 1 < strip1) request(strip + 1, b ^ 1u);              // the other box was consumed before the __syncwarp() that ended the last strip
        const unsigned buf_s = base_s + b * bufb;
        if (!mbar_wait(bar_s + 8u * b, b ? par1 : par0)) { if (lane == 0) atomicOr(&status[f], 2); return; }
        if (b) par1 ^= 1u; else par0 ^= 1u;
        const int ys = strip * PT_RH, nr = min(PT_RH, D.h - ys);
        // vertical taps of the strip's rows: lane r holds row ys + r, handed out by one shuffle per row
        const uint32_t ty_all = lane < nr ? __ldg(ytab + ys + lane) : 0u;
        const int s_first = (int)(__shfl_sync(0xffffffffu, ty_all, 0) & 0xffffu);
        const unsigned row0_s = buf_s + a_off - (unsigned)(s_first * bw);  // box row of source row s at row0_s + s * bw
        int have = -1;                                                   // source row whose horizontal pass sits in the "top" registers
        unsigned hA[4], hB[4];
        uint8_t* dst = dst0 + (size_t)ys * pitch;
        // one output row: T = horizontal pass of its upper source row (already there unless the mapping skipped a row), N = lower
        auto row = [&](int r, unsigned* T, unsigned* N) {
            const uint32_t ty = __shfl_sync(0xffffffffu, ty_all, r);
            const int s0 = (int)(ty & 0xffffu), s1 = min(s0 + 1, sh1);
            const unsigned cy1 = ty >> 16, cy0 = 256u - cy1;
            if (s0 != have) hpass(row0_s + (unsigned)(s0 * bw), T);      // warp-uniform
            hpass(row0_s + (unsigned)(s1 * bw), N);                      // (the image's bottom row clamps: s1 == s0, taken twice -- no branch)
            unsigned v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = T[k] * cy0 + (N[k] * cy1 + 32768u);
            if (store) *reinterpret_cast<uint32_t*>(dst) = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
            dst += pitch;
            have = s1;
        };
        int r = 0;
#pragma unroll 1
        for (; r + 1 < nr; r += 2) { row(r, hA, hB); row(r + 1, hB, hA); }   // ping-pong without register copies
        if (r < nr) row(r, hA, hB);
        __syncwarp();                                                    // every lane is done with this box
    }
}

}  // namespace orbx
