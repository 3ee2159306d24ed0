// orbx_kernels.cuh -- sm_100a kernels of the ORB extraction path (one stage of cv::ORB::detectAndCompute each;
// the reference invokes that operator at src/frontend.cpp:153).  Stage arithmetic follows SURVEY.md Appendix A
// and is bit-exact against OpenCV: integer stages in integer arithmetic, float stages with explicit
// round-to-nearest intrinsics (never contracted) and __fmaf_rn only where OpenCV's own build uses FMA (A.8).
//
//   k_gray          A.1   BGR -> gray (level 0)                         HBM-bound, 4 px / thread
//   k_pyr_down      A.2   INTER_LINEAR_EXACT level l from level l-1     HBM/L2-bound, 4 px / thread
//   k_fast_bands    A.3   FAST-9/16 score + 3x3 NMS -> per-row lists    smem tiles with halos, u16x2 SIMD min/max,
//                                                                       ballot compaction, raster order kept
//   k_select        A.4-6 retainBest(2n) -> Harris -> retainBest(n)     libstdc++ introselect order reproduced
//   k_describe      A.7-10 IC angle, 7x7 blur of the sampled patch, steered rBRIEF-256, cv::KeyPoint records
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "orbx_geom.h"

namespace orbx {

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ unsigned vmin2(unsigned a, unsigned b) { return __vminu2(a, b); }
__device__ __forceinline__ unsigned vmax2(unsigned a, unsigned b) { return __vmaxu2(a, b); }
__device__ __forceinline__ unsigned vmax3(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned vmin3(unsigned a, unsigned b, unsigned c) { return __vimin3_u16x2(a, b, c); }

// ------------------------------------------------------------------------------------------------ A.1 gray
// g = (3735 B + 19235 G + 9798 R + 16384) >> 15.  One thread = 4 consecutive pixels of one row.
template <int CH>
__global__ void __launch_bounds__(256) k_gray(const uint8_t* __restrict__ in, unsigned long long frame_stride,
                                              unsigned long long step, int aligned4, const __grid_constant__ Geom g,
                                              uint8_t* __restrict__ pyr)
{
    const LevelGeom& L = g.L[0];
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (y >= L.h || x >= L.pitch) return;
    const uint8_t* src = in + (size_t)f * frame_stride + (size_t)y * step + (size_t)x * CH;
    uint32_t out = 0;
    if (CH == 3) {
        if (aligned4 && x + 3 < L.w) {
            const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
            const uint32_t w0 = __ldg(s4), w1 = __ldg(s4 + 1), w2 = __ldg(s4 + 2);
            // bytes: b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
            const uint32_t p0 = (3735u * (w0 & 255u) + 19235u * ((w0 >> 8) & 255u) + 9798u * ((w0 >> 16) & 255u) + 16384u) >> 15;
            const uint32_t p1 = (3735u * (w0 >> 24) + 19235u * (w1 & 255u) + 9798u * ((w1 >> 8) & 255u) + 16384u) >> 15;
            const uint32_t p2 = (3735u * ((w1 >> 16) & 255u) + 19235u * (w1 >> 24) + 9798u * (w2 & 255u) + 16384u) >> 15;
            const uint32_t p3 = (3735u * ((w2 >> 8) & 255u) + 19235u * ((w2 >> 16) & 255u) + 9798u * (w2 >> 24) + 16384u) >> 15;
            out = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (x + k < L.w) {
                    const uint32_t p = (3735u * __ldg(src + 3 * k) + 19235u * __ldg(src + 3 * k + 1) + 9798u * __ldg(src + 3 * k + 2) + 16384u) >> 15;
                    out |= p << (8 * k);
                }
        }
    } else {
        if (aligned4 && x + 3 < L.w) out = __ldg(reinterpret_cast<const uint32_t*>(src));
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (x + k < L.w) out |= (uint32_t)__ldg(src + k) << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t*>(pyr + (size_t)f * g.pyr_frame + L.img_off + (size_t)y * L.pitch + x) = out;
}

// ------------------------------------------------------------------------------------------------ A.2 pyramid
// dst(x,y) = (h0*(256-cy) + h1*cy + 32768) >> 16,  h = p[i0]*(256-cx) + p[i1]*cx  (8.8 taps from host tables).
__global__ void __launch_bounds__(256) k_pyr_down(const __grid_constant__ Geom g, int l, uint8_t* __restrict__ pyr,
                                                  const uint32_t* __restrict__ tabs)
{
    const LevelGeom& D = g.L[l];
    const LevelGeom& S = g.L[l - 1];
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (y >= D.h || x >= D.pitch) return;
    const uint8_t* src = pyr + (size_t)f * g.pyr_frame + S.img_off;
    const uint32_t ty = __ldg(tabs + D.ytab + y);
    const int y0 = ty & 0xffff;
    const uint32_t cy1 = ty >> 16, cy0 = 256u - cy1;
    const int y1 = min(y0 + 1, S.h - 1);
    const uint8_t* r0 = src + (size_t)y0 * S.pitch;
    const uint8_t* r1 = src + (size_t)y1 * S.pitch;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (x + k < D.w) {
            const uint32_t tx = __ldg(tabs + D.xtab + x + k);
            const int i0 = tx & 0xffff;
            const uint32_t cx1 = tx >> 16, cx0 = 256u - cx1;
            const int i1 = min(i0 + 1, S.w - 1);
            const uint32_t h0 = r0[i0] * cx0 + r0[i1] * cx1;
            const uint32_t h1 = r1[i0] * cx0 + r1[i1] * cx1;
            uint32_t v = (h0 * cy0 + h1 * cy1 + 32768u) >> 16;
            v = min(v, 255u);
            out |= v << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t*>(pyr + (size_t)f * g.pyr_frame + D.img_off + (size_t)y * D.pitch + x) = out;
}

// ------------------------------------------------------------------------------------------------ A.3 FAST + NMS
// One CTA = one band of R inner rows of one level of one frame, walked left to right in chunks of CW columns, so
// every row's survivors come out in x order and the per-row lists concatenate to OpenCV's raster order.
// Only the region that can survive the 31-px border filter is evaluated.
//
// Per chunk:  load (R+8) x (CW+16) pixels widened to u16 into smem
//   phase A   every pixel pair: 4-compass-point rejection test in u16x2 SIMD (native VIMNMX.U16x2) -> pass queue
//   phase B   queued pixels: full 16-point test.  Each circle pixel is packed (p | (255-p) << 16) so ONE sliding
//             max over the 16 nine-long arcs (VIMNMX3.U16x2) yields both min-of-max(p) and max-of-min(p):
//             A = v - min_arcs max p,  -B = max_arcs min p - v,  score = max(A, -B) - 1  (corner iff > t)
//   phase C   3x3 NMS of the corners on the smem score tile -> per-row bit masks
//   phase D   ordered extraction of the bit masks (popc prefix) -> global per-row lists
template <int R, int CW, int NT>
__global__ void __launch_bounds__(NT) k_fast_bands(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr,
                                                   uint32_t* __restrict__ rowcnt, uint32_t* __restrict__ rowent)
{
    constexpr int TP = CW + 16;            // image tile pitch (pixels, u16 each)
    constexpr int TPW = TP / 2;            // ... in 32-bit words
    constexpr int TR = R + 8;              // image tile rows
    constexpr int SP = CW + 4;             // score tile pitch (bytes)
    constexpr int SR = R + 2;              // score tile rows
    constexpr int MW = CW / 32;            // mask words per row
    constexpr int T = ORBX_FAST_T;
    static_assert(R * MW <= NT, "phase D needs one thread per mask word");
    static_assert(CW % 32 == 0 && TP % 4 == 0, "tile shape");

    __shared__ __align__(16) uint16_t s_img[TR * TP];
    __shared__ __align__(4) uint8_t s_score[SR * SP];
    __shared__ uint16_t s_q[SR * SP];
    __shared__ uint32_t s_mask[R * MW];
    __shared__ uint32_t s_rowcnt[R];
    __shared__ int s_qn;

    const int tid = threadIdx.x, lane = tid & 31;
    const int f = blockIdx.y;
    // which level / band
    int l = 0;
#pragma unroll 1
    for (int i = 1; i < g.nlevels; ++i) if ((int)blockIdx.x >= g.L[i].band0) l = i;
    const LevelGeom& L = g.L[l];
    const int band = blockIdx.x - L.band0;
    if (band >= L.nbands) return;
    const int y0 = ORBX_EDGE + band * R;
    const int y1 = min(y0 + R, L.h - ORBX_EDGE);           // output rows [y0, y1)
    const int xend = L.w - ORBX_EDGE;                       // output cols [31, xend)
    const uint8_t* img = pyr + (size_t)f * g.pyr_frame + L.img_off;
    uint32_t* cnt_out = rowcnt + (size_t)f * g.cnt_frame + L.cnt_off;
    uint32_t* ent_out = rowent + (size_t)f * g.ent_frame + L.ent_off;

    if (tid < R) s_rowcnt[tid] = 0;

    const uint32_t* W = reinterpret_cast<const uint32_t*>(s_img);
    for (int ox0 = 28; ox0 < xend; ox0 += CW) {
        const int ox1 = min(ox0 + CW, xend);
        const int ix0 = ox0 - 8;
        __syncthreads();                                     // previous chunk fully consumed
        // ---- load tile rows [y0-4, y1+4), cols [ix0, ix0+TP) as u16
        {
            const int rows = y1 - y0 + 8;
            constexpr int QPR = TP / 4;                      // 4-pixel groups per tile row
            for (int i = tid; i < rows * QPR; i += NT) {
                const int ty = i / QPR, q = i - ty * QPR;
                const int gx = ix0 + q * 4;
                uint32_t w = 0;
                if (gx < L.pitch) w = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)(y0 - 4 + ty) * L.pitch + gx));
                uint2 o;
                o.x = __byte_perm(w, 0, 0x4140);
                o.y = __byte_perm(w, 0, 0x4342);
                *reinterpret_cast<uint2*>(&s_img[ty * TP + q * 4]) = o;
            }
            for (int i = tid; i < SR * SP / 4; i += NT) reinterpret_cast<uint32_t*>(s_score)[i] = 0;
            if (tid < R * MW) s_mask[tid] = 0;
            if (tid == 0) s_qn = 0;
        }
        __syncthreads();
        // ---- phase A: compass rejection on pixel pairs
        {
            constexpr int NP = SP / 2;
            constexpr unsigned K = ((511u - T) << 16) | (511u - T);
            for (int i0 = 0; i0 < SR * NP; i0 += NT) {
                const int i = i0 + tid;
                const int sy = i / NP, sx = (i - sy * NP) * 2;
                const int x = ox0 - 2 + sx, y = y0 - 1 + sy;
                unsigned m = 0;
                if (i < SR * NP && x <= ox1 && y <= y1) {
                    const int b = (sy + 3) * TPW + (sx + 6) / 2;
                    const unsigned c = W[b], n = W[b - 3 * TPW], s = W[b + 3 * TPW];
                    const unsigned e = __byte_perm(W[b + 1], W[b + 2], 0x5432);
                    const unsigned w = __byte_perm(W[b - 2], W[b - 1], 0x5432);
                    const unsigned D = vmax2(vmin2(n, s), vmin2(e, w));
                    const unsigned B = vmin2(vmax2(n, s), vmax2(e, w));
                    m = ((c + K - D) | (B + K - c)) & 0x02000200u;
                }
                const unsigned blo = __ballot_sync(0xffffffffu, m & 0x200u);
                const unsigned bhi = __ballot_sync(0xffffffffu, m & 0x02000000u);
                const int nlo = __popc(blo), tot = nlo + __popc(bhi);
                int base = 0;
                if (lane == 0 && tot) base = atomicAdd(&s_qn, tot);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (m & 0x200u) s_q[base + __popc(blo & lanemask_lt())] = (uint16_t)((sy << 9) | sx);
                if (m & 0x02000000u) s_q[base + nlo + __popc(bhi & lanemask_lt())] = (uint16_t)((sy << 9) | (sx + 1));
            }
        }
        __syncthreads();
        const int qn = s_qn;
        // ---- phase B: full segment test + score on the queued pixels
        for (int i = tid; i < qn; i += NT) {
            const int e = s_q[i];
            const int sy = e >> 9, sx = e & 511;
            const uint16_t* c = &s_img[(sy + 3) * TP + sx + 6];
            const int v = c[0];
            unsigned q[16];
#define ORBX_PK(k, dx, dy) q[k] = (unsigned)c[(dy) * TP + (dx)] * 0xFFFF0001u + 0x00FF0000u
            ORBX_PK(0, 0, 3);   ORBX_PK(1, 1, 3);   ORBX_PK(2, 2, 2);    ORBX_PK(3, 3, 1);
            ORBX_PK(4, 3, 0);   ORBX_PK(5, 3, -1);  ORBX_PK(6, 2, -2);   ORBX_PK(7, 1, -3);
            ORBX_PK(8, 0, -3);  ORBX_PK(9, -1, -3); ORBX_PK(10, -2, -2); ORBX_PK(11, -3, -1);
            ORBX_PK(12, -3, 0); ORBX_PK(13, -3, 1); ORBX_PK(14, -2, 2);  ORBX_PK(15, -1, 3);
#undef ORBX_PK
            unsigned m3[16], m9[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) m3[k] = vmax3(q[k], q[(k + 1) & 15], q[(k + 2) & 15]);
#pragma unroll
            for (int k = 0; k < 16; ++k) m9[k] = vmax3(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
            unsigned mm = vmin3(vmin3(vmin3(m9[0], m9[1], m9[2]), vmin3(m9[3], m9[4], m9[5]), vmin3(m9[6], m9[7], m9[8])),
                                vmin3(vmin3(m9[9], m9[10], m9[11]), vmin3(m9[12], m9[13], m9[14]), m9[15]),
                                0xFFFFFFFFu);
            const int A = v - (int)(mm & 0xffffu);
            const int nB = 255 - (int)(mm >> 16) - v;
            const int sc = max(A, nB);
            if (sc > T) s_score[sy * SP + sx] = (uint8_t)(sc - 1);
        }
        __syncthreads();
        // ---- phase C: 3x3 NMS (strict >) of corners inside the output region
        for (int i = tid; i < qn; i += NT) {
            const int e = s_q[i];
            const int sy = e >> 9, sx = e & 511;
            const uint8_t* p = &s_score[sy * SP + sx];
            const int s = p[0];
            if (!s) continue;
            const int x = ox0 - 2 + sx, y = y0 - 1 + sy;
            if (x < max(ox0, ORBX_EDGE) || x >= ox1 || y < y0 || y >= y1) continue;
            if (s > p[-1] && s > p[1] && s > p[-SP - 1] && s > p[-SP] && s > p[-SP + 1] && s > p[SP - 1] && s > p[SP] && s > p[SP + 1]) {
                const int bit = x - ox0;
                atomicOr(&s_mask[(y - y0) * MW + (bit >> 5)], 1u << (bit & 31));
            }
        }
        __syncthreads();
        // ---- phase D: ordered extraction; thread t owns mask word t (MW consecutive lanes = one row)
        if (tid < R * MW) {
            const int row = tid / MW, wi = tid - row * MW;
            uint32_t m = s_mask[tid];
            const int cnt = __popc(m);
            int pre = cnt;                                   // inclusive prefix inside the row's MW-lane group
#pragma unroll
            for (int d = 1; d < MW; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, pre, d, MW);
                if (wi >= d) pre += o;
            }
            const uint32_t base = s_rowcnt[row];
            __syncwarp();
            uint32_t slot = base + pre - cnt;
            uint32_t* dst = ent_out + (size_t)(y0 - ORBX_EDGE + row) * L.ent_pitch;
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int x = ox0 + wi * 32 + b;
                const uint32_t sc = s_score[(row + 1) * SP + (x - ox0 + 2)];
                dst[slot++] = (uint32_t)x | (sc << 16);
            }
            if (wi == MW - 1) s_rowcnt[row] = base + pre;
        }
    }
    __syncthreads();
    if (tid < y1 - y0) cnt_out[y0 - ORBX_EDGE + tid] = s_rowcnt[tid];
}

// ------------------------------------------------------------------------------------------------ A.6 retainBest
// Warp-cooperative, exact emulation of KeyPointsFilter::retainBest = libstdc++ std::nth_element (__introselect:
// median-of-3 to first, Hoare __unguarded_partition, insertion sort of <= 3) + bidirectional std::partition.
// The pointer scans are vectorised 32 wide with ballots; every swap is the sequential algorithm's swap, so the
// resulting ORDER is libstdc++'s.  `v` may live in shared or global memory.
__device__ __forceinline__ void elem_swap(Elem* v, int a, int b) { const Elem t = v[a]; v[a] = v[b]; v[b] = t; }

// first i in [start, end) with stopper(v[i]); `end` if none.  MODE 0: !(r > piv)   1: !(r >= piv)
template <int MODE>
__device__ __forceinline__ int scan_up(const Elem* v, int start, int end, float piv, int lane)
{
    for (int base = start; base < end; base += 32) {
        const int i = base + lane;
        bool p = false;
        if (i < end) { const float r = v[i].response; p = MODE == 0 ? !(r > piv) : !(r >= piv); }
        const unsigned b = __ballot_sync(0xffffffffu, p);
        if (b) return base + __ffs(b) - 1;
    }
    return end;
}
// last i in (stop, start] with stopper(v[i]); `stop` if none.  MODE 0: !(piv > r)   1: (r >= piv)
template <int MODE>
__device__ __forceinline__ int scan_down(const Elem* v, int start, int stop, float piv, int lane)
{
    for (int base = start; base > stop; base -= 32) {
        const int i = base - lane;
        bool p = false;
        if (i > stop) { const float r = v[i].response; p = MODE == 0 ? !(piv > r) : (r >= piv); }
        const unsigned b = __ballot_sync(0xffffffffu, p);
        if (b) return base - (__ffs(b) - 1);
    }
    return stop;
}

// returns the new length; sets *flag |= 1 if the depth limit (heap-select fallback) would have been hit
__device__ int retain_best_warp(Elem* v, int len, int m, int lane, int* flag)
{
    if (m < 0 || len <= m) return len;
    if (m == 0) return 0;
    // ---- std::nth_element(v, v + m - 1, v + len, response-greater)
    {
        const int nth = m - 1;
        int first = 0, last = len;
        int depth = 2 * (31 - __clz(len));
        while (last - first > 3) {
            if (depth == 0) { *flag |= 1; return m; }   // heap-select fallback: flagged, order not reproduced
            --depth;
            const int mid = first + (last - first) / 2;
            const int a = first + 1, b = mid, c = last - 1;
            const float ra = v[a].response, rb = v[b].response, rc = v[c].response;
            int pick;
            if (ra > rb) pick = (rb > rc) ? b : ((ra > rc) ? c : a);
            else         pick = (ra > rc) ? a : ((rb > rc) ? c : b);
            __syncwarp();
            if (lane == 0) elem_swap(v, first, pick);
            __syncwarp();
            const float piv = v[first].response;
            int f = first + 1, l = last;
            for (;;) {
                f = scan_up<0>(v, f, last, piv, lane);
                --l;
                l = scan_down<0>(v, l, first, piv, lane);
                if (!(f < l)) break;
                if (lane == 0) elem_swap(v, f, l);
                __syncwarp();
                ++f;
            }
            if (f <= nth) first = f; else last = f;
        }
        __syncwarp();
        if (lane == 0) {                                     // __insertion_sort, descending, <= 3 elements
            for (int i = first + 1; i < last; ++i) {
                const Elem val = v[i];
                int j = i;
                while (j > first && val.response > v[j - 1].response) { v[j] = v[j - 1]; --j; }
                v[j] = val;
            }
        }
        __syncwarp();
    }
    // ---- std::partition(v + m, v + len, response >= thr), bidirectional version
    const float thr = v[m - 1].response;
    int f = m, l = len;
    for (;;) {
        f = scan_up<1>(v, f, l, thr, lane);                  // first !pred in [f, l)
        if (f == l) return f;
        --l;
        l = scan_down<1>(v, l, f, thr, lane);                // last pred in (f, l]; f if none
        if (f == l) return f;
        if (lane == 0) elem_swap(v, f, l);
        __syncwarp();
        ++f;
    }
}

// ------------------------------------------------------------------------------------------------ A.5 Harris
__device__ __forceinline__ float harris_response(const uint8_t* __restrict__ img, int pitch, int x, int y)
{
    int a = 0, b = 0, c = 0;
    // 9x9 neighbourhood walked row by row with a 3-row register window
    uint8_t r0[9], r1[9], r2[9];
    const uint8_t* p = img + (size_t)(y - 4) * pitch + (x - 4);
#pragma unroll
    for (int i = 0; i < 9; ++i) { r0[i] = __ldg(p + i); r1[i] = __ldg(p + pitch + i); }
#pragma unroll
    for (int row = 0; row < 7; ++row) {
        const uint8_t* pr = p + (size_t)(row + 2) * pitch;
#pragma unroll
        for (int i = 0; i < 9; ++i) r2[i] = __ldg(pr + i);
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            const int Ix = ((int)r1[i + 1] - (int)r1[i - 1]) * 2 + ((int)r0[i + 1] - (int)r0[i - 1]) + ((int)r2[i + 1] - (int)r2[i - 1]);
            const int Iy = ((int)r2[i] - (int)r0[i]) * 2 + ((int)r2[i - 1] - (int)r0[i - 1]) + ((int)r2[i + 1] - (int)r0[i + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) { r0[i] = r1[i]; r1[i] = r2[i]; }
    }
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float det = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
    const float tr = __fadd_rn(fa, fb);
    const float ktr2 = __fmul_rn(__fmul_rn(__int_as_float(0x3d23d70a), tr), tr);       // 0.04f
    return __fmul_rn(__fsub_rn(det, ktr2), __int_as_float(0x25ddced1));                // (1/7140)^4, rounded stepwise
}

// ------------------------------------------------------------------------------------------------ selection kernel
// One CTA per (level, frame): gather the per-row FAST lists (raster order) -> retainBest(2 n_l) on the FAST score
// -> Harris on the survivors -> retainBest(n_l) on Harris.  The final list is left as the prefix of the level's
// global workspace; its length goes to fincnt.  The working array lives in shared memory when it fits.
constexpr int SEL_NT = 128;
constexpr int SEL_SMEM_ELEMS = 4096;

__global__ void __launch_bounds__(SEL_NT) k_select(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr,
                                                   const uint32_t* __restrict__ rowcnt, const uint32_t* __restrict__ rowent,
                                                   Elem* __restrict__ work, int* __restrict__ fincnt, int* __restrict__ status)
{
    __shared__ Elem s_v[SEL_SMEM_ELEMS];
    __shared__ int s_warp[SEL_NT / 32];
    __shared__ int s_n;
    const int l = blockIdx.x, f = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const LevelGeom& L = g.L[l];
    Elem* gv = work + (size_t)f * g.ws_frame + L.ws_off;
    if (L.in_w <= 0 || L.in_h <= 0) { if (tid == 0) fincnt[f * g.nlevels + l] = 0; return; }
    const uint32_t* cnt = rowcnt + (size_t)f * g.cnt_frame + L.cnt_off;
    const uint32_t* ent = rowent + (size_t)f * g.ent_frame + L.ent_off;
    // ---- gather rows in raster order
    const int nr = L.in_h;
    const int rpt = (nr + SEL_NT - 1) / SEL_NT;
    const int rb = min(tid * rpt, nr), re = min(rb + rpt, nr);
    int mine = 0;
    for (int r = rb; r < re; ++r) mine += (int)cnt[r];
    int inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < SEL_NT / 32; ++i) { const int s = s_warp[i]; if (i < wid) woff += s; total += s; }
    const int N = total;
    Elem* v = (N <= SEL_SMEM_ELEMS) ? s_v : gv;
    {
        int o = woff + inc - mine;
        for (int r = rb; r < re; ++r) {
            const int c = (int)cnt[r];
            const uint32_t* e = ent + (size_t)r * L.ent_pitch;
            for (int i = 0; i < c; ++i) {
                const uint32_t w = e[i];
                Elem el;
                el.response = (float)(w >> 16);
                el.pos = ((uint32_t)(r + ORBX_EDGE) << 16) | (w & 0xffffu);
                v[o++] = el;
            }
        }
    }
    __syncthreads();
    int flag = 0;
    // ---- retainBest(2 n_l) on the FAST score (warp 0)
    if (wid == 0) {
        const int n1 = retain_best_warp(v, N, 2 * L.quota, lane, &flag);
        if (lane == 0) s_n = n1;
    }
    __syncthreads();
    const int n1 = s_n;
    // ---- Harris on the unblurred level
    const uint8_t* img = pyr + (size_t)f * g.pyr_frame + L.img_off;
    for (int i = tid; i < n1; i += SEL_NT) {
        const uint32_t pos = v[i].pos;
        v[i].response = harris_response(img, L.pitch, (int)(pos & 0xffffu), (int)(pos >> 16));
    }
    __syncthreads();
    // ---- retainBest(n_l) on Harris (warp 0)
    if (wid == 0) {
        const int n2 = retain_best_warp(v, n1, L.quota, lane, &flag);
        if (lane == 0) { s_n = n2; fincnt[f * g.nlevels + l] = n2; if (flag) atomicOr(&status[f], flag); }
    }
    __syncthreads();
    const int n2 = s_n;
    if (v != gv) for (int i = tid; i < n2; i += SEL_NT) gv[i] = v[i];
}

// ------------------------------------------------------------------------------------------------ A.7 / A.9 scalar math
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float P1 = __int_as_float(0x4265226f), P3 = __int_as_float(0xc19556ee);
    const float P5 = __int_as_float(0x410e9fbf), P7 = __int_as_float(0xc0228ad9);
    const float eps = __int_as_float(0x25800000);            // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) { c = __fdiv_rn(ay, __fadd_rn(ax, eps)); }
    else          { c = __fdiv_rn(ax, __fadd_rn(ay, eps)); }
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(P7, c2); a = __fadd_rn(a, P5); a = __fmul_rn(a, c2); a = __fadd_rn(a, P3);
    a = __fmul_rn(a, c2);  a = __fadd_rn(a, P1); a = __fmul_rn(a, c);
    if (!(ax >= ay)) a = __fsub_rn(90.0f, a);
    if (x < 0.0f) a = __fsub_rn(180.0f, a);
    if (y < 0.0f) a = __fsub_rn(360.0f, a);
    return a;
}

// glibc 2.39 sinf/cosf restated in FP64 with separate roundings; valid for |ang| < 120.
__device__ __forceinline__ float sin_poly_d(double x, double x2)
{
    const double S0 = -0x1.555545995a603p-3, S1 = 0x1.1107605230bc4p-7, S2 = -0x1.994eb3774cf24p-13;
    const double x3 = __dmul_rn(x, x2);
    const double s1 = __dadd_rn(S1, __dmul_rn(x2, S2));
    const double x7 = __dmul_rn(x3, x2);
    const double s = __dadd_rn(x, __dmul_rn(x3, S0));
    return (float)__dadd_rn(s, __dmul_rn(x7, s1));
}
__device__ __forceinline__ float cos_poly_d(double x2, double sg)
{
    const double C0 = 1.0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double c0 = C0 * sg, c1v = C1 * sg, c2v = C2 * sg, c3v = C3 * sg, c4v = C4 * sg;   // exact sign flips
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __dadd_rn(c3v, __dmul_rn(x2, c4v));
    const double c1 = __dadd_rn(c0, __dmul_rn(x2, c1v));
    const double x6 = __dmul_rn(x4, x2);
    const double c = __dadd_rn(c1, __dmul_rn(x4, c2v));
    return (float)__dadd_rn(c, __dmul_rn(x6, c2));
}
__device__ __forceinline__ void glibc_sincosf(float ang, float* sn, float* cs)
{
    const unsigned top = (__float_as_uint(ang) >> 20) & 0x7ffu;
    double x = (double)ang;
    if (top < ((0x3f490fdbu >> 20) & 0x7ffu)) {               // |ang| < pi/4 (top-12-bit compare, as glibc)
        if (top < ((0x39800000u >> 20) & 0x7ffu)) { *sn = ang; *cs = 1.0f; return; }   // |ang| < 2^-12
        const double x2 = __dmul_rn(x, x);
        *sn = sin_poly_d(x, x2);
        *cs = cos_poly_d(x2, 1.0);
        return;
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double)n, 0x1.921FB54442D18p0));
    const double x2 = __dmul_rn(x, x);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const double tsg = (n & 2) ? -1.0 : 1.0;
    if (n & 1) { *sn = cos_poly_d(x2, tsg); *cs = sin_poly_d(x * s, x2); }
    else       { *sn = sin_poly_d(x * s, x2); *cs = cos_poly_d(x2, tsg); }
}

// ------------------------------------------------------------------------------------------------ A.7-A.10 describe
// One CTA = up to DESC_KPB keypoints of one frame.  Per keypoint: stage the 43x43 unblurred patch as f32 in smem;
// IC moments over the radius-15 disc; fastAtan2; blur the 37x37 core with OpenCV's float sepFilter2D arithmetic
// (row pass: general FMA chain, column pass: symmetric FMA chain, rint to u8); sample the 256 steered tests.
constexpr int DESC_KPB = 4;
constexpr int DESC_NT = 160;
constexpr int PATCH = 43, PPITCH = 44, CORE = 37, CPITCH = 40;

__constant__ int c_umax[16];

__global__ void __launch_bounds__(DESC_NT) k_describe(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr,
                                                      const Elem* __restrict__ work, const int* __restrict__ fincnt,
                                                      const int8_t* __restrict__ pattern, float* __restrict__ kps_out,
                                                      uint8_t* __restrict__ desc_out, int* __restrict__ counts_out, int cap)
{
    __shared__ float s_patch[DESC_KPB][PATCH * PPITCH];
    __shared__ uint8_t s_blur[DESC_KPB][CORE * CPITCH];
    __shared__ int s_mom[DESC_KPB][2];
    __shared__ int s_lvl[DESC_KPB], s_x[DESC_KPB], s_y[DESC_KPB];
    __shared__ float s_resp[DESC_KPB], s_ang[DESC_KPB], s_cos[DESC_KPB], s_sin[DESC_KPB];

    const int tid = threadIdx.x, f = blockIdx.y;
    const int slot0 = blockIdx.x * DESC_KPB;
    int total = 0;
    for (int l = 0; l < g.nlevels; ++l) total += fincnt[f * g.nlevels + l];
    if (blockIdx.x == 0 && tid == 0) counts_out[f] = total;
    const int lim = min(total, cap);
    if (slot0 >= lim) return;
    const int nk = min(DESC_KPB, lim - slot0);

    if (tid < nk) {
        int s = slot0 + tid, l = 0;
        for (; l < g.nlevels; ++l) { const int c = fincnt[f * g.nlevels + l]; if (s < c) break; s -= c; }
        const Elem e = work[(size_t)f * g.ws_frame + g.L[l].ws_off + s];
        s_lvl[tid] = l; s_x[tid] = (int)(e.pos & 0xffffu); s_y[tid] = (int)(e.pos >> 16); s_resp[tid] = e.response;
        s_mom[tid][0] = 0; s_mom[tid][1] = 0;
    }
    __syncthreads();
    // ---- stage patches (u8 -> f32, exact)
    for (int i = tid; i < nk * PATCH * PATCH; i += DESC_NT) {
        const int j = i / (PATCH * PATCH), r = (i - j * PATCH * PATCH) / PATCH, c = i - j * PATCH * PATCH - r * PATCH;
        const LevelGeom& L = g.L[s_lvl[j]];
        const uint8_t* img = pyr + (size_t)f * g.pyr_frame + L.img_off;
        s_patch[j][r * PPITCH + c] = (float)__ldg(img + (size_t)(s_y[j] - 21 + r) * L.pitch + (s_x[j] - 21 + c));
    }
    __syncthreads();
    // ---- IC moments: thread (j, u) walks its column of the disc
    if (tid < nk * 31) {
        const int j = tid / 31, u = tid - j * 31 - 15;
        const int au = u < 0 ? -u : u;
        int m10 = 0, m01 = 0;
#pragma unroll 1
        for (int vv = -15; vv <= 15; ++vv) {
            const int av = vv < 0 ? -vv : vv;
            if (au <= c_umax[av]) {
                const int I = (int)s_patch[j][(21 + vv) * PPITCH + 21 + u];
                m10 += u * I; m01 += vv * I;
            }
        }
        atomicAdd(&s_mom[j][0], m10);
        atomicAdd(&s_mom[j][1], m01);
    }
    __syncthreads();
    if (tid < nk) {
        const float ang = fast_atan2_deg((float)s_mom[tid][1], (float)s_mom[tid][0]);
        s_ang[tid] = ang;
        float sn, cs;
        glibc_sincosf(__fmul_rn(ang, __int_as_float(0x3c8efa35)), &sn, &cs);
        s_cos[tid] = cs; s_sin[tid] = sn;
    }
    // ---- blur: thread (j, c) produces column c of the 37x37 core with a 7-deep register window of row-pass values
    if (tid < nk * CORE) {
        const int j = tid / CORE, c = tid - j * CORE;
        const float k0 = __int_as_float(0x3d8fafb1), k1 = __int_as_float(0x3e06387e), k2 = __int_as_float(0x3e434a39), k3 = __int_as_float(0x3e5d4ae0);
        const float* P = &s_patch[j][c];
        float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f, w5 = 0.f, w6;
#pragma unroll 1
        for (int r = 0; r < PATCH; ++r) {
            const float* p = P + r * PPITCH;
            float acc = __fmul_rn(k0, p[0]);
            acc = __fmaf_rn(p[1], k1, acc); acc = __fmaf_rn(p[2], k2, acc); acc = __fmaf_rn(p[3], k3, acc);
            acc = __fmaf_rn(p[4], k2, acc); acc = __fmaf_rn(p[5], k1, acc); acc = __fmaf_rn(p[6], k0, acc);
            w6 = acc;
            if (r >= 6) {
                float o = __fmul_rn(k3, w3);
                o = __fmaf_rn(__fadd_rn(w4, w2), k2, o);
                o = __fmaf_rn(__fadd_rn(w5, w1), k1, o);
                o = __fmaf_rn(__fadd_rn(w6, w0), k0, o);
                int q = __float2int_rn(o);
                q = max(0, min(255, q));
                s_blur[j][(r - 6) * CPITCH + c] = (uint8_t)q;
            }
            w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6;
        }
    }
    __syncthreads();
    // ---- steered rBRIEF: thread (j, byte)
    if (tid < nk * 32) {
        const int j = tid >> 5, bi = tid & 31;
        const float a = s_cos[j], b = s_sin[j];
        const uint8_t* B = &s_blur[j][18 * CPITCH + 18];
        unsigned byte = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const char4 pt = *reinterpret_cast<const char4*>(pattern + (bi * 8 + t) * 4);
            const float x0 = (float)pt.x, y0 = (float)pt.y, x1 = (float)pt.z, y1 = (float)pt.w;
            const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int t0 = B[iy0 * CPITCH + ix0], t1 = B[iy1 * CPITCH + ix1];
            byte |= (unsigned)(t0 < t1) << t;
        }
        const size_t slot = (size_t)f * cap + slot0 + j;
        desc_out[slot * 32 + bi] = (uint8_t)byte;
        if (bi < 7) {
            const LevelGeom& L = g.L[s_lvl[j]];
            float val;
            switch (bi) {
                case 0: val = __fmul_rn((float)s_x[j], L.scale); break;
                case 1: val = __fmul_rn((float)s_y[j], L.scale); break;
                case 2: val = __fmul_rn(31.0f, L.scale); break;
                case 3: val = s_ang[j]; break;
                case 4: val = s_resp[j]; break;
                case 5: val = __int_as_float(s_lvl[j]); break;
                default: val = __int_as_float(-1); break;
            }
            kps_out[slot * 7 + bi] = val;
        }
    }
}

}  // namespace orbx
