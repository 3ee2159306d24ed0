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
//   k_blur          A.8   7x7 float-FMA Gaussian of the samplable region of every level (register sliding window)
//   k_describe      A.7-10 IC angle, steered rBRIEF-256 sampled from the blurred level, cv::KeyPoint records (warp / keypoint)
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
// One CTA = one band of R inner rows of one level of one frame, walked left to right in chunks of CWO = 252 output
// columns, so every row's survivors come out in x order and the per-row lists concatenate to OpenCV's raster order.
// Only the region that can survive the 31-px border filter is evaluated.
//
// Per chunk (score tile = (R+2) rows x 256 columns, x = ox0-2 .. ox0+253):
//   load      (R+8) x 272 pixels widened to u16 into smem (one warp per row, aligned 32-bit loads)
//   phase A   every pixel pair: 4-compass-point rejection test in u16x2 SIMD (native VIMNMX.U16x2); the pass bits
//             leave as warp ballots (one word per 32 pairs and parity); a block scan turns them into a queue
//   phase B   queued pixels: full 16-point test.  Each circle pixel is packed (p | (255-p) << 16) so ONE sliding
//             max over the 16 nine-long arcs (VIMNMX3.U16x2) yields both min-of-max(p) and max-of-min(p):
//             A = v - min_arcs max p,  -B = max_arcs min p - v,  score = max(A, -B) - 1  (corner iff > t)
//   phase C   3x3 NMS (strict >) of the corners on the smem score tile -> per-row bit masks
//   phase D   ordered extraction of the bit masks (popc prefix) -> global per-row lists
template <int R, int NT>
__global__ void __launch_bounds__(NT) k_fast_bands(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr,
                                                   uint32_t* __restrict__ rowcnt, uint32_t* __restrict__ rowent)
{
    constexpr int CWO = 252;               // output columns per chunk
    constexpr int SP = 256;                // score tile pitch (bytes) = pixels evaluated per row
    constexpr int SR = R + 2;              // score tile rows
    constexpr int TP = SP + 16;            // image tile pitch (pixels, u16 each): x = ox0-8 .. ox0+263
    constexpr int TPW = TP / 2;            // ... in 32-bit words
    constexpr int TR = R + 8;              // image tile rows
    constexpr int MW = 8;                  // mask words per row (252 bits used)
    constexpr int NPW = SR * 8;            // pass-bit words: [row][pair-column block of 32][parity]
    constexpr int T = ORBX_FAST_T;
    constexpr int NWARP = NT / 32;
    static_assert(NT == 256, "phase A maps 128 pair columns x 2 row groups onto 256 threads");
    static_assert(R * MW <= NT && NPW <= NT, "one thread per mask / pass word");

    __shared__ __align__(16) uint16_t s_img[TR * TP];
    __shared__ __align__(16) uint8_t s_score[SR * SP];
    __shared__ uint16_t s_q[SR * SP];      // pass queue: sy << 8 | sx
    __shared__ uint16_t s_cq[R * SP];      // corner queue (NMS candidates inside the output region; <= R x 252)
    __shared__ uint32_t s_pass[NPW];
    __shared__ uint32_t s_mask[R * MW];
    __shared__ uint32_t s_rowcnt[R];
    __shared__ int s_wsum[NWARP];
    __shared__ int s_qn, s_cn;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int f = blockIdx.y;
    int l = 0;
#pragma unroll 1
    for (int i = 1; i < g.nlevels; ++i) if ((int)blockIdx.x >= g.L[i].band0) l = i;
    const LevelGeom& L = g.L[l];
    const int band = blockIdx.x - L.band0;
    if (band >= L.nbands) return;
    const int y0 = ORBX_EDGE + band * R;
    const int y1 = min(y0 + R, L.h - ORBX_EDGE);           // output rows [y0, y1)
    const int nsr = y1 - y0 + 2;                            // score rows in use
    const int xend = L.w - ORBX_EDGE;                       // output cols [31, xend)
    const uint8_t* img = pyr + (size_t)f * g.pyr_frame + L.img_off;
    uint32_t* cnt_out = rowcnt + (size_t)f * g.cnt_frame + L.cnt_off;
    uint32_t* ent_out = rowent + (size_t)f * g.ent_frame + L.ent_off;

    if (tid < R) s_rowcnt[tid] = 0;

    const uint32_t* W = reinterpret_cast<const uint32_t*>(s_img);
    for (int ox0 = 28; ox0 < xend; ox0 += CWO) {
        const int ox1 = min(ox0 + CWO, xend);
        const int ix0 = ox0 - 8;
        const int need = min(ox1 + 2 - (ox0 - 2), SP);      // pixels per score row that matter: x = ox0-2 .. ox1+1
        __syncthreads();                                     // previous chunk fully consumed
        // ---- load tile rows [y0-4, y1+4), one warp per row, 4 pixels per lane and step
        {
            const int rows = y1 - y0 + 8;
            const int nq = min((need + 16 + 3) >> 2, TP / 4);
            for (int ty = wid; ty < rows; ty += NWARP) {
                const uint8_t* src = img + (size_t)(y0 - 4 + ty) * L.pitch + ix0;
                for (int q = lane; q < nq; q += 32) {
                    uint32_t w = 0;
                    if (ix0 + q * 4 < L.pitch) w = __ldg(reinterpret_cast<const uint32_t*>(src) + q);
                    uint2 o;
                    o.x = __byte_perm(w, 0, 0x4140);
                    o.y = __byte_perm(w, 0, 0x4342);
                    *reinterpret_cast<uint2*>(&s_img[ty * TP + q * 4]) = o;
                }
            }
            for (int i = tid; i < SR * SP / 16; i += NT) reinterpret_cast<uint4*>(s_score)[i] = make_uint4(0, 0, 0, 0);
            if (tid < R * MW) s_mask[tid] = 0;
            if (tid == 0) s_cn = 0;
        }
        __syncthreads();
        // ---- phase A: compass rejection; thread = pair column (tid & 127), rows tid >> 7, +2, ...
        {
            constexpr unsigned K = ((511u - T) << 16) | (511u - T);
            const int p = tid & 127, wc = (tid >> 5) & 3;
            const bool warp_on = wc * 64 < need;            // warp-uniform: this 64-pixel block holds needed pixels
            const bool col_on = 2 * p < need;
            for (int sy = tid >> 7; sy < nsr; sy += 2) {
                unsigned m = 0;
                if (warp_on) {
                    if (col_on) {
                        const int b = (sy + 3) * TPW + p + 3;
                        const unsigned c = W[b], n = W[b - 3 * TPW], s = W[b + 3 * TPW];
                        const unsigned e = __byte_perm(W[b + 1], W[b + 2], 0x5432);
                        const unsigned w = __byte_perm(W[b - 2], W[b - 1], 0x5432);
                        const unsigned D = vmax2(vmin2(n, s), vmin2(e, w));
                        const unsigned B = vmin2(vmax2(n, s), vmax2(e, w));
                        m = ((c + K - D) | (B + K - c)) & 0x02000200u;
                    }
                    const unsigned blo = __ballot_sync(0xffffffffu, m & 0x200u);
                    const unsigned bhi = __ballot_sync(0xffffffffu, m & 0x02000000u);
                    if (lane == 0) { s_pass[(sy * 4 + wc) * 2] = blo; s_pass[(sy * 4 + wc) * 2 + 1] = bhi; }
                } else if (lane == 0) {
                    s_pass[(sy * 4 + wc) * 2] = 0; s_pass[(sy * 4 + wc) * 2 + 1] = 0;
                }
            }
        }
        __syncthreads();
        // ---- pass bits -> queue (block scan of the popcounts, then every word owner expands its bits)
        {
            const int nw = nsr * 8;
            unsigned bits = tid < nw ? s_pass[tid] : 0u;
            const int cnt = __popc(bits);
            int inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
            if (lane == 31) s_wsum[wid] = inc;
            __syncthreads();
            int off = inc - cnt, tot = 0;
#pragma unroll
            for (int i = 0; i < NWARP; ++i) { const int s = s_wsum[i]; if (i < wid) off += s; tot += s; }
            if (tid == 0) s_qn = tot;
            const int sy = tid >> 3, wc = (tid >> 1) & 3, par = tid & 1;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                s_q[off++] = (uint16_t)((sy << 8) | ((wc * 32 + b) * 2 + par));
            }
        }
        __syncthreads();
        const int qn = s_qn;
        // ---- phase B: full segment test + score on the queued pixels; corners inside the output region are queued for NMS
        for (int i = tid; i < qn; i += NT) {
            const int e = s_q[i];
            const int sy = e >> 8, sx = e & 255;
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
            if (sc > T) {
                s_score[sy * SP + sx] = (uint8_t)(sc - 1);
                const int x = ox0 - 2 + sx;
                if (x >= max(ox0, ORBX_EDGE) && x < ox1 && sy >= 1 && sy <= y1 - y0) s_cq[atomicAdd(&s_cn, 1)] = (uint16_t)e;
            }
        }
        __syncthreads();
        // ---- phase C: 3x3 NMS of the queued corners
        const int cn = s_cn;
        for (int i = tid; i < cn; i += NT) {
            const int e = s_cq[i];
            const int sy = e >> 8, sx = e & 255;
            const uint8_t* p = &s_score[sy * SP + sx];
            const int s = p[0];
            if (s > p[-1] && s > p[1] && s > p[-SP - 1] && s > p[-SP] && s > p[-SP + 1] && s > p[SP - 1] && s > p[SP] && s > p[SP + 1]) {
                const int bit = sx - 2;
                atomicOr(&s_mask[(sy - 1) * MW + (bit >> 5)], 1u << (bit & 31));
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
                const int bit = wi * 32 + b;
                const uint32_t sc = s_score[(row + 1) * SP + bit + 2];
                dst[slot++] = (uint32_t)(ox0 + bit) | (sc << 16);
            }
            if (wi == MW - 1) s_rowcnt[row] = base + pre;
        }
    }
    __syncthreads();
    if (tid < y1 - y0) cnt_out[y0 - ORBX_EDGE + tid] = s_rowcnt[tid];
}


// ------------------------------------------------------------------------------------------------ A.6 retainBest
// Exact emulation of KeyPointsFilter::retainBest = libstdc++ std::nth_element (__introselect: median-of-3 to
// first, Hoare __unguarded_partition, insertion sort of <= 3) + bidirectional std::partition.  The resulting ORDER
// is libstdc++'s, which OpenCV's output order (and, through index tie-breaks, the matcher) depends on.
//   - ranges longer than SEL_PAR_MIN: one whole-CTA data-parallel Hoare pass (SURVEY A.6 / probe E21): with
//     L = ascending positions of !(a > piv) and R = descending positions of !(piv > a) (original contents), the
//     sequential loop swaps a[L[k]] <-> a[R[k]] for k < K, K = #{k : L[k] < R[k]}, and returns
//     cut = min(L[K], R[K-1]).  Ranks come from a block scan, K from a block reduction, swaps run in parallel.
//   - short ranges and the std::partition tail: warp-cooperative sequential emulation, scans vectorised with ballots.
// `v` may live in shared or global memory.
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

__device__ __forceinline__ int median3_pick(const Elem* v, int a, int b, int c)
{
    const float ra = v[a].response, rb = v[b].response, rc = v[c].response;
    if (ra > rb) return (rb > rc) ? b : ((ra > rc) ? c : a);
    return (ra > rc) ? a : ((rb > rc) ? c : b);
}

// warp-cooperative __introselect on [first, last) with the remaining depth budget; false = depth limit hit
__device__ bool introselect_warp(Elem* v, int first, int last, int nth, int depth, int lane)
{
    while (last - first > 3) {
        if (depth == 0) return false;
        --depth;
        const int pick = median3_pick(v, first + 1, first + (last - first) / 2, last - 1);
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
    if (lane == 0) {                                         // __insertion_sort, descending, <= 3 elements
        for (int i = first + 1; i < last; ++i) {
            const Elem val = v[i];
            int j = i;
            while (j > first && val.response > v[j - 1].response) { v[j] = v[j - 1]; --j; }
            v[j] = val;
        }
    }
    __syncwarp();
    return true;
}

// warp-cooperative bidirectional std::partition(v + m, v + len, response >= thr); returns the partition point
__device__ int partition_tail_warp(Elem* v, int m, int len, float thr, int lane)
{
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

constexpr int SEL_NT = 128;
constexpr int SEL_SMEM_ELEMS = 3072;
constexpr int SEL_PAR_MIN = 96;        // ranges up to this length are finished by one warp

struct SelShared {
    int warp_a[SEL_NT / 32], warp_b[SEL_NT / 32];
    int n, first, last, depth, ok;
};

// Whole-CTA KeyPointsFilter::retainBest(v, m).  All SEL_NT threads must call it.  PosT scratch lists Lp / Rp hold
// `len` positions each.  Returns the new length (valid on every thread); *flag |= 1 on the heap-select fallback.
template <typename PosT>
__device__ int retain_best_block(Elem* v, int len, int m, PosT* Lp, PosT* Rp, SelShared* sh, int* flag)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (m < 0 || len <= m) return len;
    if (m == 0) return 0;
    const int nth = m - 1;
    int first = 0, last = len, depth = 2 * (31 - __clz(len));
    bool ok = true;
    // ---- data-parallel Hoare passes while the range is long
    while (last - first > SEL_PAR_MIN) {
        if (depth == 0) { ok = false; break; }
        --depth;
        if (tid == 0) elem_swap(v, first, median3_pick(v, first + 1, first + (last - first) / 2, last - 1));
        __syncthreads();
        const float piv = v[first].response;
        const int lo = first + 1, n = last - lo;
        const int cs = (n + SEL_NT - 1) / SEL_NT;
        const int b = min(lo + tid * cs, last), e = min(b + cs, last);
        int cL = 0, cR = 0;
        for (int i = b; i < e; ++i) { const float r = v[i].response; cL += !(r > piv); cR += !(piv > r); }
        int iL = cL, iR = cR;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int oL = __shfl_up_sync(0xffffffffu, iL, d), oR = __shfl_up_sync(0xffffffffu, iR, d);
            if (lane >= d) { iL += oL; iR += oR; }
        }
        if (lane == 31) { sh->warp_a[wid] = iL; sh->warp_b[wid] = iR; }
        __syncthreads();
        int totL = 0, totR = 0, exL = iL - cL, exR = iR - cR;
#pragma unroll
        for (int i = 0; i < SEL_NT / 32; ++i) {
            const int a = sh->warp_a[i], c = sh->warp_b[i];
            if (i < wid) { exL += a; exR += c; }
            totL += a; totR += c;
        }
        {
            int kL = exL, kR = totR - 1 - exR;               // L: ascending rank;  R: rank counted from the right end
            for (int i = b; i < e; ++i) {
                const float r = v[i].response;
                if (!(r > piv)) Lp[kL++] = (PosT)i;
                if (!(piv > r)) Rp[kR--] = (PosT)i;
            }
        }
        __syncthreads();
        const int mm = min(totL, totR);
        int cnt = 0;
        for (int k = tid; k < mm; k += SEL_NT) cnt += ((int)Lp[k] < (int)Rp[k]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0) sh->warp_a[wid] = cnt;                // safe: all reads of warp_a finished before the last barrier
        __syncthreads();
        int K = 0;
#pragma unroll
        for (int i = 0; i < SEL_NT / 32; ++i) K += sh->warp_a[i];
        for (int k = tid; k < K; k += SEL_NT) elem_swap(v, (int)Lp[k], (int)Rp[k]);
        int cut = 0x7fffffff;
        if (K < totL) cut = (int)Lp[K];
        if (K > 0) cut = min(cut, (int)Rp[K - 1]);
        __syncthreads();
        if (cut <= nth) first = cut; else last = cut;
    }
    // ---- finish with one warp: short-range introselect + insertion sort, then the std::partition tail
    if (wid == 0) {
        int res = m;
        if (ok) ok = introselect_warp(v, first, last, nth, depth, lane);
        if (ok) res = partition_tail_warp(v, m, len, v[m - 1].response, lane);
        if (lane == 0) { sh->n = res; sh->ok = ok ? 1 : 0; }
    }
    __syncthreads();
    const int res = sh->n;
    if (!sh->ok) *flag |= 1;
    __syncthreads();
    return res;
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
// global workspace; its length goes to fincnt.  The working array (and the partition scratch) lives in shared
// memory when the level's candidate count fits, else in global memory.
__global__ void __launch_bounds__(SEL_NT) k_select(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr,
                                                   const uint32_t* __restrict__ rowcnt, const uint32_t* __restrict__ rowent,
                                                   Elem* __restrict__ work, uint32_t* __restrict__ selpos, int* __restrict__ fincnt,
                                                   int* __restrict__ status)
{
    __shared__ Elem s_v[SEL_SMEM_ELEMS];
    __shared__ uint16_t s_pos[2 * SEL_SMEM_ELEMS];
    __shared__ SelShared sh;
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
    if (lane == 31) sh.warp_a[wid] = inc;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < SEL_NT / 32; ++i) { const int s = sh.warp_a[i]; if (i < wid) woff += s; total += s; }
    const int N = total;
    const bool in_smem = N <= SEL_SMEM_ELEMS;
    Elem* v = in_smem ? s_v : gv;
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
    uint32_t* gpos = selpos + 2 * ((size_t)f * g.ws_frame + L.ws_off);
    // ---- retainBest(2 n_l) on the FAST score
    const int n1 = in_smem ? retain_best_block<uint16_t>(v, N, 2 * L.quota, s_pos, s_pos + SEL_SMEM_ELEMS, &sh, &flag)
                           : retain_best_block<uint32_t>(v, N, 2 * L.quota, gpos, gpos + N, &sh, &flag);
    // ---- Harris on the unblurred level
    const uint8_t* img = pyr + (size_t)f * g.pyr_frame + L.img_off;
    for (int i = tid; i < n1; i += SEL_NT) {
        const uint32_t pos = v[i].pos;
        v[i].response = harris_response(img, L.pitch, (int)(pos & 0xffffu), (int)(pos >> 16));
    }
    __syncthreads();
    // ---- retainBest(n_l) on Harris
    const int n2 = in_smem ? retain_best_block<uint16_t>(v, n1, L.quota, s_pos, s_pos + SEL_SMEM_ELEMS, &sh, &flag)
                           : retain_best_block<uint32_t>(v, n1, L.quota, gpos, gpos + N, &sh, &flag);
    if (tid == 0) { fincnt[f * g.nlevels + l] = n2; if (flag) atomicOr(&status[f], flag); }
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

// ------------------------------------------------------------------------------------------------ A.8 blur
// 7x7 sigma-2 Gaussian of every level in OpenCV's float sepFilter2D arithmetic:
//   row pass   acc = F(k0*p[x-3]); acc = fma(p[x-3+i], k_i, acc), i = 1..6
//   col pass   acc = F(k3*r[y]);   acc = fma(F(r[y+j] + r[y-j]), k_{3+j}, acc), j = 1..3;  out = rint(acc)
// Only the region a descriptor can sample is produced: [13, w-13) x [13, h-13) (keypoints keep 31 px from the
// border, rBRIEF reaches 18).  One thread = 4 adjacent columns walked down BLUR_RH rows with a 7-deep register
// window of row-pass values, so every input byte is loaded once per thread and converted to float once
// (exactly: PRMT into the mantissa of 2^23, one FADD).  No clamp is needed: the taps sum to < 1.
constexpr int BLUR_RH = 32;            // output rows per thread
constexpr int BLUR_NT = 128;           // 4 warps = 4 vertically stacked strips of a 128-column group
constexpr int BLUR_LO = 12;            // first produced row / column is BLUR_LO rounded to the quad grid (x) / 13 (y)

__device__ __forceinline__ float u8f(uint32_t w, uint32_t sel) { return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.0f; }

__global__ void __launch_bounds__(BLUR_NT) k_blur(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur)
{
    const int f = blockIdx.y;
    int l = 0;
#pragma unroll 1
    for (int i = 1; i < g.nlevels; ++i) if ((int)blockIdx.x >= g.L[i].blur0) l = i;
    const LevelGeom& L = g.L[l];
    const int tile = blockIdx.x - L.blur0;
    if (tile >= L.nblur) return;
    const int cg = tile % L.blur_cgs, sg = tile / L.blur_cgs;
    const int x0 = BLUR_LO + (cg * 32 + (threadIdx.x & 31)) * 4;                 // first of this thread's 4 columns (multiple of 4)
    const int ys = 13 + (sg * 4 + (threadIdx.x >> 5)) * BLUR_RH;                  // first output row of this warp's strip
    const int ye = min(ys + BLUR_RH, L.h - 13);
    if (x0 >= L.w - 13 || ys >= ye) return;
    const float k0 = __int_as_float(0x3d8fafb1), k1 = __int_as_float(0x3e06387e), k2 = __int_as_float(0x3e434a39), k3 = __int_as_float(0x3e5d4ae0);
    const uint8_t* src = pyr + (size_t)f * g.pyr_frame + L.img_off + (size_t)(ys - 3) * L.pitch + (x0 - 4);
    uint8_t* dst = blur + (size_t)f * g.pyr_frame + L.img_off + (size_t)ys * L.pitch + x0;
    float w[4][7];
#pragma unroll 1
    for (int r = 0; r < (ye - ys) + 6; ++r, src += L.pitch) {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(src));
        const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(src) + 1);
        const uint32_t c = (x0 + 4 < L.pitch) ? __ldg(reinterpret_cast<const uint32_t*>(src) + 2) : 0u;
        float p[10];                                                              // pixels x0-3 .. x0+6
        p[0] = u8f(a, 0x7651); p[1] = u8f(a, 0x7652); p[2] = u8f(a, 0x7653);
        p[3] = u8f(b, 0x7650); p[4] = u8f(b, 0x7651); p[5] = u8f(b, 0x7652); p[6] = u8f(b, 0x7653);
        p[7] = u8f(c, 0x7650); p[8] = u8f(c, 0x7651); p[9] = u8f(c, 0x7652);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = __fmul_rn(k0, p[j]);
            acc = __fmaf_rn(p[j + 1], k1, acc); acc = __fmaf_rn(p[j + 2], k2, acc); acc = __fmaf_rn(p[j + 3], k3, acc);
            acc = __fmaf_rn(p[j + 4], k2, acc); acc = __fmaf_rn(p[j + 5], k1, acc); acc = __fmaf_rn(p[j + 6], k0, acc);
#pragma unroll
            for (int i = 0; i < 6; ++i) w[j][i] = w[j][i + 1];
            w[j][6] = acc;
        }
        if (r >= 6) {
            uint32_t out = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o = __fmul_rn(k3, w[j][3]);
                o = __fmaf_rn(__fadd_rn(w[j][4], w[j][2]), k2, o);
                o = __fmaf_rn(__fadd_rn(w[j][5], w[j][1]), k1, o);
                o = __fmaf_rn(__fadd_rn(w[j][6], w[j][0]), k0, o);
                out |= (uint32_t)__float2int_rn(o) << (8 * j);
            }
            *reinterpret_cast<uint32_t*>(dst) = out;
            dst += L.pitch;
        }
    }
}

// ------------------------------------------------------------------------------------------------ A.7 / A.9 / A.10 describe
// One warp = one final keypoint (no block-level sync): IC moments over the radius-15 disc of the unblurred level
// (lane = column, shuffle reduction), fastAtan2, glibc-exact sin/cos, then the 37x37 window of the blurred level is
// staged in shared memory with aligned word loads and the 256 steered tests are sampled from it
// (lane = descriptor byte).  Keypoint records and descriptors leave as coalesced stores.
constexpr int DESC_KPB = 4;            // keypoints (warps) per CTA
constexpr int DESC_NT = DESC_KPB * 32;
constexpr int DWIN = 37, DWORDS = 12;  // staged window: 37 rows x 12 words (48 bytes >= 37 + 3 alignment slack)

__constant__ int c_umax[16];

__global__ void __launch_bounds__(DESC_NT) k_describe(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr,
                                                      const uint8_t* __restrict__ blur, const Elem* __restrict__ work,
                                                      const int* __restrict__ fincnt, const int8_t* __restrict__ pattern,
                                                      float* __restrict__ kps_out, uint8_t* __restrict__ desc_out,
                                                      int* __restrict__ counts_out, int cap)
{
    __shared__ uint32_t s_win[DESC_KPB][DWIN * DWORDS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, f = blockIdx.y;
    const int slot = blockIdx.x * DESC_KPB + wid;
    int total = 0, lvl = -1, idx = 0;
    {
        int s = slot;
        for (int l = 0; l < g.nlevels; ++l) {
            const int c = __ldg(fincnt + f * g.nlevels + l);
            total += c;
            if (lvl < 0) { if (s < c) { lvl = l; idx = s; } else s -= c; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) counts_out[f] = total;
    if (slot >= min(total, cap) || lvl < 0) return;          // warp-uniform
    const LevelGeom& L = g.L[lvl];
    const Elem e = work[(size_t)f * g.ws_frame + L.ws_off + idx];
    const int x = (int)(e.pos & 0xffffu), y = (int)(e.pos >> 16);
    const uint8_t* img = pyr + (size_t)f * g.pyr_frame + L.img_off;
    const uint8_t* bimg = blur + (size_t)f * g.pyr_frame + L.img_off;

    // ---- stage the blurred window rows y-18 .. y+18, bytes (x-18) .. (x+18), as aligned words
    const int xa = (x - 18) & ~3, sh = (x - 18) & 3;
    uint32_t* win = s_win[wid];
    for (int i = lane; i < DWIN * DWORDS; i += 32) {
        const int r = i / DWORDS, c = i - r * DWORDS;
        uint32_t v = 0;
        if (xa + c * 4 < L.pitch) v = __ldg(reinterpret_cast<const uint32_t*>(bimg + (size_t)(y - 18 + r) * L.pitch + xa) + c);
        win[i] = v;
    }
    // ---- IC moments on the unblurred level: lane <-> column u = lane - 15
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - 15, au = u < 0 ? -u : u;
        const uint8_t* p = img + (size_t)(y - 15) * L.pitch + (x + u);
#pragma unroll 1
        for (int vv = -15; vv <= 15; ++vv, p += L.pitch) {
            const int av = vv < 0 ? -vv : vv;
            if (au <= c_umax[av]) { const int I = __ldg(p); m10 += u * I; m01 += vv * I; }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, d); m01 += __shfl_xor_sync(0xffffffffu, m01, d); }
    const float ang = fast_atan2_deg((float)m01, (float)m10);
    float sn, cs;
    glibc_sincosf(__fmul_rn(ang, __int_as_float(0x3c8efa35)), &sn, &cs);
    __syncwarp();
    // ---- steered rBRIEF: lane <-> descriptor byte
    {
        const uint8_t* B = reinterpret_cast<const uint8_t*>(win) + 18 * (DWORDS * 4) + 18 + sh;
        unsigned byte = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const char4 pt = *reinterpret_cast<const char4*>(pattern + (lane * 8 + t) * 4);
            const float x0 = (float)pt.x, y0 = (float)pt.y, x1 = (float)pt.z, y1 = (float)pt.w;
            const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(y0, sn)));
            const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(y0, cs)));
            const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, cs), __fmul_rn(y1, sn)));
            const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sn), __fmul_rn(y1, cs)));
            const int t0 = B[iy0 * (DWORDS * 4) + ix0], t1 = B[iy1 * (DWORDS * 4) + ix1];
            byte |= (unsigned)(t0 < t1) << t;
        }
        const size_t o = (size_t)f * cap + slot;
        desc_out[o * 32 + lane] = (uint8_t)byte;
        if (lane < 7) {
            float val;
            switch (lane) {
                case 0: val = __fmul_rn((float)x, L.scale); break;
                case 1: val = __fmul_rn((float)y, L.scale); break;
                case 2: val = __fmul_rn(31.0f, L.scale); break;
                case 3: val = ang; break;
                case 4: val = e.response; break;
                case 5: val = __int_as_float(lvl); break;
                default: val = __int_as_float(-1); break;
            }
            kps_out[o * 7 + lane] = val;
        }
    }
}


}  // namespace orbx
