// orbx_map.cuh -- sm_100a kernels around the matcher (the callers either side of the hot path):
//
//   k_map_scatter       device-resident map-point table (descriptor + position + viewing normal + outlier flag per slot),
//                       replacing the per-call host gather of Mappoint::descriptor_ rows          src/frontend.cpp:169-184, :392-394
//   k_map_scatter_desc  descriptor rows copied device -> device out of the last extraction (descriptorsCurr_.row(idx).clone())
//   k_visibility        Frame::IsCouldObserveMappoint batched over the tracking map, stable-compacted in list order
//                                                                                                  src/frame.cpp:70-91
//   k_gather_desc       candidate rows -> the contiguous query matrix of match()                   src/frontend.cpp:183
//   k_filter_matches    min distance, max(min * ratio, 30) threshold, stable compaction            src/frontend.cpp:190-211
//   k_backproject       Frame::GetDepth + Camera::Pixel2World for new map points                   src/frame.cpp:43-67, src/camera.cpp:56-86
//
// The geometry is the reference's double arithmetic (Eigen / Sophus), restated with explicitly rounded operations in the
// reference's operation order; the reference binary itself is built -O3 -march=native (FMA contraction at the compiler's
// discretion), so parity for these rows is to 1e-9 relative, not bit-exact, and says so in the tests.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace orbx {

struct Pose {            // T_c_w as [R | t], row-major 3 x 4  (SE3 of src/frame.h)
    double m[12];
};
struct Cam {             // src/camera.cpp:27-31 (floats in the reference, promoted to double where they are used)
    double fx, fy, cx, cy;
};

__global__ void k_map_scatter(const int* __restrict__ slots, int n, const uint8_t* __restrict__ desc, const double* __restrict__ pos,
                              const double* __restrict__ nrm, const uint8_t* __restrict__ outl, uint8_t* t_desc, double* t_pos, double* t_nrm,
                              uint8_t* t_outl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slots[i];
    if (desc) {
        const uint4* src = reinterpret_cast<const uint4*>(desc) + 2 * (size_t)i;
        uint4* dst = reinterpret_cast<uint4*>(t_desc) + 2 * (size_t)s;
        dst[0] = src[0]; dst[1] = src[1];
    }
    if (pos) for (int k = 0; k < 3; ++k) t_pos[3 * (size_t)s + k] = pos[3 * (size_t)i + k];
    if (nrm) for (int k = 0; k < 3; ++k) t_nrm[3 * (size_t)s + k] = nrm[3 * (size_t)i + k];
    if (outl) t_outl[s] = outl[i];
}

// rows `kp_index[i]` of one frame's descriptor block -> table slots
__global__ void k_map_scatter_desc(const int* __restrict__ slots, const int* __restrict__ kp_index, int n, const uint8_t* __restrict__ frame_desc,
                                   uint8_t* t_desc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* src = reinterpret_cast<const uint4*>(frame_desc) + 2 * (size_t)kp_index[i];
    uint4* dst = reinterpret_cast<uint4*>(t_desc) + 2 * (size_t)slots[i];
    dst[0] = src[0]; dst[1] = src[1];
}

// src/frame.cpp:70-91 for one map point
__device__ __forceinline__ bool could_observe(const Pose& T, const Cam& K, int cols, int rows, const double* p, const double* nv)
{
    // Camera::World2Camera: T_c_w * p_w = R p + t
    double pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        pc[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[4 * r], p[0]), __dmul_rn(T.m[4 * r + 1], p[1])), __dmul_rn(T.m[4 * r + 2], p[2])), T.m[4 * r + 3]);
    if (pc[2] < 0.0) return false;
    // Camera::Camera2Pixel: fx * x / z + cx
    const double u = __dadd_rn(__ddiv_rn(__dmul_rn(K.fx, pc[0]), pc[2]), K.cx);
    const double v = __dadd_rn(__ddiv_rn(__dmul_rn(K.fy, pc[1]), pc[2]), K.cy);
    if (u < 0.0 || u >= (double)cols || v < 0.0 || v >= (double)rows) return false;
    // camera centre = T_c_w.inverse().translation() = -R^T t
    double d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double c = -__dadd_rn(__dadd_rn(__dmul_rn(T.m[k], T.m[3]), __dmul_rn(T.m[4 + k], T.m[7])), __dmul_rn(T.m[8 + k], T.m[11]));
        d[k] = __dsub_rn(p[k], c);
    }
    const double n2 = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
    const double nn = __dsqrt_rn(n2);
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) dot = __dadd_rn(dot, __dmul_rn(__ddiv_rn(d[k], nn), nv[k]));
    const double angle = acos(dot);
    return !(angle > 0.52359877559829887308);                // M_PI / 6 ; NaN (zero normal, |dot| > 1) is not "> pi/6": kept, as in the reference
}

// ONE CTA: candidates = list entries whose point is not an outlier and could be observed, in list order (stable).
constexpr int VIS_NT = 1024;
__global__ void __launch_bounds__(VIS_NT) k_visibility(const int* __restrict__ slots, int m, const Pose T, const Cam K, int cols, int rows,
                                                       const double* __restrict__ t_pos, const double* __restrict__ t_nrm,
                                                       const uint8_t* __restrict__ t_outl, int* __restrict__ cand, int* __restrict__ n_cand)
{
    __shared__ int s_warp[VIS_NT / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < m; base += VIS_NT) {
        const int i = base + tid;
        bool keep = false;
        if (i < m) {
            const int s = slots[i];
            keep = !t_outl[s] && could_observe(T, K, cols, rows, t_pos + 3 * (size_t)s, t_nrm + 3 * (size_t)s);
        }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < VIS_NT / 32; ++w) { const int c = s_warp[w]; if (w < wid) before += c; total += c; }
        const int out0 = s_base;
        if (keep) cand[out0 + before + __popc(b & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (tid == 0) s_base = out0 + total;
        __syncthreads();
    }
    if (tid == 0) *n_cand = s_base;
}

// query row i = table row of list entry cand[i]
__global__ void k_gather_desc(const int* __restrict__ slots, const int* __restrict__ cand, int n, const uint8_t* __restrict__ t_desc, uint8_t* q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* src = reinterpret_cast<const uint4*>(t_desc) + 2 * (size_t)slots[cand[i]];
    uint4* dst = reinterpret_cast<uint4*>(q) + 2 * (size_t)i;
    dst[0] = src[0]; dst[1] = src[1];
}

// ONE CTA: src/frontend.cpp:190-211.  min over all matches, max_dis = max<float>(min * ratio, 30), keep distance <= max_dis
// in order.  Matches against an empty train set (trainIdx < 0) are dropped and do not enter the minimum.
__global__ void __launch_bounds__(VIS_NT) k_filter_matches(const int4* __restrict__ in, int n, float ratio, int4* __restrict__ out,
                                                           int* __restrict__ n_out, float* __restrict__ minmax)
{
    __shared__ float s_min[VIS_NT / 32];
    __shared__ int s_warp[VIS_NT / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float mn = __int_as_float(0x7f800000);
    for (int i = tid; i < n; i += VIS_NT) { const int4 r = in[i]; if (r.y >= 0) mn = fminf(mn, __int_as_float(r.w)); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    if (lane == 0) s_min[wid] = mn;
    if (tid == 0) s_base = 0;
    __syncthreads();
    mn = s_min[0];
    for (int w = 1; w < VIS_NT / 32; ++w) mn = fminf(mn, s_min[w]);
    const float mx = fmaxf(__fmul_rn(mn, ratio), 30.0f);
    for (int base = 0; base < n; base += VIS_NT) {
        const int i = base + tid;
        int4 r = make_int4(0, -1, 0, 0);
        if (i < n) r = in[i];
        const bool keep = i < n && r.y >= 0 && __int_as_float(r.w) <= mx;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < VIS_NT / 32; ++w) { const int c = s_warp[w]; if (w < wid) before += c; total += c; }
        const int out0 = s_base;
        if (keep) out[out0 + before + __popc(b & ((1u << lane) - 1u))] = r;
        __syncthreads();
        if (tid == 0) s_base = out0 + total;
        __syncthreads();
    }
    if (tid == 0) { *n_out = s_base; minmax[0] = mn; minmax[1] = mx; }
}

// src/frame.cpp:43-67 (GetDepth: the pixel, then its 4-neighbours left, up, right, down) + src/camera.cpp:56-86
// (Pixel2Camera, Camera2World = T_c_w.inverse() * p_c).  Pixels whose depth lookup would leave the image are invalid
// (the reference reads out of bounds there).
__global__ void k_backproject(const float* __restrict__ kps /*7 floats per keypoint*/, int n, const uint16_t* __restrict__ depth, int w, int h,
                              size_t step_elems, double depth_scale, const Cam K, const Pose T, double* __restrict__ pos, uint8_t* __restrict__ valid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = kps[7 * (size_t)i], py = kps[7 * (size_t)i + 1];
    const int x = __float2int_rn(px), y = __float2int_rn(py);      // cvRound
    unsigned d = 0;
    bool ok = x >= 0 && y >= 0 && x < w && y < h;
    if (ok) {
        d = depth[(size_t)y * step_elems + x];
        if (d == 0) {
            const int dx[4] = {-1, 0, 1, 0}, dy[4] = {0, -1, 0, 1};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xx = x + dx[k], yy = y + dy[k];
                if (d == 0 && xx >= 0 && yy >= 0 && xx < w && yy < h) d = depth[(size_t)yy * step_elems + xx];
            }
        }
    }
    ok = ok && d != 0;
    double out[3] = {0.0, 0.0, 0.0};
    if (ok) {
        const double z = __ddiv_rn((double)d, depth_scale);
        double pc[3];
        pc[0] = __ddiv_rn(__dmul_rn(__dsub_rn((double)px, K.cx), z), K.fx);
        pc[1] = __ddiv_rn(__dmul_rn(__dsub_rn((double)py, K.cy), z), K.fy);
        pc[2] = z;
        // T_c_w.inverse() * p_c = R^T p_c + (-R^T t)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double ti = -__dadd_rn(__dadd_rn(__dmul_rn(T.m[k], T.m[3]), __dmul_rn(T.m[4 + k], T.m[7])), __dmul_rn(T.m[8 + k], T.m[11]));
            out[k] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[k], pc[0]), __dmul_rn(T.m[4 + k], pc[1])), __dmul_rn(T.m[8 + k], pc[2])), ti);
        }
    }
    pos[3 * (size_t)i] = out[0]; pos[3 * (size_t)i + 1] = out[1]; pos[3 * (size_t)i + 2] = out[2];
    valid[i] = ok ? 1 : 0;
}

}  // namespace orbx
