// orbx.cu -- host side of liborbx.so: context, device-memory layout, kernel launches and the C-ABI of include/orbx.h.
//
// The two operator calls of the reference front-end that this library stands in for:
//   cv::ORB::detectAndCompute            src/frontend.cpp:153   -> orbx_detect_and_compute*
//   cv::DescriptorMatcher::match         src/frontend.cpp:187   -> orbx_match_hamming*
// There is no CPU fallback anywhere in this file: every entry point either runs the sm_100a kernels or fails.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/orbx.h"
#include "orbx_geom.h"
#include "orbx_kernels.cuh"
#include "orbx_match.cuh"
#include "orbx_fast.cuh"
#include "orbx_pyr.cuh"
#include "orbx_desc.cuh"
#include "orbx_select.cuh"
#include "orbx_map.cuh"
#include "orbx_stage.h"
#include <unordered_map>

using namespace orbx;

namespace {

const int8_t k_pattern_host[256 * 4] = {
#include "brief_pattern.inc"
};

constexpr int FAST_R = 16, FAST_NT = 256;
constexpr int FAST_NWARP = 4;             // warps per CTA of k_fast_warp (independent units; 7 CTAs per SM by shared memory)
static_assert(FAST_R == FW_R, "band height");
constexpr int MAX_LANES = 4;               // concurrent frame-range pipelines of one device-resident extraction call
constexpr int HOST_MAX_LANES = 12, HOST_DEFAULT_LANES = 8;         // host-buffer calls: more, shorter ranges shrink the un-overlapped head (first upload) and tail
constexpr int LANE_MIN_FRAMES = 64;        // a lane must still fill the GPU on its own
constexpr int HOST_LANE_MIN_FRAMES = 16;   // host-buffer batches: lanes mainly overlap PCIe copies with kernels
constexpr int N_STAGES = 7;
const char* const k_stage_names[N_STAGES] = {"gray", "pyramid", "fast_nms", "select_harris", "blur", "describe", "match"};

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// NVTX ranges around every stage's launches and around the C-ABI calls (the reference's only timer is the wall clock around
// AddFrame, app/run_vo.cpp:104-109); header-only NVTX 3: free when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
struct NvtxStages {                      // one open range at a time: next() closes the previous stage's
    bool open = false;
    void next(const char* name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = true; }
    ~NvtxStages() { if (open) nvtxRangePop(); }
};

struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct orbx_ctx {
    int device = 0, nfeatures = 0, nlevels = 0, max_w = 0, max_h = 0, max_batch = 0;
    float scale_factor = 1.2f;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;

    Geom geom{};             // geometry of the current frame size
    FastMaps fmaps{};        // TMA tensor maps (x, y, frame) of the pyramid levels for the current geometry
    PyrMaps pmaps{};         // ... of every level as the SOURCE of the next one (k_pyr_tma boxes)
    DescMaps dmaps{};        // ... of every pyramid / blurred level with k_describe_tma's window boxes
    Geom geom_max{};         // geometry of (max_w, max_h): sizes the buffers
    int geom_w = 0, geom_h = 0;
    size_t tabs_len = 0;

    Buf pyr, blur, rowcnt, rowent, work, selpos, fincnt, selcnt, status, tabs, pattern;
    Buf in, kps, desc, counts;               // host-path staging on the device
    int out_cap = 0;
    const uint8_t* view_desc = nullptr;      // device descriptors / counts of the frames the last host call (or collect) produced
    const int* view_counts = nullptr;
    int view_frames = 0;
    Buf mq, mt, mbest, msecond, mkeys, mstatus, mcounts, mtrace;
    int* h_small = nullptr;                  // pinned: counts[max_batch] + status[max_batch] + 1
    uint8_t* h_stage_in = nullptr; size_t h_stage_in_bytes = 0;     // pinned staging of pageable caller frames (device layout)
    uint8_t* h_stage_out = nullptr; size_t h_stage_out_bytes = 0;   // ... of results bound for pageable caller buffers
    HostStager* stager = nullptr;            // copier threads, created by the first call that needs them
    int last_batch = 0;

    // device-resident map-point table: slot-addressed columns + the host's id -> slot index
    Buf t_desc, t_pos, t_nrm, t_outl, t_slots, t_cand, t_ncand, t_q, t_in, t_best, t_filtered, t_minmax, t_aux;
    int map_capacity = 0;
    std::unordered_map<long long, int> map_slot;
    std::vector<int> map_free;
    int map_next = 0;

    // asynchronous single-frame pipeline (orbx_submit_frame / orbx_collect_frame): two slots of pinned host staging and
    // device outputs, so the host can prepare / consume one frame while the GPU works on the other
    struct AsyncSlot {
        uint8_t* h_in = nullptr; size_t h_in_bytes = 0;
        uint8_t* h_out = nullptr; size_t h_out_bytes = 0;   // [counts 16 B][status 16 B][kps cap*28][desc cap*32]
        Buf d_kps, d_desc, d_counts;
        cudaEvent_t done = nullptr;
        int cap = 0;
        bool busy = false;
    } aslot[2];
    int a_head = 0, a_inflight = 0;          // oldest in-flight slot, number in flight

    bool profiling = false;
    int force_kernels = -1;                   // orbx_debug_force_kernels: -1 by launch size, 0 warp-private TMA kernels, 1 CTA-cooperative kernels
    cudaEvent_t ev[N_STAGES + 2] = {};   // 0..6 bracket the six extraction stages, 7..8 the matcher
    cudaStream_t lane[HOST_MAX_LANES] = {};   // extra frame-range pipelines (lane 0 is `stream`)
    cudaStream_t side[HOST_MAX_LANES] = {};   // per lane: high-priority stream of the small pyramid levels + their FAST bands
    cudaEvent_t ev_sfork[HOST_MAX_LANES] = {}, ev_sjoin[HOST_MAX_LANES] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[HOST_MAX_LANES] = {};
    float stage_ms[N_STAGES] = {};
    bool stage_valid[N_STAGES] = {};
};

namespace {

int fail(orbx_ctx* c, int code, const char* fmt, const char* a = "", const char* b = "")
{
    if (c) {
        char buf[512];
        snprintf(buf, sizeof buf, fmt, a, b);
        c->err = buf;
    }
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) return fail(c, ORBX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int ensure(orbx_ctx* c, Buf& b, size_t bytes)
{
    if (b.bytes >= bytes && b.p) return ORBX_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
    bytes = std::max<size_t>(bytes, 256);
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(c, ORBX_E_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e));
    b.bytes = bytes;
    return ORBX_OK;
}

// ---- geometry, identical arithmetic to OpenCV: float scale chain, cvRound sizes, float quotas
void level_sizes(int w, int h, int nlevels, float sf, int* ws, int* hs, float* sc)
{
    for (int l = 0; l < nlevels; ++l) {
        const float s = (float)pow((double)sf, (double)l);
        const float inv = 1.0f / s;
        sc[l] = s;
        ws[l] = (int)lrintf((float)w * inv);
        hs[l] = (int)lrintf((float)h * inv);
    }
}

void level_quotas(int nfeatures, float sf, int nlevels, int* q)
{
    const float factor = (float)(1.0 / (double)sf);
    const float one_minus = 1.0f - factor;
    const float num = (float)nfeatures * one_minus;
    const float den = 1.0f - (float)pow((double)factor, (double)nlevels);
    float nd = num / den;
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        q[l] = (int)lrintf(nd);
        sum += q[l];
        nd = nd * factor;
    }
    q[nlevels - 1] = std::max(nfeatures - sum, 0);
}

// INTER_LINEAR_EXACT taps of one axis, packed i0 | c1 << 16 (c1 in 8.8 fixed point, c0 = 256 - c1)
void resize_taps(int s, int d, uint32_t* out)
{
    const double inv_scale = (double)d / (double)s;
    const double scale = 1.0 / inv_scale;
    for (int x = 0; x < d; ++x) {
        const double f = scale * ((double)x + 0.5) - 0.5;
        const int i = (int)floor(f);
        uint32_t i0, c1;
        if (i < 0 || s <= 1) { i0 = 0; c1 = 0; }
        else if (i >= s - 1) { i0 = (uint32_t)(s - 1); c1 = 0; }
        else { i0 = (uint32_t)i; c1 = (uint32_t)lrint((f - (double)i) * 256.0); }
        out[x] = i0 | (c1 << 16);
    }
}

void build_geom(const orbx_ctx* c, int w, int h, Geom* g, std::vector<uint32_t>* tabs)
{
    memset(g, 0, sizeof *g);
    int ws[ORBX_LEVELS_MAX], hs[ORBX_LEVELS_MAX], q[ORBX_LEVELS_MAX];
    float sc[ORBX_LEVELS_MAX];
    level_sizes(w, h, c->nlevels, c->scale_factor, ws, hs, sc);
    level_quotas(c->nfeatures, c->scale_factor, c->nlevels, q);
    g->nlevels = c->nlevels; g->w = w; g->h = h; g->band_rows = FAST_R;
    size_t pyr = 0, cnt = 0, ent = 0, wsz = 0, tab = 0;
    int bands = 0, blurs = 0, hblks = 0, selh = 64;
    for (int l = 0; l < c->nlevels; ++l) {
        LevelGeom& L = g->L[l];
        L.w = std::max(ws[l], 0); L.h = std::max(hs[l], 0);
        L.pitch = (int)round_up((size_t)std::max(L.w, 1), 16);
        L.quota = q[l]; L.scale = sc[l];
        L.in_w = std::max(L.w - 2 * ORBX_EDGE, 0); L.in_h = std::max(L.h - 2 * ORBX_EDGE, 0);
        if (L.in_w == 0 || L.in_h == 0) { L.in_w = 0; L.in_h = 0; }
        L.ent_pitch = (int)round_up((size_t)(L.in_w + 1) / 2 + 1, 8);
        L.ws_cap = std::max(((L.in_w + 1) / 2) * ((L.in_h + 1) / 2), 1);
        L.band0 = bands; L.nbands = (L.in_h + FAST_R - 1) / FAST_R; bands += L.nbands;
        if (L.in_w > 0) {                                    // blur work items: 8-column groups x strips of blur_rh rows, BLUR_NT per CTA
            L.blur_cgs = (L.w - 13 - BLUR_LO + 7) / 8;
            // strip height: a multiple of 7 (the kernel's unrolled row window) that wastes the fewest rows on this level
            const int rows = L.h - 26;
            long best = -1;
            for (int rh = 35; rh <= 84; rh += 7) {
                const long cost = (long)((rows + rh - 1) / rh) * (rh + 6);
                if (best < 0 || cost < best) { best = cost; L.blur_rh = rh; }
            }
            L.nblur = (L.blur_cgs * ((rows + L.blur_rh - 1) / L.blur_rh) + BLUR_NT - 1) / BLUR_NT;
        }
        L.blur0 = blurs; blurs += L.nblur;
        L.hblk0 = hblks; L.hblk = (2 * L.quota * 5 / 4 + 32 + HARRIS_NT - 1) / HARRIS_NT; hblks += L.hblk;
        selh = std::max(selh, 2 * L.quota * 5 / 4 + 64);
        L.img_off = pyr; pyr += round_up((size_t)L.pitch * std::max(L.h, 1), 256);
        L.cnt_off = cnt; cnt += round_up((size_t)std::max(L.in_h, 1), 8);
        L.ent_off = ent; ent += (size_t)L.ent_pitch * std::max(L.in_h, 1);
        L.ws_off = wsz; wsz += round_up((size_t)L.ws_cap, 4);
        if (l > 0) { L.xtab = (uint32_t)tab; tab += L.w; L.ytab = (uint32_t)tab; tab += L.h; }
    }
    g->total_bands = bands; g->total_blur = blurs; g->total_hblk = hblks;
    g->selh_elems = std::min((selh + 63) / 64 * 64, 16384);
    g->pyr_frame = round_up(pyr, 256); g->cnt_frame = cnt; g->ent_frame = ent; g->ws_frame = wsz;
    if (tabs) {
        tabs->assign(std::max<size_t>(tab, 1), 0);
        for (int l = 1; l < c->nlevels; ++l) {
            const LevelGeom& L = g->L[l];
            const LevelGeom& S = g->L[l - 1];
            if (L.w > 0 && L.h > 0 && S.w > 0 && S.h > 0) {
                resize_taps(S.w, L.w, tabs->data() + L.xtab);
                resize_taps(S.h, L.h, tabs->data() + L.ytab);
                int span = 0;                                // how far a thread's 4 outputs reach from its aligned first source byte
                for (int x = 0; x < L.w; x += 4) {
                    const int a = (int)((*tabs)[L.xtab + x] & 0xffffu) & ~3;
                    const int xl = std::min(x + 3, L.w - 1);
                    span = std::max(span, (int)((*tabs)[L.xtab + xl] & 0xffffu) - a);
                }
                g->L[l].xspan = span;
                // k_pyr_tma: a lane's 4 outputs must read at most 7 consecutive source bytes, and the source box of a
                // (128-column tile, 16-row strip) must fit a TMA box (<= 256 per dimension, 16-byte aligned first column)
                LevelGeom& D = g->L[l];
                const uint32_t* xt = tabs->data() + D.xtab;
                const uint32_t* yt = tabs->data() + D.ytab;
                int quad_span = 0, bw = 0, bh = 0;
                for (int x = 0; x < D.w; x += 4)
                    quad_span = std::max(quad_span, (int)(xt[std::min(x + 3, D.w - 1)] & 0xffffu) - (int)(xt[x] & 0xffffu));
                D.pt_ncx = (D.pitch + PT_CW - 1) / PT_CW;
                for (int cx = 0; cx < D.pt_ncx; ++cx) {
                    const int xl = std::min(cx * PT_CW + PT_CW - 1, D.w - 1);
                    const int X0 = (int)(xt[cx * PT_CW] & 0xffffu) & ~15;
                    bw = std::max(bw, (int)(xt[xl] & 0xffffu) + 2 - X0);
                }
                const int nstrips = (D.h + PT_RH - 1) / PT_RH;
                for (int st = 0; st < nstrips; ++st) {
                    const int ye = std::min(st * PT_RH + PT_RH, D.h);
                    bh = std::max(bh, std::min((int)(yt[ye - 1] & 0xffffu) + 1, S.h - 1) - (int)(yt[st * PT_RH] & 0xffffu) + 1);
                }
                D.pt_bw = (int)round_up((size_t)bw, 16); D.pt_bh = bh;
                D.pt_k = 4;
                D.pt_ntask = D.pt_ncx * ((nstrips + D.pt_k - 1) / D.pt_k);
                D.pt_ok = quad_span <= 6 && D.pt_bw <= 256 && D.pt_bh <= 256 && PT_NWARP * pt_warp_bytes(D.pt_bw, D.pt_bh) + 128 <= 160 * 1024;
            }
        }
    }
}

// One 3-D u8 tensor map (x = pitch, y = rows, z = frame) per pyramid level over the context's pyramid buffer: k_fast_warp's
// image tiles are TMA boxes of FW_TP x FW_TR x 1 (out-of-range parts are zero-filled by the hardware).
int build_fast_maps(orbx_ctx* c)
{
    typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static enc_t enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
            return fail(c, ORBX_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        enc = (enc_t)fn;
    }
    const Geom& g = c->geom;
    memset(&c->fmaps, 0, sizeof c->fmaps);
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.L[l];
        if (L.nbands <= 0 || L.in_w <= 0) continue;
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.h, (cuuint64_t)c->max_batch};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.pyr_frame};
        const cuuint32_t box[3] = {FW_TP, FW_TR, 1}, es[3] = {1, 1, 1};
        const CUresult r = enc(&c->fmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (uint8_t*)c->pyr.p + L.img_off, dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, ORBX_E_CUDA, "cuTensorMapEncodeTiled failed for a pyramid level");
    }
    memset(&c->pmaps, 0, sizeof c->pmaps);
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom& D = g.L[l];
        const LevelGeom& S = g.L[l - 1];
        if (!D.pt_ok) continue;
        const cuuint64_t dims[3] = {(cuuint64_t)S.pitch, (cuuint64_t)S.h, (cuuint64_t)c->max_batch};
        const cuuint64_t strides[2] = {(cuuint64_t)S.pitch, (cuuint64_t)g.pyr_frame};
        const cuuint32_t box[3] = {(cuuint32_t)D.pt_bw, (cuuint32_t)D.pt_bh, 1}, es[3] = {1, 1, 1};
        const CUresult r = enc(&c->pmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (uint8_t*)c->pyr.p + S.img_off, dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, ORBX_E_CUDA, "cuTensorMapEncodeTiled failed for a pyramid source level");
    }
    memset(&c->dmaps, 0, sizeof c->dmaps);
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.L[l];
        if (L.in_w <= 0) continue;
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.h, (cuuint64_t)c->max_batch};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.pyr_frame};
        const cuuint32_t box_ic[3] = {DS_ICW, DS_ICH, 1}, box_bl[3] = {DS_BLW, DS_BLH, 1}, es[3] = {1, 1, 1};
        if (enc(&c->dmaps.ic[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (uint8_t*)c->pyr.p + L.img_off, dims, strides, box_ic, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
            enc(&c->dmaps.bl[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (uint8_t*)c->blur.p + L.img_off, dims, strides, box_bl, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(c, ORBX_E_CUDA, "cuTensorMapEncodeTiled failed for a descriptor window map");
    }
    return ORBX_OK;
}

int set_geometry(orbx_ctx* c, int w, int h)
{
    if (c->geom_w == w && c->geom_h == h) return ORBX_OK;
    std::vector<uint32_t> tabs;
    build_geom(c, w, h, &c->geom, &tabs);
    for (int l = 1; l < c->nlevels; ++l)
        if (c->geom.L[l].xspan > 11) { c->geom_w = c->geom_h = -1; return fail(c, ORBX_E_UNSUPPORTED, "scale_factor too large for the pyramid kernels (4 outputs must span <= 12 source bytes: about 2.6)"); }
    int rc = ensure(c, c->tabs, tabs.size() * 4);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->tabs.p, tabs.data(), tabs.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));       // `tabs` is a local
    if ((rc = build_fast_maps(c))) return rc;
    c->geom_w = w; c->geom_h = h;
    return ORBX_OK;
}

void stage_mark(orbx_ctx* c, int i)
{
    if (c->profiling) cudaEventRecord(c->ev[i], c->stream);
}

// The extraction pipeline on device-resident frames; everything asynchronous on c->stream.
// Frame range of host-buffer lane k.  The ranges taper towards both ends (weights 1, 2, 4, 4, ..., 4, 2, 1): the first
// upload has nothing to hide under and the last range's kernels + download have nothing left to hide, so both are short.
static void host_lane_range(int batch, int lanes, int k, int* f0, int* f1)
{
    auto wgt = [&](int i) { return 1 << std::min(2, std::min(i, lanes - 1 - i)); };
    long tot = 0, before = 0;
    for (int i = 0; i < lanes; ++i) { tot += wgt(i); if (i < k) before += wgt(i); }
    *f0 = (int)((long)batch * before / tot);
    *f1 = (int)((long)batch * (before + wgt(k)) / tot);
}

// The kernel sequence for frames [f0, f0 + nb) on stream `st`.  Every buffer is frame-major, so a frame range is
// just a base-pointer offset.
int run_extract_range(orbx_ctx* c, cudaStream_t st, int lane_id, bool marks, int f0, int nb,
                      const uint8_t* d_imgs, size_t step, size_t frame_stride, int channels, float* d_kps, uint8_t* d_desc, int cap,
                      int* d_counts)
{
    const Geom& g = c->geom;
    const size_t F = (size_t)f0;
    d_imgs += F * frame_stride;
    uint8_t* pyr = (uint8_t*)c->pyr.p + F * g.pyr_frame;
    uint8_t* blur = (uint8_t*)c->blur.p + F * g.pyr_frame;
    uint32_t* rowcnt = (uint32_t*)c->rowcnt.p + F * g.cnt_frame;
    uint32_t* rowent = (uint32_t*)c->rowent.p + F * g.ent_frame;
    Elem* work = (Elem*)c->work.p + F * g.ws_frame;
    uint32_t* selpos = (uint32_t*)c->selpos.p + 2 * F * g.ws_frame;
    int* fincnt = (int*)c->fincnt.p + F * g.nlevels;
    int* selcnt = (int*)c->selcnt.p + F * g.nlevels;
    int* status = (int*)c->status.p + F;
    d_kps += F * cap * 7; d_desc += F * cap * 32; d_counts += F;
    const uint32_t* tabs = (const uint32_t*)c->tabs.p;
    const unsigned B = (unsigned)nb;
    cudaStream_t side = c->side[lane_id];
    cudaEvent_t side_fork = c->ev_sfork[lane_id], side_join = c->ev_sjoin[lane_id];

    NvtxStages nv;
    nv.next("orbx:gray");
    if (marks) stage_mark(c, 0);
    {
        const int aligned16 = ((reinterpret_cast<uintptr_t>(d_imgs) | step | frame_stride) & 15) == 0;
        const dim3 blk(32, 8);
        const dim3 grd((unsigned)((g.L[0].pitch / 16 + 31) / 32), (unsigned)((g.L[0].h + 7) / 8), B);
        if (channels == 3) k_gray<3><<<grd, blk, 0, st>>>(d_imgs, frame_stride, step, aligned16, g, pyr);
        else               k_gray<1><<<grd, blk, 0, st>>>(d_imgs, frame_stride, step, aligned16, g, pyr);
        ++c->launches;
    }
    if (marks) stage_mark(c, 1);
    // Kernel choice by the size of the launch: the warp-private TMA kernels (one warp walks a whole band / a stack of strips)
    // win once there are enough warps to fill the GPU; below about 26 VGA frames' worth of pixels the CTA-cooperative
    // kernels have the shorter critical path (one 640x480 frame: pyramid + FAST 0.077 ms against 0.115 ms), which is what the
    // sequential VO loop (one frame per call, src/frontend.cpp:98-108) feels.  ORBX_*_OLD = 1 / 0 forces one or the other.
    const bool small_launch = (long)nb * g.w * g.h < 8000000L;
    static const int env_old_pyr = getenv("ORBX_PYR_OLD") ? atoi(getenv("ORBX_PYR_OLD")) : -1;
    static const int env_old_fast = getenv("ORBX_FAST_OLD") ? atoi(getenv("ORBX_FAST_OLD")) : -1;
    const bool old_pyr = c->force_kernels >= 0 ? c->force_kernels != 0 : env_old_pyr >= 0 ? env_old_pyr != 0 : small_launch;
    const bool old_fast = c->force_kernels >= 0 ? c->force_kernels != 0 : env_old_fast >= 0 ? env_old_fast != 0 : small_launch;
    auto launch_pyr = [&](int l, cudaStream_t s) {
        if (g.L[l].w <= 0 || g.L[l].h <= 0) return;
        const int nitems = (g.L[l].pitch / 8) * ((g.L[l].h + PYR_RH - 1) / PYR_RH);     // (column octet, row strip) work items
        const dim3 grd((unsigned)((nitems + PYR_NT - 1) / PYR_NT), B);
        if (g.L[l].pt_ok && !old_pyr)
            k_pyr_tma<<<dim3((unsigned)((g.L[l].pt_ntask + PT_NWARP - 1) / PT_NWARP), B), PT_NWARP * 32,
                        PT_NWARP * pt_warp_bytes(g.L[l].pt_bw, g.L[l].pt_bh) + 128, s>>>(g, c->pmaps.m[l], l, f0, (uint8_t*)c->pyr.p, tabs, status);
        else if (g.L[l].xspan <= 7) k_pyr_down<false><<<grd, PYR_NT, 0, s>>>(g, l, pyr, tabs);
        else                   k_pyr_down<true><<<grd, PYR_NT, 0, s>>>(g, l, pyr, tabs);
        ++c->launches;
    };
    // The blur (FMA-pipe-bound, needs only the pyramid) follows the small levels on the side stream, i.e. it also runs under the
    // large levels' FAST launch (ALU-pipe-bound): 0.959 -> 0.944 ms per 256 VGA frames.  Capping the occupancy of either kernel
    // so that both are resident in fixed proportions was measured and loses (FAST needs its 7 CTAs per SM).
    static const int env_co = getenv("ORBX_CO") ? atoi(getenv("ORBX_CO")) : 1;
    bool blur_done = false;
    auto launch_fast = [&](int band_lo, int band_hi, cudaStream_t s) {
        if (band_hi <= band_lo) return;
        if (old_fast) {
            if (band_lo == 0) k_fast_bands<FAST_R, FAST_NT><<<dim3((unsigned)g.total_bands, B), FAST_NT, 0, s>>>(g, pyr, rowcnt, rowent);
        } else
            k_fast_warp<FAST_NWARP><<<dim3((unsigned)((band_hi - band_lo + FAST_NWARP - 1) / FAST_NWARP), B), FAST_NWARP * 32, FAST_NWARP * FW_WARP_BYTES + 128, s>>>(
                g, c->fmaps, f0, band_lo, band_hi, rowcnt, rowent, status);
        ++c->launches;
    };
    // The small upper levels are a chain of short, latency-bound launches (at 640x480 levels 4..7 take as long as levels 1..3
    // for a fifth of the pixels).  They leave the critical path: from level `lt` on, the pyramid chain and the FAST bands of
    // those levels run on a HIGH-PRIORITY side stream (their few CTAs take the next free slots) underneath the FAST launch of the
    // large levels, which is issue-bound and has slots to spare for them only in time, not in space.
    int lt = g.nlevels;                                      // first level of the side chain
    {
        static const int tail_px = getenv("ORBX_TAIL_PX") ? atoi(getenv("ORBX_TAIL_PX")) : 100000;   // levels below this many pixels are "small"; 0: no side chain
        for (int l = g.nlevels - 1; l >= 2 && (long)g.L[l].w * g.L[l].h < tail_px; --l) lt = l;
        if (old_fast || g.total_bands <= 0 || marks) lt = g.nlevels;   // (stage timers on: one stream, clean per-stage times)
    }
    nv.next("orbx:pyramid");
    for (int l = 1; l < lt; ++l) launch_pyr(l, st);
    if (marks) stage_mark(c, 2);
    nv.next("orbx:fast_nms (+ small pyramid levels and blur on the side stream)");
    if (lt < g.nlevels) {
        CU(cudaEventRecord(side_fork, st));
        CU(cudaStreamWaitEvent(side, side_fork, 0));
        for (int l = lt; l < g.nlevels; ++l) launch_pyr(l, side);
        launch_fast(g.L[lt].band0, g.total_bands, side);
        if (env_co && g.total_blur > 0) {
            k_blur<<<dim3((unsigned)g.total_blur, B), BLUR_NT, 0, side>>>(g, pyr, blur);
            ++c->launches; blur_done = true;
        }
        CU(cudaEventRecord(side_join, side));
        launch_fast(0, g.L[lt].band0, st);
        CU(cudaStreamWaitEvent(st, side_join, 0));
    } else if (g.total_bands > 0) launch_fast(0, g.total_bands, st);
    nv.next("orbx:select_harris");
    if (marks) stage_mark(c, 3);
    // Experiment kept behind ORBX_OVERLAP=1 (and only when the blur did not already go to the side stream): the selection (three
    // launches of single-warp work per (frame, level), latency-bound) on the high-priority side stream UNDER the blur, which needs
    // only the pyramid.  Measured slower than the sequential order (0.983 ms against 0.962 ms per 256 VGA frames): the selection's
    // dependent launches wait for slots the blur's CTAs hold.
    static const int env_overlap = getenv("ORBX_OVERLAP") ? atoi(getenv("ORBX_OVERLAP")) : 0;
    const bool overlap_sel = env_overlap && !marks && g.total_blur > 0 && !blur_done;
    cudaStream_t ss = overlap_sel ? side : st;
    if (overlap_sel) { CU(cudaEventRecord(side_fork, st)); CU(cudaStreamWaitEvent(side, side_fork, 0)); }
    {
        // Selection kernel family by problem size: one WARP per (frame, level) when the candidate lists are short and there are
        // thousands of them (256 VGA frames: 0.116 ms against 0.17 ms), one CTA per (frame, level) when the lists are long -- 512
        // threads wide there (64 frames of 1920x1080: 0.30 ms against 0.43 ms with 128 threads and 0.65 ms with warps;
        // 3840x2160: 1.21 against 1.88 and 2.57 ms).
        static const int env_old_sel = getenv("ORBX_SELECT_OLD") ? atoi(getenv("ORBX_SELECT_OLD")) : -1;
        const bool old_sel = c->force_kernels >= 0 ? c->force_kernels != 0 : env_old_sel >= 0 ? env_old_sel != 0 : (long)g.w * g.h > 600000L;
        static const int env_sel_nt = getenv("ORBX_SELECT_NT") ? atoi(getenv("ORBX_SELECT_NT")) : 0;   // A/B timing only
        const bool wide = env_sel_nt ? env_sel_nt > SEL_NT : (long)g.w * g.h > 600000L;
        if (old_sel) {
            if (wide) k_select<SEL_NT_WIDE><<<dim3(B, (unsigned)g.nlevels), SEL_NT_WIDE, 0, ss>>>(g, pyr, rowcnt, rowent, work, selpos, fincnt, status);
            else      k_select<SEL_NT><<<dim3(B, (unsigned)g.nlevels), SEL_NT, 0, ss>>>(g, pyr, rowcnt, rowent, work, selpos, fincnt, status);
            ++c->launches;
        }
        else {
            // (the warp going straight on to Harris and the second retainBest -- one launch instead of three -- was measured: 0.1435 ms
            //  against 0.1159 ms per 256 VGA frames; the 14 serial Harris evaluations per lane cost the level-0 warps more than the
            //  two launch gaps and the Harris kernel's tail)
            static const int env_fused = getenv("ORBX_SELECT_FUSED") ? atoi(getenv("ORBX_SELECT_FUSED")) : 0;
            if (env_fused) {
                k_select_fast<true><<<dim3(B, (unsigned)g.nlevels), 32, 0, ss>>>(g, pyr, rowcnt, rowent, work, selpos, selcnt, fincnt);
                ++c->launches;
            } else {
                k_select_fast<false><<<dim3(B, (unsigned)g.nlevels), 32, 0, ss>>>(g, pyr, rowcnt, rowent, work, selpos, selcnt, fincnt);
                if (g.total_hblk > 0) k_harris<<<dim3((unsigned)g.total_hblk, B), HARRIS_NT, 0, ss>>>(g, pyr, work, selcnt);
                k_select_harris<<<dim3(B, (unsigned)g.nlevels), 32, (size_t)g.selh_elems * 12, ss>>>(g, work, selpos, selcnt, fincnt, g.selh_elems);
                c->launches += 3;
            }
        }
    }
    if (overlap_sel) CU(cudaEventRecord(side_join, side));
    nv.next("orbx:blur");
    if (marks) stage_mark(c, 4);
    if (g.total_blur > 0 && !blur_done) {
        k_blur<<<dim3((unsigned)g.total_blur, B), BLUR_NT, 0, st>>>(g, pyr, blur);
        ++c->launches;
    }
    if (overlap_sel) CU(cudaStreamWaitEvent(st, side_join, 0));
    nv.next("orbx:describe");
    if (marks) stage_mark(c, 5);
    static const int old_desc = getenv("ORBX_DESC_OLD") ? atoi(getenv("ORBX_DESC_OLD")) : 0;   // A/B timing only
    if (old_desc)
        k_describe<<<dim3((unsigned)((std::max(cap, 1) + DESC_KPB * DESC_KPW - 1) / (DESC_KPB * DESC_KPW)), B), DESC_NT, 0, st>>>(
            g, pyr, blur, work, fincnt, (const float4*)c->pattern.p, d_kps, d_desc, d_counts, cap);
    else {
        // groups of 4 slots per warp: four per warp when the batch is large (the warp's set-up is shared, and there are still
        // enough CTAs for several waves), one group per warp (lowest latency) when it is a single frame
        static const int env_gpw = getenv("ORBX_DESC_GPW") ? atoi(getenv("ORBX_DESC_GPW")) : 0;
        const long groups = (long)nb * ((std::min(std::max(cap, 1), std::max(c->nfeatures, 1)) + DS_G - 1) / DS_G);
        const int gpw = env_gpw > 0 ? std::min(env_gpw, DS_MAX_GPW) : (groups >= 148L * 24 * 8 ? 4 : 1);   // measured at 256 VGA frames: 1 / 2 / 4 / 8 / 16 -> 0.192 / 0.192 / 0.185 / 0.205 / 0.259 ms
        const int gcap = (std::max(cap, 1) + DS_G - 1) / DS_G;
        k_describe_tma<<<dim3((unsigned)((gcap + DS_NW * gpw - 1) / (DS_NW * gpw)), B), DS_NW * 32, DS_NW * DS_WARP_BYTES + 128, st>>>(
            g, c->dmaps, f0, work, fincnt, (const float4*)c->pattern.p, d_kps, d_desc, d_counts, cap, gpw, status);
    }
    ++c->launches;
    if (marks) stage_mark(c, 6);
    return ORBX_OK;
}

// The extraction pipeline on device-resident frames; everything asynchronous, ordered on c->stream.
// Large batches are cut into `lanes` frame ranges that run the same kernel sequence on their own streams: the stages
// bound differently (FAST issue-bound, pyramid / selection / describe latency-bound), so CTAs of different stages
// from different lanes share the SMs and fill each other's stalls.  Frames are independent, so no lane ever waits
// for another; the lanes fork from and join back into c->stream with events.
int run_extract(orbx_ctx* c, const uint8_t* d_imgs, int batch, int w, int h, size_t step, size_t frame_stride, int channels,
                float* d_kps, uint8_t* d_desc, int cap, int* d_counts)
{
    int rc = set_geometry(c, w, h);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->status.p, 0, sizeof(int) * (size_t)batch, c->stream));
    static const int env_lanes = getenv("ORBX_LANES") ? atoi(getenv("ORBX_LANES")) : 1;   // 1: every kernel fills the GPU on its own now (3 lanes measured 4 % slower)
    int lanes = std::max(1, std::min(env_lanes, MAX_LANES));
    if (batch < lanes * LANE_MIN_FRAMES) lanes = 1;
    if (c->profiling) lanes = 1;
    if (lanes == 1) {
        rc = run_extract_range(c, c->stream, 0, c->profiling, 0, batch, d_imgs, step,
                               frame_stride, channels, d_kps, d_desc, cap, d_counts);
        if (rc) return rc;
    } else {
        CU(cudaEventRecord(c->ev_fork, c->stream));
        for (int k = 0; k < lanes; ++k) {
            const int f0 = (int)((long)batch * k / lanes), f1 = (int)((long)batch * (k + 1) / lanes);
            cudaStream_t st = k == 0 ? c->stream : c->lane[k];
            if (k > 0) CU(cudaStreamWaitEvent(st, c->ev_fork, 0));
            rc = run_extract_range(c, st, k, false, f0, f1 - f0, d_imgs, step, frame_stride, channels, d_kps, d_desc, cap, d_counts);
            if (rc) return rc;
            if (k > 0) { CU(cudaEventRecord(c->ev_join[k], st)); CU(cudaStreamWaitEvent(c->stream, c->ev_join[k], 0)); }
        }
    }
    CU(cudaGetLastError());
    c->last_batch = batch;
    if (c->profiling) for (int i = 0; i < 6; ++i) c->stage_valid[i] = true;
    return ORBX_OK;
}

// Host frames [f0, f1) -> the context's staging buffer on stream `st`.  Tightly packed rows go as plain 1-D copies, and
// frames that are also contiguous in host memory (a video buffer, a numpy batch) as ONE copy per range: large linear
// transfers are what PCIe moves fastest.
int upload_frames(orbx_ctx* c, cudaStream_t st, const uint8_t* const* imgs, int f0, int f1, size_t step, size_t row, int h, size_t dstep, size_t fstride)
{
    if (step == row && dstep == row) {
        int i = f0;
        while (i < f1) {
            int j = i + 1;
            while (j < f1 && imgs[j] == imgs[j - 1] + fstride) ++j;
            CU(cudaMemcpyAsync((uint8_t*)c->in.p + fstride * i, imgs[i], fstride * (size_t)(j - i), cudaMemcpyHostToDevice, st));
            i = j;
        }
        return ORBX_OK;
    }
    for (int i = f0; i < f1; ++i)
        CU(cudaMemcpy2DAsync((uint8_t*)c->in.p + fstride * i, dstep, imgs[i], step, row, (size_t)h, cudaMemcpyHostToDevice, st));
    return ORBX_OK;
}

int check_args_extract(orbx_ctx* c, int batch, int w, int h, int channels, int cap)
{
    if (!c) return ORBX_E_ARG;
    if (batch < 0 || batch > c->max_batch) return fail(c, ORBX_E_ARG, "batch outside [0, max_batch]");
    if (w < 0 || h < 0 || w > c->max_w || h > c->max_h) return fail(c, ORBX_E_ARG, "frame larger than the context maximum");
    if (channels != 1 && channels != 3) return fail(c, ORBX_E_UNSUPPORTED, "channels must be 1 (gray) or 3 (BGR)");
    if (cap < 0) return fail(c, ORBX_E_ARG, "negative capacity");
    return ORBX_OK;
}

// A matcher status word was found set (an mbarrier wait timed out on the device): clear all of them, so that the next
// call starts clean without a per-call memset on the matcher's critical path, and report.
int match_timed_out(orbx_ctx* c)
{
    cudaMemset(c->mstatus.p, 0, sizeof(int) * 16);
    return fail(c, ORBX_E_INTERNAL, "matcher pipeline timed out on an mbarrier (device status set)");
}

// The matcher on device-resident operands, asynchronous on `st` (default: the context's stream).  `slot` selects the
// status word (one per lane, so concurrent lanes never reset each other's); lanes never take the split-train path,
// whose key scratch is shared.
int run_match(orbx_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int stride_rows, const int* d_counts, int nsets,
              int4* d_best, int4* d_second, cudaStream_t st = nullptr, int slot = 0)
{
    NvtxRange nvr("orbx:match_hamming");
    const bool lane_call = st != nullptr;
    if (!st) st = c->stream;
    if (nt >= MT_MAX_TRAIN) return fail(c, ORBX_E_UNSUPPORTED, "train set larger than 2^20 - 1 rows");
    const bool knn2 = d_second != nullptr;
    // CTA-pair kernel (tcgen05 cta_group::2, query tiles in twos) whenever the (query tile, set) pairs alone fill the GPU;
    // small problems keep the single-CTA kernel, whose CTAs are independent units for the train-row split.
    // ORBX_MATCH_2CTA = 0 / 1 forces one of them (experiments).
    static const int pair_env = getenv("ORBX_MATCH_2CTA") ? atoi(getenv("ORBX_MATCH_2CTA")) : -1;
    const int tiles_1 = (nq + MT_QROWS - 1) / MT_QROWS, tiles_2 = (tiles_1 + 1) & ~1;
    const bool pair = pair_env >= 0 ? pair_env != 0 : (tiles_1 >= 2 && (long)tiles_2 * nsets >= 148);
    const int tiles_m = pair ? tiles_2 : tiles_1;
    const int ntile_n = (nt + MT_BN - 1) / MT_BN;
    constexpr int kSMs = 148;                                // one CTA per SM (163 KB of shared memory each)
    // Few (query tile, set) pairs: split the train rows across CTAs and merge with atomicMax on the packed key.
    int nsplit = 1;
    if (!knn2 && !lane_call) {
        const long base = (long)tiles_m * nsets;
        if (base < kSMs) nsplit = (int)std::min<long>(ntile_n, std::max<long>(1, (kSMs + base - 1) / base));
    }
    const int tiles_per_split = (ntile_n + nsplit - 1) / nsplit;
    const int rows_per_split = tiles_per_split * MT_BN;
    nsplit = (nt + rows_per_split - 1) / rows_per_split;
    // Many sets: each CTA keeps its expanded query tile and walks sets z, z + zgroups, ... (persistent pipelines).
    // zgroups: waves(z) * sets-per-CTA(z) is the time in units of one (query tile, set) pass; a CTA's setup (TMEM
    // allocation, query expansion, pipeline fill) costs about a third of such a pass for ~1000-row sets, less for longer ones
    int zgroups = 1;
    {
        const long per_z = (long)tiles_m * nsplit;
        const long setup20 = std::max<long>(1, std::min<long>(20, 20L * 4 / std::max(1, tiles_per_split)));   // in 1/20 passes
        long best_cost = -1;
        for (int z = 1; z <= nsets; ++z) {
            const long waves = (per_z * z + kSMs - 1) / kSMs;
            const long cost = waves * (20L * ((nsets + z - 1) / z) + setup20);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; zgroups = z; }
        }
    }
    int* keys = nullptr;
    const size_t nout = (size_t)nq * nsets;
    int rc;
    if ((rc = ensure(c, c->mstatus, sizeof(int) * 16))) return rc;
    int* d_status = (int*)c->mstatus.p + slot;               // zero since orbx_create; re-zeroed by match_timed_out() after a failure
    if (!lane_call) stage_mark(c, 7);
    if (nsplit > 1) {
        if ((rc = ensure(c, c->mkeys, nout * sizeof(int)))) return rc;
        keys = (int*)c->mkeys.p;
        k_match_keys_init<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(keys, nout);
        ++c->launches;
    }
    const dim3 grd((unsigned)tiles_m, (unsigned)nsplit, (unsigned)zgroups);
    if (getenv("ORBX_MATCH_TRACE") && !c->mtrace.p) { if ((rc = ensure(c, c->mtrace, 16 * 16 * 8))) return rc; cudaMemset(c->mtrace.p, 0, 16 * 16 * 8); }
    static const int dbg = getenv("ORBX_MATCH_DBG") ? atoi(getenv("ORBX_MATCH_DBG")) : 0;   // perf experiments only: skips a role's work (wrong results)
    if (pair && knn2) k_hamming_umma2<true><<<grd, MT_THREADS, MT_SMEM_BYTES, st>>>(d_q, nq, d_t, nt, stride_rows, d_counts, nsets, rows_per_split, d_best, d_second, keys, d_status, dbg, (long long*)c->mtrace.p);
    else if (pair) k_hamming_umma2<false><<<grd, MT_THREADS, MT_SMEM_BYTES, st>>>(d_q, nq, d_t, nt, stride_rows, d_counts, nsets, rows_per_split, d_best, d_second, keys, d_status, dbg, (long long*)c->mtrace.p);
    else if (knn2) k_hamming_umma<true><<<grd, MT_THREADS, MT_SMEM_BYTES, st>>>(d_q, nq, d_t, nt, stride_rows, d_counts, nsets, rows_per_split, d_best, d_second, keys, d_status, dbg, (long long*)c->mtrace.p);
    else      k_hamming_umma<false><<<grd, MT_THREADS, MT_SMEM_BYTES, st>>>(d_q, nq, d_t, nt, stride_rows, d_counts, nsets, rows_per_split, d_best, d_second, keys, d_status, dbg, (long long*)c->mtrace.p);
    ++c->launches;
    if (nsplit > 1) {
        k_match_finalize<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(keys, nq, nout, d_best);
        ++c->launches;
    }
    if (!lane_call) { stage_mark(c, 8); if (c->profiling) c->stage_valid[6] = true; }
    CU(cudaGetLastError());
    return ORBX_OK;
}

int match_host(orbx_ctx* c, const uint8_t* query, int nq, const uint8_t* train, int nt, orbx_match* out, int* n_out, bool knn2)
{
    if (!c) return ORBX_E_ARG;
    if (n_out) *n_out = 0;
    if (nq < 0 || nt < 0) return fail(c, ORBX_E_ARG, "negative row count");
    if (nq == 0 || nt == 0) return ORBX_OK;                 // cv: empty query or train -> no matches, no throw
    if (!query || !train || !out) return fail(c, ORBX_E_ARG, "null pointer");
    CU(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(c, c->mq, (size_t)nq * 32))) return rc;
    if ((rc = ensure(c, c->mt, (size_t)nt * 32))) return rc;
    if ((rc = ensure(c, c->mbest, (size_t)nq * 16))) return rc;
    if (knn2 && (rc = ensure(c, c->msecond, (size_t)nq * 16))) return rc;
    CU(cudaMemcpyAsync(c->mq.p, query, (size_t)nq * 32, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->mt.p, train, (size_t)nt * 32, cudaMemcpyHostToDevice, c->stream));
    if ((rc = run_match(c, (const uint8_t*)c->mq.p, nq, (const uint8_t*)c->mt.p, nt, nt, nullptr, 1, (int4*)c->mbest.p, knn2 ? (int4*)c->msecond.p : nullptr))) return rc;
    if (!knn2) {
        CU(cudaMemcpyAsync(out, c->mbest.p, (size_t)nq * 16, cudaMemcpyDeviceToHost, c->stream));
    } else {
        CU(cudaMemcpy2DAsync(out, 32, c->mbest.p, 16, 16, (size_t)nq, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpy2DAsync(out + 1, 32, c->msecond.p, 16, 16, (size_t)nq, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(c->h_small, c->mstatus.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->h_small[0]) return match_timed_out(c);
    if (n_out) *n_out = nq;
    return ORBX_OK;
}

}  // namespace

// =================================================================================================== C-ABI
extern "C" {

const char* orbx_version(void) { return "orbx 0.1 (sm_100a; tcgen05 int8 matcher)"; }

int orbx_create(orbx_ctx** out, int device, int nfeatures, float scale_factor, int nlevels, int max_w, int max_h, int max_batch)
{
    if (!out) return ORBX_E_ARG;
    *out = nullptr;
    if (nfeatures < 0 || nlevels < 1 || nlevels > ORBX_LEVELS_MAX || !(scale_factor > 1.0f) || max_w < 1 || max_h < 1 ||
        max_w > 65535 || max_h > 65535 || max_batch < 1 || max_batch > 65535)
        return ORBX_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return ORBX_E_CUDA;
    orbx_ctx* c = new (std::nothrow) orbx_ctx();
    if (!c) return ORBX_E_NOMEM;
    c->device = device; c->nfeatures = nfeatures; c->scale_factor = scale_factor; c->nlevels = nlevels;
    c->max_w = max_w; c->max_h = max_h; c->max_batch = max_batch;
    auto bail = [&](int code) { orbx_destroy(c); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(ORBX_E_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(ORBX_E_CUDA);
    if (prop.major != 10) return bail(ORBX_E_CUDA);          // sm_100a only: no other code path exists
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(ORBX_E_CUDA);
    for (int k = 0; k < HOST_MAX_LANES; ++k)
        if ((k > 0 && cudaStreamCreateWithFlags(&c->lane[k], cudaStreamNonBlocking) != cudaSuccess) || cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming) != cudaSuccess) return bail(ORBX_E_CUDA);
    if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess) return bail(ORBX_E_CUDA);
    {
        int prio_lo = 0, prio_hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) return bail(ORBX_E_CUDA);
        for (int k = 0; k < HOST_MAX_LANES; ++k)
            if (cudaStreamCreateWithPriority(&c->side[k], cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
                cudaEventCreateWithFlags(&c->ev_sfork[k], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&c->ev_sjoin[k], cudaEventDisableTiming) != cudaSuccess) return bail(ORBX_E_CUDA);
    }
    for (int i = 0; i < N_STAGES + 2; ++i) if (cudaEventCreate(&c->ev[i]) != cudaSuccess) return bail(ORBX_E_CUDA);
    build_geom(c, max_w, max_h, &c->geom_max, nullptr);
    const Geom& g = c->geom_max;
    const size_t B = (size_t)max_batch;
    if (ensure(c, c->pyr, g.pyr_frame * B + (size_t)(96 + 8) * g.L[0].pitch) ||   // + slack: the blur streams up to ~100 rows past a level's end (masked outputs)
        ensure(c, c->blur, g.pyr_frame * B) || ensure(c, c->rowcnt, g.cnt_frame * 4 * B) || ensure(c, c->rowent, g.ent_frame * 4 * B) ||
        ensure(c, c->work, g.ws_frame * sizeof(Elem) * B) || ensure(c, c->selpos, g.ws_frame * 8 * B) || ensure(c, c->fincnt, sizeof(int) * ORBX_LEVELS_MAX * B) || ensure(c, c->selcnt, sizeof(int) * ORBX_LEVELS_MAX * B) ||
        ensure(c, c->status, sizeof(int) * B) || ensure(c, c->pattern, sizeof(float) * 1024) || ensure(c, c->mstatus, sizeof(int) * 16))
        return bail(ORBX_E_NOMEM);
    if (cudaMemset(c->mstatus.p, 0, sizeof(int) * 16) != cudaSuccess) return bail(ORBX_E_CUDA);
    {
        float pat[1024];                                     // the rBRIEF pattern as float4 (x0, y0, x1, y1) per test,
        for (int t = 0; t < 8; ++t)                          // transposed to [t][lane]: test lane * 8 + t  (k_describe)
            for (int lane = 0; lane < 32; ++lane)
                for (int k = 0; k < 4; ++k) pat[(t * 32 + lane) * 4 + k] = (float)k_pattern_host[(lane * 8 + t) * 4 + k];
        if (cudaMemcpy(c->pattern.p, pat, sizeof pat, cudaMemcpyHostToDevice) != cudaSuccess) return bail(ORBX_E_CUDA);
    }
    if (cudaMemset(c->status.p, 0, sizeof(int) * B) != cudaSuccess) return bail(ORBX_E_CUDA);
    if (cudaMallocHost((void**)&c->h_small, sizeof(int) * (2 * B + 16)) != cudaSuccess) return bail(ORBX_E_NOMEM);
    if (cudaFuncSetAttribute(k_hamming_umma2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_hamming_umma2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_hamming_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_hamming_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_fast_warp<FAST_NWARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_NWARP * FW_WARP_BYTES + 128) != cudaSuccess ||
        cudaFuncSetAttribute(k_pyr_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_describe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_NW * DS_WARP_BYTES + 128) != cudaSuccess ||
        cudaFuncSetAttribute(k_select_harris, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 12) != cudaSuccess)
        return bail(ORBX_E_CUDA);
    *out = c;
    return ORBX_OK;
}

void orbx_destroy(orbx_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    Buf* bufs[] = {&c->pyr, &c->blur, &c->rowcnt, &c->rowent, &c->work, &c->selpos, &c->fincnt, &c->selcnt, &c->status, &c->tabs, &c->pattern, &c->in, &c->kps,
                   &c->desc, &c->counts, &c->mq, &c->mt, &c->mbest, &c->msecond, &c->mkeys, &c->mstatus, &c->mcounts, &c->mtrace,
                   &c->t_desc, &c->t_pos, &c->t_nrm, &c->t_outl, &c->t_slots, &c->t_cand, &c->t_ncand, &c->t_q, &c->t_in, &c->t_best, &c->t_filtered,
                   &c->t_minmax, &c->t_aux};
    for (Buf* b : bufs) if (b->p) cudaFree(b->p);
    if (c->h_small) cudaFreeHost(c->h_small);
    delete c->stager;
    if (c->h_stage_in) cudaFreeHost(c->h_stage_in);
    if (c->h_stage_out) cudaFreeHost(c->h_stage_out);
    for (auto& a : c->aslot) {
        if (a.h_in) cudaFreeHost(a.h_in);
        if (a.h_out) cudaFreeHost(a.h_out);
        Buf* ab[] = {&a.d_kps, &a.d_desc, &a.d_counts};
        for (Buf* b : ab) if (b->p) cudaFree(b->p);
        if (a.done) cudaEventDestroy(a.done);
    }
    for (int i = 0; i < N_STAGES + 2; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int k = 0; k < HOST_MAX_LANES; ++k) { if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]); if (c->lane[k]) { cudaStreamSynchronize(c->lane[k]); cudaStreamDestroy(c->lane[k]); } }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int k = 0; k < HOST_MAX_LANES; ++k) {
        if (c->side[k]) { cudaStreamSynchronize(c->side[k]); cudaStreamDestroy(c->side[k]); }
        if (c->ev_sfork[k]) cudaEventDestroy(c->ev_sfork[k]);
        if (c->ev_sjoin[k]) cudaEventDestroy(c->ev_sjoin[k]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* orbx_last_error(const orbx_ctx* c) { return c ? c->err.c_str() : "null context"; }
void* orbx_stream(orbx_ctx* c) { return c ? (void*)c->stream : nullptr; }
uint64_t orbx_launch_count(const orbx_ctx* c) { return c ? c->launches : 0; }

int orbx_synchronize(orbx_ctx* c)
{
    if (!c) return ORBX_E_ARG;
    CU(cudaSetDevice(c->device));
    const int b = c->last_batch;
    if (b > 0) CU(cudaMemcpyAsync(c->h_small, c->status.p, sizeof(int) * (size_t)b, cudaMemcpyDeviceToHost, c->stream));
    int* h_m = c->h_small + 2 * (size_t)c->max_batch;        // the asynchronous matcher entry points defer their status words to this call
    CU(cudaMemcpyAsync(h_m, c->mstatus.p, sizeof(int) * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 16; ++i) if (h_m[i]) return match_timed_out(c);
    for (int i = 0; i < b; ++i)
        if (c->h_small[i]) return fail(c, ORBX_E_INTERNAL, "device-side status set for a frame");
    return ORBX_OK;
}

int orbx_detect_and_compute_device(orbx_ctx* c, const uint8_t* d_imgs, int batch, int w, int h, size_t step, size_t frame_stride,
                                   int channels, orbx_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_counts)
{
    int rc = check_args_extract(c, batch, w, h, channels, cap);
    if (rc) return rc;
    if (batch == 0) return ORBX_OK;
    if (!d_imgs || !d_kps || !d_desc || !d_counts) return fail(c, ORBX_E_ARG, "null pointer");
    if (w == 0 || h == 0) return fail(c, ORBX_E_ARG, "empty frames are only accepted by the host entry points");
    if (step < (size_t)w * channels) return fail(c, ORBX_E_ARG, "step smaller than a row");
    CU(cudaSetDevice(c->device));
    return run_extract(c, d_imgs, batch, w, h, step, frame_stride, channels, (float*)d_kps, d_desc, cap, d_counts);
}

// ---- host-buffer batches: one core for orbx_detect_and_compute_batch (nmaps = 0) and orbx_extract_match_batch.
// Frame ranges ("lanes") on their own streams: upload -> kernels -> (matches) -> download per lane, so the H2D copy of lane
// k+1 runs under the kernels of lane k and the D2H of lane k under the kernels of lane k+1 (PCIe is the bound here).
// Caller buffers that are NOT page-locked (the reference's cv::Mat frames and std::vector results) go through the context's
// pinned staging, copied by the HostStager threads (orbx_stage.h), so the lanes overlap for them as well.
static bool host_ptr_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int ensure_pinned(orbx_ctx* c, uint8_t*& p, size_t& have, size_t bytes)
{
    if (have >= bytes && p) return ORBX_OK;
    if (p) { cudaFreeHost(p); p = nullptr; have = 0; }
    bytes = std::max<size_t>(bytes, 4096);
    if (cudaMallocHost((void**)&p, bytes) != cudaSuccess) { cudaGetLastError(); return fail(c, ORBX_E_NOMEM, "cudaMallocHost failed for the pinned staging buffer"); }
    have = bytes;
    return ORBX_OK;
}

static HostStager* stager_of(orbx_ctx* c)
{
    if (!c->stager) {
        int n = getenv("ORBX_STAGE_THREADS") ? atoi(getenv("ORBX_STAGE_THREADS")) : (int)std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
        c->stager = new HostStager(std::max(0, std::min(n, 32)));
    }
    return c->stager;
}

static int host_batch(orbx_ctx* c, const uint8_t* const* imgs, int batch, int w, int h, size_t step, int channels,
                      orbx_keypoint* kps, uint8_t* desc, int cap, int* n_out, const uint8_t* const* queries, const int* nq,
                      int nmaps, orbx_match* const* best)
{
    int rc;
    NvtxRange nvr("orbx:host_batch (stage / upload / extract / match / download)");
    CU(cudaSetDevice(c->device));
    const size_t row = (size_t)w * channels, dstep = round_up(row, 16), fstride = dstep * h;   // 16: k_gray takes 128-bit loads
    const size_t capz = (size_t)std::max(cap, 1);
    size_t qrows = 0, mrows = 0;
    for (int j = 0; j < nmaps; ++j) { qrows += (size_t)nq[j]; mrows += (size_t)nq[j] * batch; }
    if ((rc = ensure(c, c->in, fstride * batch)) || (rc = ensure(c, c->kps, sizeof(orbx_keypoint) * capz * batch)) ||
        (rc = ensure(c, c->desc, (size_t)32 * capz * batch)) || (rc = ensure(c, c->counts, sizeof(int) * (size_t)batch)))
        return rc;
    if (nmaps > 0 && ((rc = ensure(c, c->mq, std::max<size_t>(qrows, 1) * 32)) || (rc = ensure(c, c->mbest, std::max<size_t>(mrows, 1) * 16)) ||
                      (rc = ensure(c, c->mstatus, sizeof(int) * 16))))
        return rc;
    for (int i = 0; i < batch; ++i) if (!imgs[i]) return fail(c, ORBX_E_ARG, "null frame pointer");
    if ((rc = set_geometry(c, w, h))) return rc;

    // ---- which caller buffers need staging
    static const int stage_env = getenv("ORBX_STAGE") ? atoi(getenv("ORBX_STAGE")) : 1;      // 0: never stage (A/B timing only)
    const bool stage_in = stage_env && !(host_ptr_is_pinned(imgs[0]) && host_ptr_is_pinned(imgs[batch - 1] + (size_t)(h - 1) * step + row - 1));
    bool stage_out = false;
    if (stage_env && cap > 0) stage_out = !(host_ptr_is_pinned(kps) && host_ptr_is_pinned(desc));
    for (int j = 0; j < nmaps && stage_env && !stage_out; ++j) if (nq[j] > 0 && !host_ptr_is_pinned(best[j])) stage_out = true;
    const size_t o_kps = 0, o_desc = round_up(sizeof(orbx_keypoint) * capz * batch, 256), o_best = o_desc + round_up((size_t)32 * capz * batch, 256);
    if (stage_in && (rc = ensure_pinned(c, c->h_stage_in, c->h_stage_in_bytes, fstride * batch))) return rc;
    if (stage_out && (rc = ensure_pinned(c, c->h_stage_out, c->h_stage_out_bytes, o_best + mrows * 16))) return rc;
    HostStager* hs = (stage_in || stage_out) ? stager_of(c) : nullptr;
    static const int stage_nt = getenv("ORBX_STAGE_NT") ? atoi(getenv("ORBX_STAGE_NT")) : 1;   // non-temporal stores into the pinned input staging
    static const size_t wake_bytes = getenv("ORBX_STAGE_WAKE_BYTES") ? (size_t)atoll(getenv("ORBX_STAGE_WAKE_BYTES")) : (size_t)2 << 20;
    const bool wake_in = fstride * batch >= wake_bytes, wake_out = (size_t)60 * capz * batch + mrows * 16 >= wake_bytes;
    orbx_keypoint* kps_dl = stage_out ? (orbx_keypoint*)(c->h_stage_out + o_kps) : kps;       // where the D2H copies land
    uint8_t* desc_dl = stage_out ? c->h_stage_out + o_desc : desc;

    CU(cudaMemsetAsync(c->status.p, 0, sizeof(int) * (size_t)batch, c->stream));
    {
        size_t off = 0;
        for (int j = 0; j < nmaps; ++j) {
            if (nq[j] > 0) CU(cudaMemcpyAsync((uint8_t*)c->mq.p + off * 32, queries[j], (size_t)nq[j] * 32, cudaMemcpyHostToDevice, c->stream));
            off += (size_t)nq[j];
        }
    }
    static const int host_lanes_max = std::max(1, std::min(HOST_MAX_LANES, getenv("ORBX_HOST_LANES") ? atoi(getenv("ORBX_HOST_LANES")) : HOST_DEFAULT_LANES));
    int* h_counts = c->h_small;
    int* h_status = c->h_small + batch;
    int* h_mstatus = c->h_small + 2 * batch;
    const int lanes = c->profiling ? 1 : std::max(1, std::min(host_lanes_max, batch / HOST_LANE_MIN_FRAMES));
    // staged inputs: all pieces are queued up front, lane by lane, so the copier threads run ahead of the lane loop
    std::atomic<int> in_done[HOST_MAX_LANES], out_done(0);
    int in_target[HOST_MAX_LANES] = {}, out_target = 0;
    struct Drain {                                           // no queued piece may outlive the counters above, whichever way this call ends
        HostStager* hs; std::atomic<int>* in; int* in_t; int n; std::atomic<int>* out; int* out_t;
        ~Drain() { if (hs) { for (int k = 0; k < n; ++k) hs->help_until(in[k], in_t[k]); hs->help_until(*out, *out_t); } }
    } drain{hs, in_done, in_target, stage_in ? lanes : 0, &out_done, &out_target};
    if (stage_in)
        for (int k = 0; k < lanes; ++k) {
            int f0, f1;
            host_lane_range(batch, lanes, k, &f0, &f1);
            in_done[k].store(0, std::memory_order_relaxed);
            for (int i = f0; i < f1; ++i) in_target[k] += hs->submit(c->h_stage_in + fstride * i, dstep, imgs[i], step, row, (size_t)h, &in_done[k], wake_in, stage_nt != 0);
        }
    CU(cudaEventRecord(c->ev_fork, c->stream));
    for (int k = 0; k < lanes; ++k) {
        int f0, f1;
        host_lane_range(batch, lanes, k, &f0, &f1);
        const size_t n = (size_t)(f1 - f0);
        cudaStream_t st = k == 0 ? c->stream : c->lane[k];
        if (k > 0) CU(cudaStreamWaitEvent(st, c->ev_fork, 0));
        if (stage_in) {
            hs->help_until(in_done[k], in_target[k]);        // (the pinned image has the device layout: one copy per lane)
            CU(cudaMemcpyAsync((uint8_t*)c->in.p + fstride * f0, c->h_stage_in + fstride * f0, fstride * n, cudaMemcpyHostToDevice, st));
        } else if ((rc = upload_frames(c, st, imgs, f0, f1, step, row, h, dstep, fstride))) return rc;
        if ((rc = run_extract_range(c, st, k, c->profiling, f0, f1 - f0, (const uint8_t*)c->in.p, dstep,
                                    fstride, channels, (float*)c->kps.p, (uint8_t*)c->desc.p, cap, (int*)c->counts.p)))
            return rc;
        CU(cudaMemcpyAsync(h_counts + f0, (int*)c->counts.p + f0, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_status + f0, (int*)c->status.p + f0, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
        if (cap > 0) {
            // outputs are [batch][cap] on both sides: bulk copies per lane (records past n_out[i] are unspecified)
            CU(cudaMemcpyAsync(kps_dl + (size_t)f0 * cap, (orbx_keypoint*)c->kps.p + (size_t)f0 * cap, sizeof(orbx_keypoint) * (size_t)cap * n, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(desc_dl + (size_t)f0 * cap * 32, (uint8_t*)c->desc.p + (size_t)f0 * cap * 32, (size_t)32 * cap * n, cudaMemcpyDeviceToHost, st));
        }
        size_t qoff = 0, moff = 0;
        for (int j = 0; j < nmaps; ++j) {
            if (nq[j] > 0) {
                int4* d_best = (int4*)c->mbest.p + moff + (size_t)f0 * nq[j];
                // train sets = this lane's frames, straight from the extraction outputs ([frame][cap][32] + counts)
                if ((rc = run_match(c, (const uint8_t*)c->mq.p + qoff * 32, nq[j], (const uint8_t*)c->desc.p + (size_t)f0 * cap * 32, cap, cap,
                                    (const int*)c->counts.p + f0, (int)n, d_best, nullptr, st, k)))
                    return rc;
                orbx_match* dl = stage_out ? (orbx_match*)(c->h_stage_out + o_best) + moff + (size_t)f0 * nq[j] : best[j] + (size_t)f0 * nq[j];
                CU(cudaMemcpyAsync(dl, d_best, (size_t)16 * nq[j] * n, cudaMemcpyDeviceToHost, st));
            }
            qoff += (size_t)nq[j]; moff += (size_t)nq[j] * batch;
        }
        if (nmaps > 0) CU(cudaMemcpyAsync(h_mstatus + k, (int*)c->mstatus.p + k, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (k > 0 || stage_out) CU(cudaEventRecord(c->ev_join[k], st));
        if (k > 0) CU(cudaStreamWaitEvent(c->stream, c->ev_join[k], 0));
    }
    CU(cudaGetLastError());
    c->last_batch = batch;
    c->out_cap = cap;
    c->view_desc = (const uint8_t*)c->desc.p; c->view_counts = (const int*)c->counts.p; c->view_frames = batch;
    if (c->profiling) for (int i = 0; i < 6; ++i) c->stage_valid[i] = true;
    if (stage_out) {
        // results leave the pinned staging lane by lane, as soon as a lane's downloads have landed, under the later lanes' work:
        // only the records a frame really has are copied (the caller's records past n_out[i] stay untouched)
        for (int k = 0; k < lanes; ++k) {
            int f0, f1;
            host_lane_range(batch, lanes, k, &f0, &f1);
            CU(cudaEventSynchronize(c->ev_join[k]));
            for (int i = f0; i < f1; ++i) {
                const size_t ni = (size_t)std::max(0, std::min(h_counts[i], cap));
                if (ni > 0) {
                    out_target += hs->submit(kps + (size_t)i * cap, 0, kps_dl + (size_t)i * cap, 0, sizeof(orbx_keypoint) * ni, 1, &out_done, wake_out);
                    out_target += hs->submit(desc + (size_t)i * cap * 32, 0, desc_dl + (size_t)i * cap * 32, 0, 32 * ni, 1, &out_done, wake_out);
                }
            }
            size_t moff = 0;
            for (int j = 0; j < nmaps; ++j) {
                if (nq[j] > 0)
                    out_target += hs->submit(best[j] + (size_t)f0 * nq[j], 0, (orbx_match*)(c->h_stage_out + o_best) + moff + (size_t)f0 * nq[j], 0,
                                             (size_t)16 * nq[j] * (size_t)(f1 - f0), 1, &out_done, wake_out);
                moff += (size_t)nq[j] * batch;
            }
        }
        hs->help_until(out_done, out_target);
    }
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < lanes; ++k) if (nmaps > 0 && h_mstatus[k]) return match_timed_out(c);
    bool over = false;
    for (int i = 0; i < batch; ++i) {
        if (h_status[i]) return fail(c, ORBX_E_INTERNAL, "device-side status set for a frame");
        n_out[i] = h_counts[i];
        if (h_counts[i] > cap) over = true;
    }
    if (over) return fail(c, ORBX_E_CAPACITY, "output capacity too small; n_out holds the needed counts");
    return ORBX_OK;
}

int orbx_detect_and_compute_batch(orbx_ctx* c, const uint8_t* const* imgs, int batch, int w, int h, size_t step, int channels,
                                  orbx_keypoint* kps, uint8_t* desc, int cap, int* n_out)
{
    int rc = check_args_extract(c, batch, w, h, channels, cap);
    if (rc) return rc;
    if (n_out) for (int i = 0; i < batch; ++i) n_out[i] = 0;
    if (batch == 0 || w == 0 || h == 0) return ORBX_OK;     // cv: empty image -> silent return, no keypoints
    if (!imgs || !n_out || (cap > 0 && (!kps || !desc))) return fail(c, ORBX_E_ARG, "null pointer");
    if (step < (size_t)w * channels) return fail(c, ORBX_E_ARG, "step smaller than a row");
    return host_batch(c, imgs, batch, w, h, step, channels, kps, desc, cap, n_out, nullptr, nullptr, 0, nullptr);
}

int orbx_extract_match_batch(orbx_ctx* c, const uint8_t* const* imgs, int batch, int w, int h, size_t step, int channels,
                             orbx_keypoint* kps, uint8_t* desc, int cap, int* n_out, const uint8_t* const* queries, const int* nq,
                             int nmaps, orbx_match* const* best)
{
    int rc = check_args_extract(c, batch, w, h, channels, cap);
    if (rc) return rc;
    if (nmaps < 0 || nmaps > 8) return fail(c, ORBX_E_ARG, "nmaps outside [0, 8]");
    if (n_out) for (int i = 0; i < batch; ++i) n_out[i] = 0;
    if (batch == 0) return ORBX_OK;
    if (!imgs || !n_out || cap <= 0 || !kps || !desc) return fail(c, ORBX_E_ARG, "null pointer / zero capacity");
    if (nmaps > 0 && (!queries || !nq || !best)) return fail(c, ORBX_E_ARG, "null map arguments");
    for (int j = 0; j < nmaps; ++j) if (nq[j] < 0 || (nq[j] > 0 && (!queries[j] || !best[j]))) return fail(c, ORBX_E_ARG, "bad map");
    if (w == 0 || h == 0) {                                  // cv: empty image -> no keypoints; a query against an empty train set -> no matches
        return ORBX_OK;
    }
    if (step < (size_t)w * channels) return fail(c, ORBX_E_ARG, "step smaller than a row");
    if (cap >= MT_MAX_TRAIN) return fail(c, ORBX_E_UNSUPPORTED, "capacity larger than 2^20 - 1 rows");
    return host_batch(c, imgs, batch, w, h, step, channels, kps, desc, cap, n_out, queries, nq, nmaps, best);
}

int orbx_host_register(orbx_ctx* c, void* ptr, size_t bytes)
{
    if (!c) return ORBX_E_ARG;
    if (!ptr || bytes == 0) return fail(c, ORBX_E_ARG, "null pointer / empty range");
    CU(cudaSetDevice(c->device));
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return ORBX_OK;
}

int orbx_host_unregister(orbx_ctx* c, void* ptr)
{
    if (!c) return ORBX_E_ARG;
    if (!ptr) return fail(c, ORBX_E_ARG, "null pointer");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));                    // nothing of this context may still be copying from / into it
    CU(cudaHostUnregister(ptr));
    return ORBX_OK;
}

int orbx_detect_and_compute(orbx_ctx* c, const uint8_t* img, int w, int h, size_t step, int channels, orbx_keypoint* kps,
                            uint8_t* desc, int cap, int* n_out)
{
    if (!c) return ORBX_E_ARG;
    if (!n_out) return fail(c, ORBX_E_ARG, "null n_out");
    *n_out = 0;
    if (w == 0 || h == 0 || !img) { if (!img && w > 0 && h > 0) return fail(c, ORBX_E_ARG, "null image"); return ORBX_OK; }
    const uint8_t* one[1] = {img};
    return orbx_detect_and_compute_batch(c, one, 1, w, h, step, channels, kps, desc, cap, n_out);
}

int orbx_match_hamming(orbx_ctx* c, const uint8_t* q, int nq, const uint8_t* t, int nt, orbx_match* out, int* n_out)
{
    return match_host(c, q, nq, t, nt, out, n_out, false);
}
int orbx_match_hamming_knn2(orbx_ctx* c, const uint8_t* q, int nq, const uint8_t* t, int nt, orbx_match* out, int* n_out)
{
    return match_host(c, q, nq, t, nt, out, n_out, true);
}

int orbx_match_hamming_device(orbx_ctx* c, const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, int nsets,
                              orbx_match* d_best, orbx_match* d_second)
{
    if (!c) return ORBX_E_ARG;
    if (nq < 0 || nt < 0 || nsets < 0 || nsets > 65535) return fail(c, ORBX_E_ARG, "bad sizes");
    if (nq == 0 || nt == 0 || nsets == 0) return ORBX_OK;
    if (!d_query || !d_train || !d_best) return fail(c, ORBX_E_ARG, "null pointer");
    CU(cudaSetDevice(c->device));
    return run_match(c, d_query, nq, d_train, nt, nt, nullptr, nsets, (int4*)d_best, (int4*)d_second);
}

int orbx_match_hamming_device_ragged(orbx_ctx* c, const uint8_t* d_query, int nq, const uint8_t* d_train, int train_stride_rows,
                                     const int* d_train_counts, int nsets, orbx_match* d_best, orbx_match* d_second)
{
    if (!c) return ORBX_E_ARG;
    if (nq < 0 || train_stride_rows < 0 || nsets < 0 || nsets > 65535) return fail(c, ORBX_E_ARG, "bad sizes");
    if (nq == 0 || train_stride_rows == 0 || nsets == 0) return ORBX_OK;
    if (!d_query || !d_train || !d_best || !d_train_counts) return fail(c, ORBX_E_ARG, "null pointer");
    CU(cudaSetDevice(c->device));
    return run_match(c, d_query, nq, d_train, train_stride_rows, train_stride_rows, d_train_counts, nsets, (int4*)d_best, (int4*)d_second);
}

int orbx_match_hamming_sets(orbx_ctx* c, const uint8_t* query, int nq, const uint8_t* train, int train_stride_rows,
                            const int* train_counts, int nsets, orbx_match* best, orbx_match* second)
{
    if (!c) return ORBX_E_ARG;
    if (nq < 0 || train_stride_rows < 0 || nsets < 0 || nsets > 65535) return fail(c, ORBX_E_ARG, "bad sizes");
    if (nq == 0 || train_stride_rows == 0 || nsets == 0) return ORBX_OK;
    if (!query || !train || !best || !train_counts) return fail(c, ORBX_E_ARG, "null pointer");
    CU(cudaSetDevice(c->device));
    const size_t tbytes = (size_t)nsets * train_stride_rows * 32, obytes = (size_t)nsets * nq * 16;
    int rc;
    if ((rc = ensure(c, c->mq, (size_t)nq * 32)) || (rc = ensure(c, c->mt, tbytes)) || (rc = ensure(c, c->mbest, obytes)) ||
        (rc = ensure(c, c->mcounts, sizeof(int) * (size_t)nsets)) || (second && (rc = ensure(c, c->msecond, obytes))))
        return rc;
    CU(cudaMemcpyAsync(c->mq.p, query, (size_t)nq * 32, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->mt.p, train, tbytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->mcounts.p, train_counts, sizeof(int) * (size_t)nsets, cudaMemcpyHostToDevice, c->stream));
    if ((rc = run_match(c, (const uint8_t*)c->mq.p, nq, (const uint8_t*)c->mt.p, train_stride_rows, train_stride_rows, (const int*)c->mcounts.p,
                        nsets, (int4*)c->mbest.p, second ? (int4*)c->msecond.p : nullptr)))
        return rc;
    CU(cudaMemcpyAsync(best, c->mbest.p, obytes, cudaMemcpyDeviceToHost, c->stream));
    if (second) CU(cudaMemcpyAsync(second, c->msecond.p, obytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_small, c->mstatus.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->h_small[0]) return match_timed_out(c);
    return ORBX_OK;
}

int orbx_filter_matches(orbx_match* m, int n, float ratio)
{
    if (!m || n <= 0) return 0;
    // records with trainIdx < 0 are the "empty train set" placeholders of the batched matchers (distance 0): they are neither
    // matches nor candidates for the minimum -- same rule as the device filter k_filter_matches
    bool any = false;
    float mn = 0.0f;
    for (int i = 0; i < n; ++i) if (m[i].trainIdx >= 0 && (!any || m[i].distance < mn)) { mn = m[i].distance; any = true; }
    if (!any) return 0;
    const float mx = std::max(mn * ratio, 30.0f);
    int k = 0;
    for (int i = 0; i < n; ++i) if (m[i].trainIdx >= 0 && m[i].distance <= mx) m[k++] = m[i];
    return k;
}

int orbx_level_geometry(const orbx_ctx* c, int w, int h, int* ws, int* hs, float* scales, int* quotas)
{
    if (!c || w < 0 || h < 0) return ORBX_E_ARG;
    int a[ORBX_LEVELS_MAX], b[ORBX_LEVELS_MAX], q[ORBX_LEVELS_MAX];
    float s[ORBX_LEVELS_MAX];
    level_sizes(w, h, c->nlevels, c->scale_factor, a, b, s);
    level_quotas(c->nfeatures, c->scale_factor, c->nlevels, q);
    for (int l = 0; l < c->nlevels; ++l) {
        if (ws) ws[l] = a[l];
        if (hs) hs[l] = b[l];
        if (scales) scales[l] = s[l];
        if (quotas) quotas[l] = q[l];
    }
    return ORBX_OK;
}

int orbx_debug_read_level(orbx_ctx* c, int frame, int level, uint8_t* out, size_t out_bytes)
{
    if (!c || !out) return ORBX_E_ARG;
    if (frame < 0 || frame >= c->last_batch || level < 0 || level >= c->geom.nlevels) return fail(c, ORBX_E_ARG, "frame/level out of range");
    const LevelGeom& L = c->geom.L[level];
    if (out_bytes < (size_t)L.w * L.h) return fail(c, ORBX_E_CAPACITY, "buffer too small");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (L.w > 0 && L.h > 0)
        CU(cudaMemcpy2D(out, (size_t)L.w, (const uint8_t*)c->pyr.p + (size_t)frame * c->geom.pyr_frame + L.img_off, (size_t)L.pitch,
                        (size_t)L.w, (size_t)L.h, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_debug_read_fast(orbx_ctx* c, int frame, int level, int32_t* x, int32_t* y, int32_t* score, int cap, int* n_out)
{
    if (!c || !n_out) return ORBX_E_ARG;
    *n_out = 0;
    if (frame < 0 || frame >= c->last_batch || level < 0 || level >= c->geom.nlevels) return fail(c, ORBX_E_ARG, "frame/level out of range");
    const Geom& g = c->geom;
    const LevelGeom& L = g.L[level];
    if (L.in_h <= 0) return ORBX_OK;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> cnt((size_t)L.in_h), ent((size_t)L.in_h * L.ent_pitch);
    CU(cudaMemcpy(cnt.data(), (const uint32_t*)c->rowcnt.p + (size_t)frame * g.cnt_frame + L.cnt_off, cnt.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ent.data(), (const uint32_t*)c->rowent.p + (size_t)frame * g.ent_frame + L.ent_off, ent.size() * 4, cudaMemcpyDeviceToHost));
    int n = 0;
    for (int r = 0; r < L.in_h; ++r)
        for (uint32_t i = 0; i < cnt[r]; ++i, ++n)
            if (n < cap) {
                const uint32_t e = ent[(size_t)r * L.ent_pitch + i];
                if (x) x[n] = (int)(e & 0xffffu);
                if (y) y[n] = r + ORBX_EDGE;
                if (score) score[n] = (int)(e >> 16);
            }
    *n_out = n;
    return n > cap ? ORBX_E_CAPACITY : ORBX_OK;
}

int orbx_debug_force_kernels(orbx_ctx* c, int mode)
{
    if (!c || mode < -1 || mode > 1) return ORBX_E_ARG;
    c->force_kernels = mode;
    return ORBX_OK;
}

int orbx_set_profiling(orbx_ctx* c, int enable)
{
    if (!c) return ORBX_E_ARG;
    c->profiling = enable != 0;
    for (int i = 0; i < N_STAGES; ++i) c->stage_valid[i] = false;
    return ORBX_OK;
}

int orbx_debug_match_trace(orbx_ctx* c, long long* out /*256*/)
{
    if (!c || !out || !c->mtrace.p) return ORBX_E_ARG;
    cudaStreamSynchronize(c->stream);
    return cudaMemcpy(out, c->mtrace.p, 16 * 16 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? ORBX_OK : ORBX_E_CUDA;
}

int orbx_debug_stage_times(orbx_ctx* c, const char** names, float* ms, int cap)
{
    if (!c) return ORBX_E_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    int n = 0;
    for (int i = 0; i < N_STAGES && n < cap; ++i) {
        if (!c->stage_valid[i]) continue;
        float t = 0.f;
        const int e0 = i < 6 ? i : 7;
        if (cudaEventElapsedTime(&t, c->ev[e0], c->ev[e0 + 1]) != cudaSuccess) continue;
        if (names) names[n] = k_stage_names[i];
        if (ms) ms[n] = t;
        ++n;
    }
    return n;
}

}  // extern "C"

// =================================================================================================== map table / tracking
namespace {

// grow a table column, keeping its contents
int grow_keep(orbx_ctx* c, Buf& b, size_t old_bytes, size_t new_bytes)
{
    if (b.bytes >= new_bytes && b.p) return ORBX_OK;
    void* np = nullptr;
    if (cudaMalloc(&np, std::max<size_t>(new_bytes, 256)) != cudaSuccess) return fail(c, ORBX_E_NOMEM, "cudaMalloc failed (map table)");
    if (cudaMemsetAsync(np, 0, std::max<size_t>(new_bytes, 256), c->stream) != cudaSuccess) { cudaFree(np); return fail(c, ORBX_E_CUDA, "memset failed"); }
    if (b.p && old_bytes) {
        if (cudaMemcpyAsync(np, b.p, old_bytes, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) { cudaFree(np); return fail(c, ORBX_E_CUDA, "copy failed"); }
    }
    cudaStreamSynchronize(c->stream);
    if (b.p) cudaFree(b.p);
    b.p = np; b.bytes = std::max<size_t>(new_bytes, 256);
    return ORBX_OK;
}

int map_reserve(orbx_ctx* c, int capacity)
{
    if (capacity <= c->map_capacity) return ORBX_OK;
    int ncap = std::max(capacity, std::max(1024, c->map_capacity * 2));
    const size_t o = (size_t)c->map_capacity, n = (size_t)ncap;
    int rc;
    if ((rc = grow_keep(c, c->t_desc, o * 32, n * 32)) || (rc = grow_keep(c, c->t_pos, o * 24, n * 24)) || (rc = grow_keep(c, c->t_nrm, o * 24, n * 24)) ||
        (rc = grow_keep(c, c->t_outl, o, n)))
        return rc;
    c->map_capacity = ncap;
    return ORBX_OK;
}

// ids -> slots (allocating new slots when `create`); the slot list is left in c->t_slots on the device
int map_slots(orbx_ctx* c, const int64_t* ids, int n, bool create, std::vector<int>* slots)
{
    slots->resize((size_t)n);
    if (create) {
        int fresh = 0;
        for (int i = 0; i < n; ++i) if (!c->map_slot.count((long long)ids[i])) ++fresh;     // duplicates inside one call over-count: harmless
        const int need = c->map_next - (int)c->map_free.size() + fresh;
        int rc = map_reserve(c, std::max(need, c->map_next));
        if (rc) return rc;
    }
    for (int i = 0; i < n; ++i) {
        auto it = c->map_slot.find((long long)ids[i]);
        if (it != c->map_slot.end()) { (*slots)[i] = it->second; continue; }
        if (!create) return fail(c, ORBX_E_ARG, "unknown map-point id");
        int s;
        if (!c->map_free.empty()) { s = c->map_free.back(); c->map_free.pop_back(); }
        else {
            s = c->map_next++;
            if (s >= c->map_capacity) { int rc = map_reserve(c, s + 1); if (rc) return rc; }
        }
        c->map_slot.emplace((long long)ids[i], s);
        (*slots)[i] = s;
    }
    int rc = ensure(c, c->t_slots, sizeof(int) * (size_t)std::max(n, 1));
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->t_slots.p, slots->data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    return ORBX_OK;
}

}  // namespace

extern "C" {

int orbx_map_reserve(orbx_ctx* c, int capacity)
{
    if (!c || capacity < 0) return ORBX_E_ARG;
    CU(cudaSetDevice(c->device));
    return map_reserve(c, capacity);
}

int orbx_map_size(const orbx_ctx* c) { return c ? (int)c->map_slot.size() : 0; }

int orbx_map_clear(orbx_ctx* c)
{
    if (!c) return ORBX_E_ARG;
    c->map_slot.clear(); c->map_free.clear(); c->map_next = 0;
    return ORBX_OK;
}

int orbx_map_upsert(orbx_ctx* c, const int64_t* ids, int n, const uint8_t* desc, const double* pos, const double* norm, const uint8_t* outlier)
{
    if (!c || n < 0) return ORBX_E_ARG;
    if (n == 0) return ORBX_OK;
    if (!ids) return fail(c, ORBX_E_ARG, "null ids");
    CU(cudaSetDevice(c->device));
    std::vector<int> slots;
    std::vector<char> is_new((size_t)n);
    for (int i = 0; i < n; ++i) is_new[i] = !c->map_slot.count((long long)ids[i]);
    int rc = map_slots(c, ids, n, true, &slots);
    if (rc) return rc;
    // staging: [desc n*32][pos n*24][norm n*24][outlier n]; a column the caller leaves NULL keeps its value for existing
    // points and starts as zero for new ones (outlier_ = false, norm_ = 0: src/mappoint.cpp:35)
    const size_t N = (size_t)n, o_pos = N * 32, o_nrm = o_pos + N * 24, o_out = o_nrm + N * 24, total = o_out + N;
    if ((rc = ensure(c, c->t_in, total))) return rc;
    uint8_t* st = (uint8_t*)c->t_in.p;
    if (desc) CU(cudaMemcpyAsync(st, desc, N * 32, cudaMemcpyHostToDevice, c->stream));
    if (pos) CU(cudaMemcpyAsync(st + o_pos, pos, N * 24, cudaMemcpyHostToDevice, c->stream));
    if (norm) CU(cudaMemcpyAsync(st + o_nrm, norm, N * 24, cudaMemcpyHostToDevice, c->stream));
    if (outlier) CU(cudaMemcpyAsync(st + o_out, outlier, N, cudaMemcpyHostToDevice, c->stream));
    if (!desc || !pos || !norm || !outlier) {
        // zero the columns of NEW points that were not supplied: one scatter of zeros for those slots
        std::vector<int> ns;
        for (int i = 0; i < n; ++i) if (is_new[i]) ns.push_back(slots[i]);
        if (!ns.empty()) {
            const size_t Z = ns.size();
            const size_t zoff = round_up(Z * 4, 16);        // the zero block is read as uint4 / double: keep it 16-byte aligned
            if ((rc = ensure(c, c->t_aux, zoff + Z * 32))) return rc;
            CU(cudaMemcpyAsync(c->t_aux.p, ns.data(), Z * 4, cudaMemcpyHostToDevice, c->stream));
            uint8_t* z = (uint8_t*)c->t_aux.p + zoff;
            CU(cudaMemsetAsync(z, 0, Z * 32, c->stream));
            k_map_scatter<<<(unsigned)((Z + 255) / 256), 256, 0, c->stream>>>((const int*)c->t_aux.p, (int)Z, desc ? nullptr : z, pos ? nullptr : (const double*)z,
                                                                              norm ? nullptr : (const double*)z, outlier ? nullptr : z, (uint8_t*)c->t_desc.p,
                                                                              (double*)c->t_pos.p, (double*)c->t_nrm.p, (uint8_t*)c->t_outl.p);
            ++c->launches;
            CU(cudaStreamSynchronize(c->stream));            // `ns` is a local
        }
    }
    k_map_scatter<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const int*)c->t_slots.p, n, desc ? st : nullptr, pos ? (const double*)(st + o_pos) : nullptr,
                                                                      norm ? (const double*)(st + o_nrm) : nullptr, outlier ? st + o_out : nullptr,
                                                                      (uint8_t*)c->t_desc.p, (double*)c->t_pos.p, (double*)c->t_nrm.p, (uint8_t*)c->t_outl.p);
    ++c->launches;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));                    // host sources are borrowed only for the duration of the call
    return ORBX_OK;
}

int orbx_map_upsert_from_frame(orbx_ctx* c, const int64_t* ids, int n, int frame, const int32_t* kp_index, const double* pos, const double* norm)
{
    if (!c || n < 0) return ORBX_E_ARG;
    if (n == 0) return ORBX_OK;
    if (!ids || !kp_index) return fail(c, ORBX_E_ARG, "null pointer");
    if (!c->view_desc || c->out_cap <= 0 || frame < 0 || frame >= c->view_frames) return fail(c, ORBX_E_ARG, "no host-API extraction holds that frame");
    for (int i = 0; i < n; ++i) if (kp_index[i] < 0 || kp_index[i] >= c->out_cap) return fail(c, ORBX_E_ARG, "keypoint index out of range");
    int rc = orbx_map_upsert(c, ids, n, nullptr, pos, norm, nullptr);
    if (rc) return rc;
    std::vector<int> slots;
    if ((rc = map_slots(c, ids, n, false, &slots))) return rc;
    if ((rc = ensure(c, c->t_aux, sizeof(int) * (size_t)n))) return rc;
    CU(cudaMemcpyAsync(c->t_aux.p, kp_index, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    k_map_scatter_desc<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const int*)c->t_slots.p, (const int*)c->t_aux.p, n,
                                                                           c->view_desc + (size_t)frame * c->out_cap * 32, (uint8_t*)c->t_desc.p);
    ++c->launches;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return ORBX_OK;
}

int orbx_map_erase(orbx_ctx* c, const int64_t* ids, int n)
{
    if (!c || n < 0 || (n > 0 && !ids)) return ORBX_E_ARG;
    for (int i = 0; i < n; ++i) {
        auto it = c->map_slot.find((long long)ids[i]);
        if (it == c->map_slot.end()) continue;
        c->map_free.push_back(it->second);
        c->map_slot.erase(it);
    }
    return ORBX_OK;
}

int orbx_track_match(orbx_ctx* c, const int64_t* ids, int m, const double* pose_Tcw, const double* cam, int cols, int rows,
                     const uint8_t* train, int nt_or_frame, float match_ratio, int32_t* cand, int* n_cand, orbx_match* matches, int* n_matches,
                     float* min_dis, float* max_dis)
{
    if (!c) return ORBX_E_ARG;
    if (n_cand) *n_cand = 0;
    if (n_matches) *n_matches = 0;
    if (m < 0 || !pose_Tcw || !cam || !n_cand || !n_matches) return fail(c, ORBX_E_ARG, "bad arguments");
    if (m == 0) return ORBX_OK;
    if (!ids || !cand || !matches) return fail(c, ORBX_E_ARG, "null pointer");
    CU(cudaSetDevice(c->device));
    // train set: host rows, or frame `nt_or_frame` of the last host-API extraction (still on the device)
    const uint8_t* d_train = nullptr;
    const int* d_tcount = nullptr;
    int nt = 0, stride = 0;
    int rc;
    if (train) {
        nt = nt_or_frame;
        if (nt < 0) return fail(c, ORBX_E_ARG, "negative train count");
        if (nt > 0) {
            if ((rc = ensure(c, c->mt, (size_t)nt * 32))) return rc;
            CU(cudaMemcpyAsync(c->mt.p, train, (size_t)nt * 32, cudaMemcpyHostToDevice, c->stream));
            d_train = (const uint8_t*)c->mt.p;
        }
        stride = nt;
    } else {
        const int f = nt_or_frame;
        if (!c->view_desc || c->out_cap <= 0 || f < 0 || f >= c->view_frames) return fail(c, ORBX_E_ARG, "no host-API extraction holds that frame");
        d_train = c->view_desc + (size_t)f * c->out_cap * 32;
        d_tcount = c->view_counts + f;
        nt = stride = c->out_cap;
    }
    std::vector<int> slots;
    if ((rc = map_slots(c, ids, m, false, &slots))) return rc;
    if ((rc = ensure(c, c->t_cand, sizeof(int) * (size_t)m)) || (rc = ensure(c, c->t_ncand, 16)) || (rc = ensure(c, c->t_q, (size_t)m * 32)) ||
        (rc = ensure(c, c->t_best, (size_t)m * 16)) || (rc = ensure(c, c->t_filtered, (size_t)m * 16)) || (rc = ensure(c, c->t_minmax, 16)))
        return rc;
    Pose T; Cam K;
    for (int i = 0; i < 12; ++i) T.m[i] = pose_Tcw[i];
    K.fx = cam[0]; K.fy = cam[1]; K.cx = cam[2]; K.cy = cam[3];
    k_visibility<<<1, VIS_NT, 0, c->stream>>>((const int*)c->t_slots.p, m, T, K, cols, rows, (const double*)c->t_pos.p, (const double*)c->t_nrm.p,
                                              (const uint8_t*)c->t_outl.p, (int*)c->t_cand.p, (int*)c->t_ncand.p);
    ++c->launches;
    int* h = c->h_small;
    CU(cudaMemcpyAsync(h, c->t_ncand.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));                    // the candidate count sizes the matcher's grid
    const int nc = h[0];
    *n_cand = nc;
    if (nc > 0) CU(cudaMemcpyAsync(cand, c->t_cand.p, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, c->stream));
    if (nc == 0 || nt == 0) { CU(cudaStreamSynchronize(c->stream)); return ORBX_OK; }   // cv: empty query or train -> no matches
    k_gather_desc<<<(unsigned)((nc + 255) / 256), 256, 0, c->stream>>>((const int*)c->t_slots.p, (const int*)c->t_cand.p, nc, (const uint8_t*)c->t_desc.p,
                                                                       (uint8_t*)c->t_q.p);
    ++c->launches;
    if ((rc = run_match(c, (const uint8_t*)c->t_q.p, nc, d_train, nt, stride, d_tcount, 1, (int4*)c->t_best.p, nullptr))) return rc;
    k_filter_matches<<<1, VIS_NT, 0, c->stream>>>((const int4*)c->t_best.p, nc, match_ratio, (int4*)c->t_filtered.p, (int*)c->t_ncand.p + 1,
                                                  (float*)c->t_minmax.p);
    ++c->launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h, (int*)c->t_ncand.p + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(h + 1, c->t_minmax.p, 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(h + 3, c->mstatus.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h[3]) return match_timed_out(c);
    const int nm = h[0];
    if (nm > 0) CU(cudaMemcpy(matches, c->t_filtered.p, (size_t)nm * 16, cudaMemcpyDeviceToHost));
    *n_matches = nm;
    if (min_dis) memcpy(min_dis, h + 1, sizeof(float));
    if (max_dis) memcpy(max_dis, h + 2, sizeof(float));
    return ORBX_OK;
}

int orbx_backproject(orbx_ctx* c, const orbx_keypoint* kps, int n, const uint16_t* depth, int w, int h, size_t step_bytes, float depth_scale,
                     const double* cam, const double* pose_Tcw, double* pos_w, uint8_t* valid)
{
    if (!c || n < 0) return ORBX_E_ARG;
    if (n == 0) return ORBX_OK;
    if (!kps || !depth || !cam || !pose_Tcw || !pos_w || !valid || w <= 0 || h <= 0 || step_bytes < (size_t)w * 2 || (step_bytes & 1))
        return fail(c, ORBX_E_ARG, "bad arguments");
    CU(cudaSetDevice(c->device));
    const size_t N = (size_t)n, dbytes = step_bytes * (size_t)h;
    const size_t o_depth = round_up(N * 28, 256), o_pos = round_up(o_depth + dbytes, 256), o_valid = o_pos + N * 24, total = o_valid + N;
    int rc = ensure(c, c->t_in, total);
    if (rc) return rc;
    uint8_t* st = (uint8_t*)c->t_in.p;
    CU(cudaMemcpyAsync(st, kps, N * 28, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(st + o_depth, depth, dbytes, cudaMemcpyHostToDevice, c->stream));
    Pose T; Cam K;
    for (int i = 0; i < 12; ++i) T.m[i] = pose_Tcw[i];
    K.fx = cam[0]; K.fy = cam[1]; K.cx = cam[2]; K.cy = cam[3];
    k_backproject<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>((const float*)st, n, (const uint16_t*)(st + o_depth), w, h, step_bytes / 2,
                                                                      (double)depth_scale, K, T, (double*)(st + o_pos), st + o_valid);
    ++c->launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pos_w, st + o_pos, N * 24, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(valid, st + o_valid, N, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return ORBX_OK;
}

}  // extern "C"

// =================================================================================================== async single-frame pipeline
extern "C" {

int orbx_submit_frame(orbx_ctx* c, const uint8_t* img, int w, int h, size_t step, int channels)
{
    if (!c) return ORBX_E_ARG;
    if (!img || w <= 0 || h <= 0) return fail(c, ORBX_E_ARG, "empty or null image (use the synchronous call for empty frames)");
    if (w > c->max_w || h > c->max_h) return fail(c, ORBX_E_ARG, "frame larger than the context maximum");
    if (channels != 1 && channels != 3) return fail(c, ORBX_E_UNSUPPORTED, "channels must be 1 (gray) or 3 (BGR)");
    if (step < (size_t)w * channels) return fail(c, ORBX_E_ARG, "step smaller than a row");
    if (c->a_inflight >= 2) return fail(c, ORBX_E_BUSY, "two frames already in flight: collect one first");
    CU(cudaSetDevice(c->device));
    orbx_ctx::AsyncSlot& a = c->aslot[(c->a_head + c->a_inflight) & 1];
    const int cap = std::max(2 * c->nfeatures, 64);
    const size_t row = (size_t)w * channels, dstep = round_up(row, 16), fstride = dstep * h;
    const size_t out_bytes = 32 + (size_t)cap * 60;
    int rc;
    if (a.h_in_bytes < fstride) {
        if (a.h_in) cudaFreeHost(a.h_in);
        a.h_in = nullptr; a.h_in_bytes = 0;
        if (cudaMallocHost((void**)&a.h_in, fstride) != cudaSuccess) return fail(c, ORBX_E_NOMEM, "pinned allocation failed");
        a.h_in_bytes = fstride;
    }
    if (a.h_out_bytes < out_bytes) {
        if (a.h_out) cudaFreeHost(a.h_out);
        a.h_out = nullptr; a.h_out_bytes = 0;
        if (cudaMallocHost((void**)&a.h_out, out_bytes) != cudaSuccess) return fail(c, ORBX_E_NOMEM, "pinned allocation failed");
        a.h_out_bytes = out_bytes;
    }
    if (!a.done && cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming) != cudaSuccess) return fail(c, ORBX_E_CUDA, "event creation failed");
    if ((rc = ensure(c, c->in, fstride)) || (rc = ensure(c, a.d_kps, sizeof(orbx_keypoint) * (size_t)cap)) || (rc = ensure(c, a.d_desc, (size_t)32 * cap)) ||
        (rc = ensure(c, a.d_counts, 32)))
        return rc;
    a.cap = cap;
    if (c->view_desc == (const uint8_t*)a.d_desc.p) {         // the last collected frame lives in this slot: its descriptors are about to be
        c->view_desc = nullptr; c->view_counts = nullptr; c->view_frames = 0;   // overwritten, so "frame 0 of the last extraction" no longer exists
    }
    if ((rc = set_geometry(c, w, h))) return rc;
    for (int y = 0; y < h; ++y) memcpy(a.h_in + (size_t)y * dstep, img + (size_t)y * step, row);     // the caller's buffer is free on return
    // everything below is stream-ordered behind the previous frame: the shared per-frame workspace is reused safely
    CU(cudaMemcpyAsync(c->in.p, a.h_in, fstride, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->status.p, 0, sizeof(int), c->stream));
    if ((rc = run_extract_range(c, c->stream, 0, false, 0, 1, (const uint8_t*)c->in.p, dstep, fstride, channels, (float*)a.d_kps.p,
                                (uint8_t*)a.d_desc.p, cap, (int*)a.d_counts.p)))
        return rc;
    CU(cudaMemcpyAsync(a.h_out, a.d_counts.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(a.h_out + 16, c->status.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(a.h_out + 32, a.d_kps.p, sizeof(orbx_keypoint) * (size_t)cap, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(a.h_out + 32 + (size_t)cap * 28, a.d_desc.p, (size_t)32 * cap, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(a.done, c->stream));
    CU(cudaGetLastError());
    a.busy = true;
    ++c->a_inflight;
    return ORBX_OK;
}

int orbx_collect_frame(orbx_ctx* c, orbx_keypoint* kps, uint8_t* desc, int cap, int* n_out)
{
    if (!c || !n_out) return ORBX_E_ARG;
    *n_out = 0;
    if (c->a_inflight <= 0) return fail(c, ORBX_E_ARG, "no frame in flight");
    if (cap < 0 || (cap > 0 && (!kps || !desc))) return fail(c, ORBX_E_ARG, "null output");
    CU(cudaSetDevice(c->device));
    orbx_ctx::AsyncSlot& a = c->aslot[c->a_head];
    CU(cudaEventSynchronize(a.done));
    int n, status;
    memcpy(&n, a.h_out, sizeof n);
    memcpy(&status, a.h_out + 16, sizeof status);
    *n_out = n;
    if (n > a.cap) {                                          // more ties than the internal capacity: redo synchronously is the caller's fallback
        a.busy = false; c->a_head ^= 1; --c->a_inflight;
        return fail(c, ORBX_E_CAPACITY, "internal capacity exceeded; use orbx_detect_and_compute for this frame");
    }
    if (status) {                                             // device-side flag (selection fallback bound / TMA timeout): never hand out a possibly wrong frame
        a.busy = false; c->a_head ^= 1; --c->a_inflight;
        *n_out = 0;
        return fail(c, ORBX_E_INTERNAL, "device-side status set for the collected frame");
    }
    if (n > cap) return fail(c, ORBX_E_CAPACITY, "output capacity too small; n_out holds the needed count (frame still in flight)");
    memcpy(kps, a.h_out + 32, sizeof(orbx_keypoint) * (size_t)n);
    memcpy(desc, a.h_out + 32 + (size_t)a.cap * 28, (size_t)32 * n);
    // the collected frame becomes "frame 0 of the last extraction" for orbx_track_match / orbx_map_upsert_from_frame
    c->view_desc = (const uint8_t*)a.d_desc.p; c->view_counts = (const int*)a.d_counts.p; c->view_frames = 1; c->out_cap = a.cap;
    a.busy = false;
    c->a_head ^= 1;
    --c->a_inflight;
    return ORBX_OK;
}

}  // extern "C"
