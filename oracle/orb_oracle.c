/*
 * orb_oracle.c -- CPU restatement of the ORB-extract + brute-force-Hamming-match hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product path is the CUDA library (liborbx.so).
 *
 * What it restates.  The reference (BowenBZ/RGBD_VisualOdometry) runs this path through two
 * OpenCV operators:
 *   - cv::ORB::detectAndCompute        src/frontend.cpp:153  (operator built at src/frontend.cpp:35-37)
 *   - cv::DescriptorMatcher::match     src/frontend.cpp:187  (query = map candidates, train = frame)
 * OpenCV (pinned 3.1 by CMakeLists.txt:23) is an un-vendored dependency that is absent from
 * /root/reference, so the arithmetic below restates OpenCV's published algorithm stage by stage as
 * specified in SURVEY.md Appendix A (A.1 gray, A.2 pyramid, A.3 FAST+NMS, A.4/A.6 retainBest in
 * libstdc++ order, A.5 Harris, A.7 IC angle, A.8 blur, A.9 steered rBRIEF, A.10 record, A.11 match).
 *
 * Parity pin: this restatement is checked bit-for-bit (every KeyPoint field, keypoint order, all
 * descriptor bytes, all DMatch fields) against the in-image OpenCV build (cv2 4.13.0) by
 * tests/test_oracle.py and against the committed fixtures under tests/golden/
 * (generated from cv2 by tools/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC  (see oracle/Makefile).
 * Every float expression that OpenCV rounds separately is written as separate statements; the two
 * places OpenCV's AVX2 build really uses FMA (the blur, A.8) call fmaf() explicitly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORBO_MAX_LEVELS 16
#define ORBO_EDGE 31
#define ORBO_FAST_T 20
#define ORBO_HALF_PATCH 15

typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orbo_keypoint; /* == cv::KeyPoint */
typedef struct { int32_t queryIdx, trainIdx, imgIdx; float distance; } orbo_match;            /* == cv::DMatch  */
typedef struct { int32_t x, y; float response; } orbo_cand;                                    /* level-space candidate */

static const int16_t k_pattern[256 * 4] = {
#include "brief_pattern.inc"
};

static inline int orbo_rint_f(float v) { return (int)lrintf(v); }   /* cvRound: round-half-even */

/* ---------------------------------------------------------------- A.1 gray */
void orbo_gray(const uint8_t* bgr, int w, int h, size_t step, uint8_t* gray)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t* p = bgr + (size_t)y * step;
        uint8_t* g = gray + (size_t)y * w;
        for (int x = 0; x < w; ++x, p += 3)
            g[x] = (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15);
    }
}

/* ---------------------------------------------------------------- A.2 pyramid geometry */
void orbo_level_geometry(int w, int h, int nlevels, float scale_factor, int* ws, int* hs, float* scales)
{
    for (int l = 0; l < nlevels; ++l) {
        float s = (float)pow((double)scale_factor, (double)l);
        float inv = 1.0f / s;
        scales[l] = s;
        ws[l] = orbo_rint_f((float)w * inv);
        hs[l] = orbo_rint_f((float)h * inv);
    }
}

/* per-axis taps of INTER_LINEAR_EXACT: ofs[i], c1[i] in 8.8 fixed point (c0 = 256 - c1). */
void orbo_resize_taps(int s, int d, int32_t* ofs, int32_t* c1)
{
    double inv_scale = (double)d / (double)s;
    double scale = 1.0 / inv_scale;
    for (int x = 0; x < d; ++x) {
        double f = scale * ((double)x + 0.5) - 0.5;
        int i = (int)floor(f);
        if (i < 0 || s <= 1) { ofs[x] = 0; c1[x] = 0; }
        else if (i >= s - 1) { ofs[x] = s - 1; c1[x] = 0; }
        else { ofs[x] = i; c1[x] = (int)lrint((f - (double)i) * 256.0); }
    }
}

void orbo_resize_exact(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh)
{
    int32_t* xo = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(dw + dh));
    int32_t* xc = xo + dw; int32_t* yo = xc + dw; int32_t* yc = yo + dh;
    orbo_resize_taps(sw, dw, xo, xc);
    orbo_resize_taps(sh, dh, yo, yc);
    for (int y = 0; y < dh; ++y) {
        const uint8_t* r0 = src + (size_t)yo[y] * sw;
        const uint8_t* r1 = src + (size_t)(yo[y] + (yc[y] ? 1 : 0)) * sw;
        uint32_t cy1 = (uint32_t)yc[y], cy0 = 256u - cy1;
        for (int x = 0; x < dw; ++x) {
            int i0 = xo[x], i1 = xo[x] + (xc[x] ? 1 : 0);
            uint32_t cx1 = (uint32_t)xc[x], cx0 = 256u - cx1;
            uint32_t h0 = r0[i0] * cx0 + r0[i1] * cx1;      /* 8.8, exact */
            uint32_t h1 = r1[i0] * cx0 + r1[i1] * cx1;
            uint32_t v = (h0 * cy0 + h1 * cy1 + 32768u) >> 16;
            dst[(size_t)y * dw + x] = (uint8_t)(v > 255u ? 255u : v);
        }
    }
    free(xo);
}

/* ---------------------------------------------------------------- A.3 FAST-9/16 score + 3x3 NMS */
static const int k_circle[16][2] = {
    {0,3},{1,3},{2,2},{3,1},{3,0},{3,-1},{2,-2},{1,-3},{0,-3},{-1,-3},{-2,-2},{-3,-1},{-3,0},{-3,1},{-2,2},{-1,3}};

/* returns 0 for a non-corner, else the FAST score (>= t) */
static int orbo_fast_score_px(const uint8_t* c, int stride, int t)
{
    int d[25];
    int v = c[0];
    for (int k = 0; k < 16; ++k) d[k] = v - c[k_circle[k][1] * stride + k_circle[k][0]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int A = -1000, Bm = 1000;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { if (d[k + j] < mn) mn = d[k + j]; if (d[k + j] > mx) mx = d[k + j]; }
        if (mn > A) A = mn;
        if (mx < Bm) Bm = mx;
    }
    if (!(A > t || Bm < -t)) return 0;
    int a0 = A > t ? A : t;
    int b0 = Bm < -a0 ? Bm : -a0;
    return -b0 - 1;
}

/* score map (0 = not a corner) over the whole level; rows/cols closer than 3 px to the edge are 0 */
void orbo_fast_score_map(const uint8_t* img, int w, int h, int t, uint8_t* score)
{
    memset(score, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x)
            score[(size_t)y * w + x] = (uint8_t)orbo_fast_score_px(img + (size_t)y * w + x, w, t);
}

/* raster-ordered NMS survivors; returns the count (writes at most cap) */
int orbo_fast_nms(const uint8_t* img, int w, int h, int t, orbo_cand* out, int cap)
{
    if (w < 7 || h < 7) return 0;
    uint8_t* sc = (uint8_t*)malloc((size_t)w * h);
    orbo_fast_score_map(img, w, h, t, sc);
    int n = 0;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = sc + (size_t)y * w + x;
            int s = p[0];
            if (!s) continue;
            if (s > p[-1] && s > p[1] && s > p[-w - 1] && s > p[-w] && s > p[-w + 1] &&
                s > p[w - 1] && s > p[w] && s > p[w + 1]) {
                if (n < cap) { out[n].x = x; out[n].y = y; out[n].response = (float)s; }
                ++n;
            }
        }
    free(sc);
    return n;
}

/* ---------------------------------------------------------------- A.4 quotas */
void orbo_quotas(int nfeatures, float scale_factor, int nlevels, int* n_l)
{
    float factor = (float)(1.0 / (double)scale_factor);
    float one_minus = 1.0f - factor;
    float num = (float)nfeatures * one_minus;
    float den = 1.0f - (float)pow((double)factor, (double)nlevels);
    float nd = num / den;
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        n_l[l] = orbo_rint_f(nd);
        sum += n_l[l];
        nd = nd * factor;
    }
    n_l[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

/* ---------------------------------------------------------------- A.6 retainBest, libstdc++ order */
#define GT(p, q) (v[p].response > v[q].response)
static inline void cswap(orbo_cand* v, int a, int b) { orbo_cand t = v[a]; v[a] = v[b]; v[b] = t; }

/* libstdc++ heap primitives (bits/stl_heap.h) with comp(a, b) = a.response > b.response, on the range v[first ..):
 * __push_heap, __adjust_heap, __make_heap, __pop_heap, __heap_select -- the depth-limit fallback of __introselect.
 * Indices are relative to `first`, exactly as the iterator arithmetic of the library. */
static void orbo_push_heap(orbo_cand* b, int hole, int top, orbo_cand value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && b[parent].response > value.response) {
        b[hole] = b[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    b[hole] = value;
}
static void orbo_adjust_heap(orbo_cand* b, int hole, int len, orbo_cand value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (b[child].response > b[child - 1].response) --child;
        b[hole] = b[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        b[hole] = b[child - 1];
        hole = child - 1;
    }
    orbo_push_heap(b, hole, top, value);
}
/* how often the fallback ran (tests assert that their tie-heavy cases really exercise it) */
int orbo_heap_select_calls = 0;
static void orbo_heap_select(orbo_cand* v, int first, int middle, int last)
{
    ++orbo_heap_select_calls;
    orbo_cand* b = v + first;
    const int len = middle - first;
    if (len >= 2) {                                           /* __make_heap */
        int parent = (len - 2) / 2;
        for (;;) {
            orbo_cand value = b[parent];
            orbo_adjust_heap(b, parent, len, value);
            if (parent == 0) break;
            --parent;
        }
    }
    for (int i = middle; i < last; ++i)
        if (v[i].response > b[0].response) {                  /* __pop_heap(first, middle, i) */
            orbo_cand value = v[i];
            v[i] = b[0];
            orbo_adjust_heap(b, 0, len, value);
        }
}

/* std::nth_element (libstdc++ __introselect), including its depth-limit fallback (__heap_select + iter_swap). */
static int orbo_nth_element(orbo_cand* v, int first, int nth, int last)
{
    if (first == last || nth == last) return 0;
    int len = last - first, lg = 0;
    while ((len >> (lg + 1)) > 0) ++lg;
    int depth = 2 * lg;
    while (last - first > 3) {
        if (depth == 0) {
            orbo_heap_select(v, first, nth + 1, last);
            cswap(v, first, nth);
            return 0;
        }
        --depth;
        int mid = first + (last - first) / 2;
        int a = first + 1, b = mid, c = last - 1, pick;
        if (GT(a, b)) pick = GT(b, c) ? b : (GT(a, c) ? c : a);
        else          pick = GT(a, c) ? a : (GT(b, c) ? c : b);
        cswap(v, first, pick);
        int f = first + 1, l = last;
        for (;;) {
            while (GT(f, first)) ++f;
            --l;
            while (GT(first, l)) --l;
            if (!(f < l)) break;
            cswap(v, f, l);
            ++f;
        }
        if (f <= nth) first = f; else last = f;
    }
    /* __insertion_sort, descending by response */
    for (int i = first + 1; i < last; ++i) {
        orbo_cand val = v[i];
        if (val.response > v[first].response) {
            memmove(v + first + 1, v + first, sizeof(orbo_cand) * (size_t)(i - first));
            v[first] = val;
        } else {
            int j = i;
            while (val.response > v[j - 1].response) { v[j] = v[j - 1]; --j; }
            v[j] = val;
        }
    }
    return 0;
}

/* KeyPointsFilter::retainBest.  Returns the new length, or -1 on the flagged fallback. */
int orbo_retain_best(orbo_cand* v, int len, int m)
{
    if (m < 0 || len <= m) return len;
    if (m == 0) return 0;
    if (orbo_nth_element(v, 0, m - 1, len) < 0) return -1;
    float thr = v[m - 1].response;
    int f = m, l = len;                       /* bidirectional std::partition, pred = response >= thr */
    for (;;) {
        for (;;) { if (f == l) return f; if (v[f].response >= thr) ++f; else break; }
        --l;
        for (;;) { if (f == l) return f; if (!(v[l].response >= thr)) --l; else break; }
        cswap(v, f, l);
        ++f;
    }
}

/* ---------------------------------------------------------------- A.5 Harris */
float orbo_harris(const uint8_t* img, int stride, int x, int y)
{
    int a = 0, b = 0, c = 0;
    for (int dy = -3; dy <= 3; ++dy)
        for (int dx = -3; dx <= 3; ++dx) {
            const uint8_t* p = img + (size_t)(y + dy) * stride + (x + dx);
            int Ix = (p[1] - p[-1]) * 2 + (p[-stride + 1] - p[-stride - 1]) + (p[stride + 1] - p[stride - 1]);
            int Iy = (p[stride] - p[-stride]) * 2 + (p[stride - 1] - p[-stride - 1]) + (p[stride + 1] - p[-stride + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    float fa = (float)a, fb = (float)b, fc = (float)c;
    float s = 1.0f / (4.0f * 7.0f * 255.0f);
    float s2 = s * s; float s3 = s2 * s; float s4 = s3 * s;
    float ab = fa * fb;
    float cc = fc * fc;
    float det = ab - cc;
    float tr = fa + fb;
    float ktr = 0.04f * tr;
    float ktr2 = ktr * tr;
    float r = det - ktr2;
    return r * s4;
}

/* ---------------------------------------------------------------- A.7 IC angle */
static const int k_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

float orbo_fast_atan2(float y, float x)
{
    const float P1 = 57.283627f, P3 = -18.667446f, P5 = 8.9140005f, P7 = -2.5397246f;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        float den = ax + (float)2.220446049250313e-16;
        c = ay / den; c2 = c * c;
        a = P7 * c2; a = a + P5; a = a * c2; a = a + P3; a = a * c2; a = a + P1; a = a * c;
    } else {
        float den = ay + (float)2.220446049250313e-16;
        c = ax / den; c2 = c * c;
        a = P7 * c2; a = a + P5; a = a * c2; a = a + P3; a = a * c2; a = a + P1; a = a * c;
        a = 90.0f - a;
    }
    if (x < 0) a = 180.0f - a;
    if (y < 0) a = 360.0f - a;
    return a;
}

void orbo_ic_moments(const uint8_t* img, int stride, int x, int y, int* m01o, int* m10o)
{
    const uint8_t* c = img + (size_t)y * stride + x;
    int m01 = 0, m10 = 0;
    for (int u = -ORBO_HALF_PATCH; u <= ORBO_HALF_PATCH; ++u) m10 += u * c[u];
    for (int v = 1; v <= ORBO_HALF_PATCH; ++v) {
        int vs = 0, d = k_umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = c[u + v * stride], vm = c[u - v * stride];
            vs += vp - vm;
            m10 += u * (vp + vm);
        }
        m01 += v * vs;
    }
    *m01o = m01; *m10o = m10;
}

float orbo_ic_angle(const uint8_t* img, int stride, int x, int y)
{
    int m01, m10;
    orbo_ic_moments(img, stride, x, y, &m01, &m10);
    return orbo_fast_atan2((float)m01, (float)m10);
}

/* ---------------------------------------------------------------- A.8 blur (float sepFilter2D, FMA order) */
static const uint32_t k_gauss_bits[7] = {0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u, 0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u};
static inline float bits2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline int reflect101(int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; } return p; }

void orbo_blur7(const uint8_t* src, int w, int h, uint8_t* dst)
{
    float k[7];
    for (int i = 0; i < 7; ++i) k[i] = bits2f(k_gauss_bits[i]);
    float* rows = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * w;
        float* r = rows + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = k[0] * (float)s[reflect101(x - 3, w)];
            for (int i = 1; i < 7; ++i) acc = fmaf((float)s[reflect101(x - 3 + i, w)], k[i], acc);
            r[x] = acc;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float acc = k[3] * rows[(size_t)y * w + x];
            for (int j = 1; j <= 3; ++j) {
                float sum = rows[(size_t)reflect101(y + j, h) * w + x] + rows[(size_t)reflect101(y - j, h) * w + x];
                acc = fmaf(sum, k[3 + j], acc);
            }
            int v = (int)lrintf(acc);
            dst[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    free(rows);
}

/* ---------------------------------------------------------------- A.9 glibc 2.39 sinf/cosf in FP64 */
static const double kC0 = 1.0, kC1 = -0x1.ffffffd0c621cp-2, kC2 = 0x1.55553e1068f19p-5,
                    kC3 = -0x1.6c087e89a359dp-10, kC4 = 0x1.99343027bf8c3p-16;
static const double kS0 = -0x1.555545995a603p-3, kS1 = 0x1.1107605230bc4p-7, kS2 = -0x1.994eb3774cf24p-13;

static inline float sin_poly(double x, double x2) {
    double x3 = x * x2; double s1 = kS1 + x2 * kS2; double x7 = x3 * x2; double s = x + x3 * kS0;
    return (float)(s + x7 * s1);
}
static inline float cos_poly(double x2, double sgn) {
    double c0 = kC0 * sgn, c1v = kC1 * sgn, c2v = kC2 * sgn, c3v = kC3 * sgn, c4v = kC4 * sgn;
    double x4 = x2 * x2; double c2 = c3v + x2 * c4v; double c1 = c0 + x2 * c1v; double x6 = x4 * x2; double c = c1 + x4 * c2v;
    return (float)(c + x6 * c2);
}
static inline uint32_t abstop12(float x) { uint32_t b; memcpy(&b, &x, 4); return (b >> 20) & 0x7ff; }

/* valid for |ang| < 120 (rad) which covers [0, 2*pi] */
void orbo_sincosf(float ang, float* sn, float* cs)
{
    double x = (double)ang;
    if (abstop12(ang) < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (abstop12(ang) < abstop12(0x1p-12f)) { *sn = ang; *cs = 1.0f; return; }
        *sn = sin_poly(x, x2);
        *cs = cos_poly(x2, 1.0);
        return;
    }
    double r = x * 0x1.45F306DC9C883p+23;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = x - (double)n * 0x1.921FB54442D18p0;
    double x2 = x * x;
    static const double sign[4] = {1.0, -1.0, -1.0, 1.0};
    double s = sign[n & 3];
    double tsgn = (n & 2) ? -1.0 : 1.0;       /* glibc switches to the negated-C table when n&2 */
    /* sinf: poly parity n;  cosf: poly parity n^1 (same n, same sign, same table) */
    if (n & 1) { *sn = cos_poly(x2, tsgn);     *cs = sin_poly(x * s, x2); }
    else       { *sn = sin_poly(x * s, x2);    *cs = cos_poly(x2, tsgn); }
}

/* descriptor of one keypoint at integer level coordinates on the BLURRED level */
void orbo_brief(const uint8_t* blurred, int stride, int x, int y, float angle_deg, uint8_t* desc)
{
    float ang = angle_deg * 0.017453292f;
    float a, b;
    orbo_sincosf(ang, &b, &a);
    const uint8_t* c = blurred + (size_t)y * stride + x;
    for (int i = 0; i < 32; ++i) {
        unsigned byte = 0;
        for (int j = 0; j < 8; ++j) {
            const int16_t* p = k_pattern + (i * 8 + j) * 4;
            float x0a = (float)p[0] * a, y0b = (float)p[1] * b, x0b = (float)p[0] * b, y0a = (float)p[1] * a;
            float x1a = (float)p[2] * a, y1b = (float)p[3] * b, x1b = (float)p[2] * b, y1a = (float)p[3] * a;
            float fx0 = x0a - y0b, fy0 = x0b + y0a, fx1 = x1a - y1b, fy1 = x1b + y1a;
            int t0 = c[orbo_rint_f(fy0) * stride + orbo_rint_f(fx0)];
            int t1 = c[orbo_rint_f(fy1) * stride + orbo_rint_f(fx1)];
            byte |= (unsigned)(t0 < t1) << j;
        }
        desc[i] = (uint8_t)byte;
    }
}

/* ---------------------------------------------------------------- full operator: detectAndCompute */
/* error codes */
#define ORBO_OK 0
#define ORBO_E_CAP (-2)          /* output capacity too small; *n_out = needed */
#define ORBO_E_HEAPSELECT (-7)   /* libstdc++ heap_select fallback would be needed (flagged, not emulated) */
#define ORBO_E_ARG (-1)

/* Optional stage dump (any pointer may be NULL). */
typedef struct {
    uint8_t* levels[ORBO_MAX_LEVELS];      /* caller-allocated w_l*h_l each: unblurred level */
    uint8_t* blurred[ORBO_MAX_LEVELS];     /* caller-allocated: blurred level */
    orbo_cand* fast[ORBO_MAX_LEVELS];      /* caller-allocated, fast_cap entries: raster NMS list after border filter */
    int fast_cap;
    int n_fast[ORBO_MAX_LEVELS];           /* out: border-filtered FAST+NMS count */
    int n_sel1[ORBO_MAX_LEVELS];           /* out: after retainBest(2 n_l) */
    int n_sel2[ORBO_MAX_LEVELS];           /* out: after retainBest(n_l)  */
} orbo_dump;

int orbo_detect_and_compute(const uint8_t* img, int w, int h, size_t step, int channels,
                            int nfeatures, float scale_factor, int nlevels,
                            orbo_keypoint* kps, uint8_t* desc, int cap, int* n_out, orbo_dump* dump)
{
    if (n_out) *n_out = 0;
    if (!img || w <= 0 || h <= 0 || nlevels < 1 || nlevels > ORBO_MAX_LEVELS || (channels != 1 && channels != 3)) return ORBO_E_ARG;
    int ws[ORBO_MAX_LEVELS], hs[ORBO_MAX_LEVELS], nl[ORBO_MAX_LEVELS];
    float scales[ORBO_MAX_LEVELS];
    orbo_level_geometry(w, h, nlevels, scale_factor, ws, hs, scales);
    orbo_quotas(nfeatures, scale_factor, nlevels, nl);

    uint8_t* lev[ORBO_MAX_LEVELS];
    memset(lev, 0, sizeof lev);
    lev[0] = (uint8_t*)malloc((size_t)w * h);
    if (channels == 3) orbo_gray(img, w, h, step, lev[0]);
    else for (int y = 0; y < h; ++y) memcpy(lev[0] + (size_t)y * w, img + (size_t)y * step, (size_t)w);
    int rc = ORBO_OK, total = 0;
    for (int l = 1; l < nlevels; ++l) {
        if (ws[l] <= 0 || hs[l] <= 0) { ws[l] = hs[l] = 0; continue; }
        lev[l] = (uint8_t*)malloc((size_t)ws[l] * hs[l]);
        orbo_resize_exact(lev[l - 1], ws[l - 1], hs[l - 1], lev[l], ws[l], hs[l]);
    }
    for (int l = 0; l < nlevels && rc == ORBO_OK; ++l) {
        int lw = ws[l], lh = hs[l];
        if (dump) { dump->n_fast[l] = dump->n_sel1[l] = dump->n_sel2[l] = 0; }
        if (!lev[l]) continue;
        if (dump && dump->levels[l]) memcpy(dump->levels[l], lev[l], (size_t)lw * lh);
        int ncap = ((lw + 1) / 2) * ((lh + 1) / 2) + 1;
        orbo_cand* cand = (orbo_cand*)malloc(sizeof(orbo_cand) * (size_t)ncap);
        int n = orbo_fast_nms(lev[l], lw, lh, ORBO_FAST_T, cand, ncap);
        /* runByImageBorder */
        int m = 0;
        if (lw > 2 * ORBO_EDGE && lh > 2 * ORBO_EDGE)
            for (int i = 0; i < n; ++i)
                if (cand[i].x >= ORBO_EDGE && cand[i].x < lw - ORBO_EDGE && cand[i].y >= ORBO_EDGE && cand[i].y < lh - ORBO_EDGE)
                    cand[m++] = cand[i];
        n = m;
        if (dump) { dump->n_fast[l] = n; if (dump->fast[l]) memcpy(dump->fast[l], cand, sizeof(orbo_cand) * (size_t)(n < dump->fast_cap ? n : dump->fast_cap)); }
        n = orbo_retain_best(cand, n, 2 * nl[l]);
        if (n < 0) { rc = ORBO_E_HEAPSELECT; free(cand); break; }
        if (dump) dump->n_sel1[l] = n;
        for (int i = 0; i < n; ++i) cand[i].response = orbo_harris(lev[l], lw, cand[i].x, cand[i].y);
        n = orbo_retain_best(cand, n, nl[l]);
        if (n < 0) { rc = ORBO_E_HEAPSELECT; free(cand); break; }
        if (dump) dump->n_sel2[l] = n;
        uint8_t* bl = NULL;
        if (n > 0 || (dump && dump->blurred[l])) {
            bl = (uint8_t*)malloc((size_t)lw * lh);
            orbo_blur7(lev[l], lw, lh, bl);
            if (dump && dump->blurred[l]) memcpy(dump->blurred[l], bl, (size_t)lw * lh);
        }
        for (int i = 0; i < n; ++i, ++total) {
            if (total >= cap) continue;
            orbo_keypoint* k = kps + total;
            float ang = orbo_ic_angle(lev[l], lw, cand[i].x, cand[i].y);
            k->x = (float)cand[i].x * scales[l];
            k->y = (float)cand[i].y * scales[l];
            k->size = 31.0f * scales[l];
            k->angle = ang;
            k->response = cand[i].response;
            k->octave = l;
            k->class_id = -1;
            orbo_brief(bl, lw, cand[i].x, cand[i].y, ang, desc + (size_t)total * 32);
        }
        free(bl);
        free(cand);
    }
    for (int l = 0; l < nlevels; ++l) free(lev[l]);
    if (n_out) *n_out = total;
    if (rc == ORBO_OK && total > cap) rc = ORBO_E_CAP;
    return rc;
}

/* ---------------------------------------------------------------- A.11 brute-force Hamming */
static inline int hamming32(const uint8_t* a, const uint8_t* b)
{
    uint64_t x[4], y[4];
    memcpy(x, a, 32); memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

/* BFMatcher(NORM_HAMMING).match: one DMatch per query row; ties -> lowest train index. Returns count. */
int orbo_match_hamming(const uint8_t* query, int nq, const uint8_t* train, int nt, orbo_match* out)
{
    if (nq <= 0 || nt <= 0) return 0;
    for (int i = 0; i < nq; ++i) {
        int best = 1 << 30, bj = -1;
        for (int j = 0; j < nt; ++j) {
            int d = hamming32(query + (size_t)i * 32, train + (size_t)j * 32);
            if (d < best) { best = d; bj = j; }
        }
        out[i].queryIdx = i; out[i].trainIdx = bj; out[i].imgIdx = 0; out[i].distance = (float)best;
    }
    return nq;
}

/* knnMatch(k=2): two smallest by (distance, index); out has 2*nq entries (second = trainIdx -1 if nt < 2).
 * Returns the number of query rows. */
int orbo_match_hamming_knn2(const uint8_t* query, int nq, const uint8_t* train, int nt, orbo_match* out)
{
    if (nq <= 0 || nt <= 0) return 0;
    for (int i = 0; i < nq; ++i) {
        int b0 = 1 << 30, j0 = -1, b1 = 1 << 30, j1 = -1;
        for (int j = 0; j < nt; ++j) {
            int d = hamming32(query + (size_t)i * 32, train + (size_t)j * 32);
            if (d < b0) { b1 = b0; j1 = j0; b0 = d; j0 = j; }
            else if (d < b1) { b1 = d; j1 = j; }
        }
        out[2 * i].queryIdx = i; out[2 * i].trainIdx = j0; out[2 * i].imgIdx = 0; out[2 * i].distance = (float)b0;
        out[2 * i + 1].queryIdx = i; out[2 * i + 1].trainIdx = j1; out[2 * i + 1].imgIdx = 0;
        out[2 * i + 1].distance = j1 >= 0 ? (float)b1 : 0.0f;
    }
    return nq;
}

/* The reference's host-side post-filter (src/frontend.cpp:190-211): keep distance <= max(min*ratio, 30). */
int orbo_filter_matches(const orbo_match* in, int n, float ratio, orbo_match* out)
{
    if (n <= 0) return 0;
    float mn = in[0].distance;
    for (int i = 1; i < n; ++i) if (in[i].distance < mn) mn = in[i].distance;
    float mx = mn * ratio; if (mx < 30.0f) mx = 30.0f;
    int m = 0;
    for (int i = 0; i < n; ++i) if (in[i].distance <= mx) out[m++] = in[i];
    return m;
}
