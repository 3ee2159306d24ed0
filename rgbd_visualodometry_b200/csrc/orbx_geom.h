// orbx_geom.h -- pyramid geometry and the device-memory layout of one extraction batch (host + device).
//
// Everything a kernel needs to address a frame's data is in `Geom`, passed by value (__grid_constant__).
// Layout is FRAME-MAJOR: each buffer has a per-frame stride, and inside a frame's block every pyramid level
// sits at a fixed offset, so one frame's working set stays together in L2 across the pipeline.
//
//   pyr     u8    [frame][level: h_l rows x pitch_l]            the unblurred pyramid (level 0 = gray)
//   blur    u8    same layout as pyr                            7x7 Gaussian of each level (only where descriptors sample)
//   rowcnt  u32   [frame][level: inner row r = y-31]            FAST+NMS survivors per inner row
//   rowent  u32   [frame][level: inner row r][ent_pitch_l]      survivors of that row in x order: x | score<<16
//   work    Elem  [frame][level: ws_cap_l]                      selection workspace; its prefix is the final list
//   fincnt  i32   [frame][level]                                final keypoints per level
//   status  i32   [frame]                                       deferred per-frame status bits
#pragma once
#include <stdint.h>

#define ORBX_EDGE 31          // cv::ORB edgeThreshold (OpenCV default)
#define ORBX_FAST_T 20        // cv::ORB fastThreshold
#define ORBX_LEVELS_MAX 16

struct Elem {                 // one selection candidate: response (FAST score, later Harris) + packed position
    float response;
    uint32_t pos;             // y << 16 | x  (level coordinates)
};

struct LevelGeom {
    int w, h, pitch;          // pitch: bytes per row, multiple of 16
    int quota;                // n_l  
    float scale;              // F(pow((double)scaleFactor, l))
    int in_w, in_h;           // inner (border-filtered) extent: max(w-62,0), max(h-62,0)
    int ent_pitch;            // row-list capacity (entries) per inner row
    int ws_cap;               // workspace capacity (entries): ceil(in_w/2)*ceil(in_h/2)
    int band0;                // index of this level's first band in the per-frame band list
    int nbands;
    int hblk0, hblk;          // k_harris: first block of this level in the per-frame block list, blocks of the level
    int blur0, nblur, blur_cgs;  // blur tiles: first tile index, tile count (= column groups x strip groups), column groups
    int blur_rh;              // output rows per blur strip: a multiple of 7 chosen per level so the strips fit the level tightly
    uint32_t xtab, ytab;      // offsets (u32 units) of the INTER_LINEAR_EXACT tap tables: i0 | c1 << 16
    int pt_ok;                // 1: the TMA pyramid kernel (k_pyr_tma) produces this level; 0: the register-window kernel k_pyr_down
    int pt_ncx, pt_k, pt_ntask;   // column tiles of 128 outputs, strips per warp task, warp tasks per frame
    int pt_bw, pt_bh;         // TMA box (bytes x rows) of source level l-1 that covers one (column tile, strip)
    int xspan;                // max over 4-column groups of (left tap of the group's last column) - (its aligned first byte): <= 7 narrow kernel, <= 11 wide
    unsigned long long img_off;   // bytes, inside the frame's pyr block
    unsigned long long cnt_off;   // u32 units, inside the frame's rowcnt block
    unsigned long long ent_off;   // u32 units, inside the frame's rowent block
    unsigned long long ws_off;    // Elem units, inside the frame's work block
};

struct Geom {
    int nlevels, w, h;
    int total_bands;          // per frame
    int total_blur;           // blur tiles per frame
    int total_hblk;           // k_harris blocks per frame
    int selh_elems;           // k_select_harris: candidates of one level held in shared memory
    int band_rows;            // rows per FAST band
    unsigned long long pyr_frame, cnt_frame, ent_frame, ws_frame;   // per-frame strides (bytes / u32 / u32 / Elem)
    LevelGeom L[ORBX_LEVELS_MAX];
};
