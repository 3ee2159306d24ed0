// examples/frontend_shim.cpp -- the FrontEnd shim of INTEGRATION.md section 1 as a compilable C++11 translation unit.
//
// The reference's front-end (src/frontend.cpp) is C++ on OpenCV types; OpenCV's C++ headers are not in this image, so the three
// types the shim touches are declared here with OpenCV's own field layout (cv::KeyPoint 28 bytes, cv::DMatch 16 bytes, and the
// handful of cv::Mat members the shim uses).  Everything from "---- the shim" on is the text a maintainer pastes into
// src/frontend.cpp:33-37, :150-154 and :186-187.  main() feeds it one synthetic BGR frame (argv[1]: raw 640 x 480 x 3 bytes)
// and prints the keypoint count and byte checksums, which tests/test_abi.py (build, no GPU) and tests/test_gpu_parity.py (run)
// compare with the Python binding's result for the same frame.
//   g++ -std=c++11 -Iinclude examples/frontend_shim.cpp -Lrgbd_visualodometry_b200 -lorbx -Wl,-rpath,$PWD/rgbd_visualodometry_b200
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "orbx.h"

namespace cv {                                   // layout stand-ins for the three OpenCV types (see the header comment)
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
enum { CV_8U = 0 };
struct Mat {
    std::vector<uint8_t> store;
    uint8_t* data = nullptr;
    int rows = 0, cols = 0, ch = 1;
    size_t step = 0;
    int channels() const { return ch; }
    void create(int r, int c, int /*type*/) { rows = r; cols = c; ch = 1; step = (size_t)c; store.assign((size_t)r * c, 0); data = store.data(); }
    Mat rowRange(int r0, int r1) const { Mat m; m.create(r1 - r0, cols, CV_8U); memcpy(m.data, data + (size_t)r0 * step, (size_t)(r1 - r0) * step); return m; }
};
}  // namespace cv
using std::vector;

struct Frame { cv::Mat color_; };
struct Config {                                   // config/default.yaml:18-21
    template <class T> static T get(const char* key);
};
template <> int Config::get<int>(const char* key) { return !strcmp(key, "number_of_features") ? 500 : 8; }
template <> double Config::get<double>(const char*) { return 1.2; }

// ---- the shim ------------------------------------------------------------------------------------
class FrontEnd {
public:
    FrontEnd()
    {
        static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint) && sizeof(cv::DMatch) == sizeof(orbx_match), "layout");
        if (orbx_create(&orbx_, /*device*/0, Config::get<int>("number_of_features"), (float)Config::get<double>("scale_factor"),
                        Config::get<int>("level_pyramid"), /*max_w*/1920, /*max_h*/1080, /*max_batch*/1) != ORBX_OK)
            throw std::runtime_error("orbx_create failed: no sm_100 CUDA device");   // there is no CPU fallback
    }
    ~FrontEnd() { orbx_destroy(orbx_); }

    void ExtractKeyPointsAndComputeDescriptors()                                     // src/frontend.cpp:150-154
    {
        const cv::Mat& img = frameCurr_->color_;
        int cap = 2 * Config::get<int>("number_of_features"), n = 0;                 // retainBest keeps ties: n may exceed nfeatures
        for (;;) {
            keypointsCurr_.resize(cap);
            descriptorsCurr_.create(cap, 32, cv::CV_8U);
            int rc = orbx_detect_and_compute(orbx_, img.data, img.cols, img.rows, img.step, img.channels(),
                                             reinterpret_cast<orbx_keypoint*>(keypointsCurr_.data()), descriptorsCurr_.data, cap, &n);
            if (rc == ORBX_E_CAPACITY) { cap = n; continue; }                        // never truncated silently: retry with the need
            if (rc != ORBX_OK) throw std::runtime_error(orbx_last_error(orbx_));
            break;
        }
        keypointsCurr_.resize(n);
        descriptorsCurr_ = n ? descriptorsCurr_.rowRange(0, n) : cv::Mat();
    }

    vector<cv::DMatch> MatchAgainst(const cv::Mat& mptCandidatesDescriptors)         // src/frontend.cpp:186-187
    {
        vector<cv::DMatch> matches(mptCandidatesDescriptors.rows);
        int nm = 0;                                                                  // query = map candidates, train = frame
        if (orbx_match_hamming(orbx_, mptCandidatesDescriptors.data, mptCandidatesDescriptors.rows, descriptorsCurr_.data, descriptorsCurr_.rows,
                               reinterpret_cast<orbx_match*>(matches.data()), &nm) != ORBX_OK)
            throw std::runtime_error(orbx_last_error(orbx_));
        matches.resize(nm);
        return matches;
    }

    Frame* frameCurr_ = nullptr;
    vector<cv::KeyPoint> keypointsCurr_;
    cv::Mat descriptorsCurr_;

private:
    orbx_ctx* orbx_ = nullptr;
};
// ---- end of the shim --------------------------------------------------------------------------------------------------------

static uint32_t fnv1a(const void* p, size_t n)
{
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; ++i) { h ^= ((const uint8_t*)p)[i]; h *= 16777619u; }
    return h;
}

int main(int argc, char** argv)
{
    try {
        FrontEnd fe;
        if (argc < 2) { printf("shim ok: context created (pass a raw 640x480 BGR file to extract)\n"); return 0; }
        Frame fr;
        fr.color_.create(480, 640 * 3, cv::CV_8U);
        fr.color_.cols = 640; fr.color_.ch = 3; fr.color_.step = 640 * 3;
        FILE* f = fopen(argv[1], "rb");
        if (!f || fread(fr.color_.data, 1, 640 * 480 * 3, f) != 640 * 480 * 3) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
        fclose(f);
        fe.frameCurr_ = &fr;
        fe.ExtractKeyPointsAndComputeDescriptors();
        cv::Mat map;                                                                  // "map" = every third descriptor of the frame itself
        const int m = fe.descriptorsCurr_.rows / 3;
        map.create(m, 32, cv::CV_8U);
        for (int i = 0; i < m; ++i) memcpy(map.data + (size_t)i * 32, fe.descriptorsCurr_.data + (size_t)3 * i * 32, 32);
        vector<cv::DMatch> matches = fe.MatchAgainst(map);
        printf("shim: keypoints %d kp_fnv %08x desc_fnv %08x matches %d match_fnv %08x\n", (int)fe.keypointsCurr_.size(),
               fnv1a(fe.keypointsCurr_.data(), fe.keypointsCurr_.size() * sizeof(cv::KeyPoint)),
               fnv1a(fe.descriptorsCurr_.data, (size_t)fe.descriptorsCurr_.rows * 32), (int)matches.size(),
               fnv1a(matches.data(), matches.size() * sizeof(cv::DMatch)));
        return 0;
    } catch (const std::exception& e) {
        printf("shim: %s\n", e.what());
        return 3;                                                                    // no GPU: the constructor throws, nothing falls back
    }
}
