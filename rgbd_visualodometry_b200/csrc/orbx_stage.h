// orbx_stage.h -- host side only: pageable caller buffers <-> pinned staging, copied by a small pool of threads.
//
// Why: the reference's frames are pageable cv::Mat (src/frame.cpp:28, filled by cv::imread at app/run_vo.cpp:91-92) and its
// results are std::vector<cv::KeyPoint> / cv::Mat / std::vector<cv::DMatch> (frontend.h:69-70).  cudaMemcpyAsync on pageable
// memory is staged by the driver synchronously and single-threaded (about 11 GB/s on the GPU box), and it blocks the calling
// thread, so the frame-range lanes of the host-buffer entry points stop overlapping.  Here the library stages such buffers
// itself: the copies pageable -> pinned (inputs) and pinned -> pageable (results) are cut into pieces and taken by
// ORBX_STAGE_THREADS copier threads plus the calling thread while it waits, every lane's H2D / D2H then runs from / into
// pinned memory as one asynchronous copy, and the lanes overlap as they do for a caller that pinned its buffers.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define ORBX_STAGE_NT 1
#endif
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace orbx {

class HostStager {
public:
    struct Job {                       // `rows` rows of `row_bytes` from src (pitch spitch) to dst (pitch dpitch)
        uint8_t* dst; const uint8_t* src; size_t dpitch, spitch, row_bytes, rows;
        std::atomic<int>* done;        // incremented once when the piece has landed
        bool nt;                       // destination is pinned staging that only the DMA engine will read: non-temporal stores
    };
    static constexpr size_t PIECE = 256 * 1024;

    explicit HostStager(int nthreads)
    {
        for (int i = 0; i < nthreads; ++i) th_.emplace_back([this] { worker(); });
    }
    ~HostStager()
    {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int threads() const { return (int)th_.size(); }

    // Queue a 2-D copy cut into pieces of about PIECE bytes (whole rows; one contiguous block when the pitches equal the
    // row length).  Returns the number of pieces, i.e. how much `done` will grow.
    // nt = true: the destination is written with non-temporal stores (no read-for-ownership of the destination lines, nothing of it
    // left in the CPU caches): right for pinned staging that is read next by the GPU's copy engine, wrong for results the caller reads.
    int submit(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows, std::atomic<int>* done, bool wake = true,
               bool nt = false)
    {
        if (rows == 0 || row_bytes == 0) return 0;
        if (dpitch == row_bytes && spitch == row_bytes) { row_bytes *= rows; dpitch = spitch = row_bytes; rows = 1; }
        int n = 0;
        std::lock_guard<std::mutex> lk(mu_);
        if (rows == 1) {
            for (size_t o = 0; o < row_bytes; o += PIECE, ++n)
                q_.push_back(Job{(uint8_t*)dst + o, (const uint8_t*)src + o, 0, 0, row_bytes - o < PIECE ? row_bytes - o : PIECE, 1, done, nt});
        } else {
            const size_t per = PIECE / row_bytes > 0 ? PIECE / row_bytes : 1;
            for (size_t r = 0; r < rows; r += per, ++n)
                q_.push_back(Job{(uint8_t*)dst + r * dpitch, (const uint8_t*)src + r * spitch, dpitch, spitch, row_bytes, rows - r < per ? rows - r : per, done, nt});
        }
        if (wake) { if (n > 1) cv_.notify_all(); else cv_.notify_one(); }   // (!wake: a small job the calling thread takes itself in help_until)
        return n;
    }

    // The calling thread copies queued pieces itself until `done` reaches `target` (pieces are taken in queue order, so the
    // ones it waits for are never behind work it would not do).
    void help_until(const std::atomic<int>& done, int target)
    {
        while (done.load(std::memory_order_acquire) < target) {
            Job j;
            bool got = false;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!q_.empty()) { j = q_.front(); q_.pop_front(); got = true; }
            }
            if (got) run(j); else std::this_thread::yield();
        }
    }

private:
    static void copy_nt(uint8_t* d, const uint8_t* s, size_t n)
    {
#ifdef ORBX_STAGE_NT
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
        if (head && head <= n) { memcpy(d, s, head); d += head; s += head; n -= head; }
        if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
            for (; n >= 64; n -= 64, d += 64, s += 64) {
                const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
                const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
                _mm_stream_si128(reinterpret_cast<__m128i*>(d), a); _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
                _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c); _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
            }
        }
#endif
        if (n) memcpy(d, s, n);
    }
    static void run(const Job& j)
    {
        if (j.nt) {
            for (size_t r = 0; r < j.rows; ++r) copy_nt(j.dst + r * j.dpitch, j.src + r * j.spitch, j.row_bytes);
#ifdef ORBX_STAGE_NT
            _mm_sfence();                                   // the streamed lines are globally visible before `done` says so
#endif
        } else {
            for (size_t r = 0; r < j.rows; ++r) memcpy(j.dst + r * j.dpitch, j.src + r * j.spitch, j.row_bytes);
        }
        j.done->fetch_add(1, std::memory_order_release);
    }
    void worker()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;              // stop_ and drained
                j = q_.front(); q_.pop_front();
            }
            run(j);
        }
    }
    std::vector<std::thread> th_;
    std::deque<Job> q_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

}  // namespace orbx
