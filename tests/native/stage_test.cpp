// CPU-only unit test of the host staging pool (rgbd_visualodometry_b200/csrc/orbx_stage.h): strided 2-D copies cut into pieces,
// copier threads + the calling thread helping, jobs queued without waking the pool, several counters in flight, clean shutdown
// with work still queued.  Built and run by tests/test_stage.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../rgbd_visualodometry_b200/csrc/orbx_stage.h"

using orbx::HostStager;

static int check(const std::vector<uint8_t>& dst, size_t dpitch, const std::vector<uint8_t>& src, size_t spitch, size_t row, size_t rows, const char* what)
{
    for (size_t r = 0; r < rows; ++r)
        for (size_t x = 0; x < row; ++x)
            if (dst[r * dpitch + x] != src[r * spitch + x]) { printf("FAIL %s: row %zu col %zu\n", what, r, x); return 1; }
    for (size_t r = 0; r < rows; ++r)
        for (size_t x = row; x < dpitch; ++x)
            if (dst[r * dpitch + x] != 0xEE) { printf("FAIL %s: padding of row %zu overwritten\n", what, r); return 1; }
    return 0;
}

int main()
{
    int bad = 0;
    for (int nthreads : {0, 1, 4}) {
        HostStager hs(nthreads);
        // (a) a padded image: 480 rows of 1920 bytes out of a 1999-byte pitch into a 1920 + 16 pitch; several pieces
        const size_t row = 1920, rows = 480, sp = 1999, dp = 1936;
        std::vector<uint8_t> src(sp * rows), dst(dp * rows, 0xEE);
        for (size_t i = 0; i < src.size(); ++i) src[i] = (uint8_t)(i * 2654435761u >> 11);
        std::atomic<int> done(0);
        const int n = hs.submit(dst.data(), dp, src.data(), sp, row, rows, &done, nthreads > 0);
        // ... and the same copy with non-temporal stores into a deliberately misaligned destination
        std::vector<uint8_t> dst_nt(dp * rows + 64, 0xEE);
        std::atomic<int> done_nt(0);
        const int n_nt = hs.submit(dst_nt.data() + 5, dp, src.data() + 3, sp, row, rows, &done_nt, nthreads > 0, true);
        hs.help_until(done_nt, n_nt);
        for (size_t r = 0; r < rows; ++r)
            if (memcmp(dst_nt.data() + 5 + r * dp, src.data() + 3 + r * sp, row) != 0 || dst_nt[5 + r * dp + row] != 0xEE) { printf("FAIL: non-temporal copy, row %zu\n", r); ++bad; break; }
        if (n < 2) { printf("FAIL: expected several pieces, got %d\n", n); ++bad; }
        hs.help_until(done, n);
        bad += check(dst, dp, src, sp, row, rows, "padded image");
        // (b) a contiguous block (pitches == row): one flat copy cut at PIECE boundaries, odd size
        const size_t flat = 3 * HostStager::PIECE + 12345;
        std::vector<uint8_t> a(flat), b(flat, 0);
        for (size_t i = 0; i < flat; ++i) a[i] = (uint8_t)(i * 40503u >> 7);
        std::atomic<int> d2(0);
        const int n2 = hs.submit(b.data(), 1000, a.data(), 1000, 1000, flat / 1000, &d2, false) + hs.submit(b.data() + flat / 1000 * 1000, 0, a.data() + flat / 1000 * 1000, 0, flat % 1000, 1, &d2, false);
        hs.help_until(d2, n2);
        if (a != b) { printf("FAIL: flat copy differs (threads %d)\n", nthreads); ++bad; }
        // (c) many small jobs on several counters at once, queued without waking anybody: the caller alone must finish them
        std::vector<std::vector<uint8_t>> S(64, std::vector<uint8_t>(777)), D(64, std::vector<uint8_t>(777, 0));
        std::atomic<int> c0(0), c1(0);
        int t0 = 0, t1 = 0;
        for (int i = 0; i < 64; ++i) {
            for (auto& v : S[i]) v = (uint8_t)(i + 3);
            (i & 1 ? t1 : t0) += hs.submit(D[i].data(), 0, S[i].data(), 0, 777, 1, i & 1 ? &c1 : &c0, false);
        }
        hs.help_until(c1, t1);
        hs.help_until(c0, t0);
        for (int i = 0; i < 64; ++i) if (S[i] != D[i]) { printf("FAIL: small job %d\n", i); ++bad; }
        // (d) empty submissions are no-ops
        if (hs.submit(D[0].data(), 0, S[0].data(), 0, 0, 1, &c0) != 0 || hs.submit(D[0].data(), 8, S[0].data(), 8, 8, 0, &c0) != 0) { printf("FAIL: empty job produced pieces\n"); ++bad; }
    }
    {   // (e) destruction with queued work that nobody waits for must not hang (workers drain the queue, then stop)
        std::vector<uint8_t> a(1 << 20, 7), b(1 << 20, 0);
        std::atomic<int> d(0);
        int n;
        { HostStager hs(2); n = hs.submit(b.data(), 0, a.data(), 0, a.size(), 1, &d); }
        if (d.load() != n || a != b) { printf("FAIL: jobs lost at shutdown (%d of %d)\n", d.load(), n); ++bad; }
    }
    printf(bad ? "STAGE TEST FAILED\n" : "STAGE TEST OK\n");
    return bad != 0;
}
