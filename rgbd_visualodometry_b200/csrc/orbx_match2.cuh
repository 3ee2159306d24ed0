// orbx_match2.cuh -- the Hamming matcher as a CTA PAIR (tcgen05 cta_group::2): same pipeline as k_hamming_umma
// (orbx_match.cuh), but two CTAs on the SMs of one TPC share every B tile.  Each CTA still owns 256 query rows (two
// 128-row A tiles parked in its own tensor memory) and its own accumulators and epilogue; of a 96-row train tile each
// CTA expands only HALF (48 rows, two threads per row) into its own shared memory, and one UMMA of M = 256 issued by the
// leader feeds both SMs' tensor cores (the hardware exchanges the B halves).  Per SM that halves the bit expansion and the
// shared-memory operand reads of every MMA -- the two things that kept the single-CTA kernel at ~57 instead of 48 cycles
// per MMA.  Barriers: "stage full" and "accumulators free" live in the leader (the peer arrives on them through the
// cluster address space); "stage free" and "accumulators full" are signalled in BOTH CTAs by multicast tcgen05.commit.
// Row code of the bias k-step: local row i of CTA r carries 95 - (48 r + i); for a partial last tile issued with a smaller
// UMMA N the peer holds rows N/2 .. N-1, still in decreasing code order, and the epilogue maps the code back with N.
#pragma once
#include "orbx_match.cuh"

namespace orbx {

constexpr int MT2_HALF = MT_BN / 2;                          // train rows of a B tile held by each CTA of the pair
constexpr int MT2_GROUP_WARPS = MT_BN / 32;                  // expander warps per group (two threads per local row)
// instruction descriptor: as MT_IDESC but M = 256 (cta_group::2)
constexpr uint32_t MT2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MT_BN >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");   // (default .release.cta, as CUTLASS' ClusterBarrier::arrive(cta_id): the cluster-scope form stalls the warp for ~1000 cycles)
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma2_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma2_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

template <bool KNN2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MT_THREADS, 1)
k_hamming_umma2(const uint8_t* __restrict__ query, int nq, const uint8_t* __restrict__ train, int nt, int train_stride_rows,
               const int* __restrict__ train_counts, int nsets, int rows_per_split, int4* __restrict__ best,
               int4* __restrict__ second, int* __restrict__ keys, int* __restrict__ status, int dbg, long long* __restrict__ trace)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + MT_SMEM_BAR;
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * MT_STAGES, bar_tfull = bar0 + 16 * MT_STAGES, bar_tempty = bar_tfull + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + MT_SMEM_BAR + 16 * MT_STAGES + 32);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.y;
    const int q0 = blockIdx.x * MT_QROWS;
    const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the MMAs of the pair), 1 = peer

    // ---- setup: barriers, TMEM, A tiles (expanded in registers and parked in tensor memory)
    if (tid == 0) {
        // full / tempty are only used in the leader (the peer arrives on them remotely); one arrive per warp
        for (int s = 0; s < MT_STAGES; ++s) { mbar_init(bar_full + 8 * s, 2 * MT2_GROUP_WARPS); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MT_ISSUER_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(MT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < (128 * 128 + MT_BN * 128) / 16; i += MT_THREADS) reinterpret_cast<uint4*>(smem + MT_SMEM_CA)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // bias k-step operands: logical byte k = 0 of row r sits at (r / 8) * 1024 + (r % 8) * 128 + ((0 ^ (r % 8)) << 4)
    if (tid < 128) smem[MT_SMEM_CA + (tid >> 3) * 1024 + (tid & 7) * 128 + ((tid & 7) << 4)] = 1;
    else if (tid < 128 + MT2_HALF) {                         // this CTA's half of the bias B tile: local row i <-> code 95 - (48 rank + i)
        const int i = tid - 128;
        smem[MT_SMEM_CB + (i >> 3) * 1024 + (i & 7) * 128 + ((i & 7) << 4)] = (uint8_t)(MT_BN - 1 - ((int)rank * MT2_HALF + i));
    }
    const uint32_t tmem_base = *tmem_slot;
    bool ok = true;
    if (warp >= MT_EPI_WARP0 && warp < MT_EPI_WARP0 + 8) {   // thread <-> query row <-> TMEM lane (same mapping as the epilogue)
        const int q = q0 + ((warp - MT_EPI_WARP0) >> 2) * 128 + (warp & 3) * 32 + lane;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (q < nq) {
            const uint4* p = reinterpret_cast<const uint4*>(query + (size_t)q * 32);
            lo = __ldg(p); hi = __ldg(p + 1);
        }
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + MT_TMEM_A + (uint32_t)((warp - MT_EPI_WARP0) >> 2) * 64u;
#pragma unroll
        for (int half = 0; half < 2; ++half) {               // 64 columns = K 256 int8, four per 32-bit column
            uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = pm127x4((w[(half * 32 + i) >> 3] >> (((half * 32 + i) & 7) * 4)) & 15u);
            tc_st32(ta + half * 32, r);
        }
        tc_wait_st();
    }
    fence_proxy_async_smem();                                // constant bias tiles (generic-proxy stores) -> visible to the tensor core
    tc_fence_before();
    cluster_sync_all();                                      // both CTAs: barriers initialised, TMEM allocated, A tiles parked
    tc_fence_after();
    const uint32_t lead_full = mapa_u32(bar_full, 0), lead_tempty = mapa_u32(bar_tempty, 0);   // the leader's barriers, cluster addresses

    if (warp < MT_EXP_WARPS) {
        // ================= expanders: one B-tile row per thread; raw rows arrive through a cp.async ring
        //                   MT_PREFETCH tiles ahead, so global-load latency never sits on the pipeline's critical path
        // a B tile of N rows is split between the CTAs: rows [0, N/2) live in the leader's shared memory, [N/2, N) in the peer's.
        // Two threads per local row (one 16-byte K half each): 96 threads per group, two groups on alternate tiles.
        const int grp = tid / MT_BN, tl = tid - grp * MT_BN;
        const int r = tl % MT2_HALF, hk = tl / MT2_HALF;      // local row, K half
        struct TileIter {
            int set, i; MatchSetRange rg;
        };
        auto seek = [&](TileIter& it) -> bool {              // make (set, i) name an existing tile, skipping empty sets
            while (it.set < nsets && it.i >= it.rg.ntiles) {
                it.set += gridDim.z; it.i = 0;
                if (it.set < nsets) it.rg = match_set_range(train_counts, it.set, nt, split, rows_per_split);
            }
            return it.set < nsets;
        };
        auto step = [&](TileIter& it) -> bool { if (it.set < nsets) ++it.i; return seek(it); };
        TileIter cur = {(int)blockIdx.z, 0, {0, 0, 0}};
        if (cur.set < nsets) cur.rg = match_set_range(train_counts, cur.set, nt, split, rows_per_split);
        seek(cur);
        for (int k = 0; k < grp; ++k) step(cur);             // group g owns tiles g, g + GROUPS, g + 2 GROUPS, ...
        TileIter ahead = cur;
        uint8_t* raw = smem + MT_SMEM_RAW + (grp * MT_PREFETCH * MT_BN + tl) * 16;
        auto fetch = [&](int slot) {                         // raw half row of the tile `ahead` names -> ring slot (own piece only)
            uint8_t* dst = raw + slot * (MT_BN * 16);
            if (ahead.set < nsets) {
                const int j0t = ahead.rg.n0 + ahead.i * MT_BN;
                const int nh = ((min(MT_BN, ahead.rg.n1 - j0t) + 31) & ~31) >> 1;    // rows per CTA of this tile (UMMA N / 2)
                // rows past the set's end repeat the set's LAST row: a copy scores like the original but carries a smaller
                // index code, so it can never win -- and the epilogue needs no per-column masking
                const int j = min(j0t + (int)rank * nh + r, ahead.rg.n1 - 1);
                if (r < nh) {
                    const uint8_t* src = train + ((size_t)ahead.set * train_stride_rows + j) * 32 + hk * 16;
                    const uint32_t d = smem_u32(dst);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
                } else {
                    reinterpret_cast<uint4*>(dst)[0] = make_uint4(0, 0, 0, 0);
                }
#pragma unroll 1
                for (int k = 0; k < MT_EXP_GROUPS; ++k) step(ahead);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");   // one group per tile slot, empty or not
        };
#pragma unroll 1
        for (int k = 0; k < MT_PREFETCH; ++k) fetch(k);
        int t = grp, n = 0;                                  // global tile number (pipeline stage / phase), own tile count (ring slot)
        while (cur.set < nsets) {
            asm volatile("cp.async.wait_group %0;" ::"n"(MT_PREFETCH - 1) : "memory");
            const int slot = n % MT_PREFETCH;
            const uint4 cv = reinterpret_cast<const uint4*>(raw + slot * (MT_BN * 16))[0];
            fetch(slot);                                     // refill the slot just consumed
            const int s = t % MT_STAGES;
            const uint32_t ph = (uint32_t)(t / MT_STAGES) & 1u;
            const bool tr_on = trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t >= 40 && t < 56 && r == 0;
            if (tr_on) trace[(t - 40) * 16 + 8] = clock64();
            if (!mbar_wait(bar_empty + 8 * s, ph ^ 1u)) { ok = false; break; }
            if (tr_on) trace[(t - 40) * 16 + 9] = clock64();
            if (!(dbg & 4)) expand_half_row(smem + MT_SMEM_B + s * MT_B_BYTES, MT2_HALF, r, hk, cv);
            fence_proxy_async_smem();                        // every writer fences, then one lane arrives for the warp (on the leader's barrier)
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_full + 8 * s);
            if (tr_on) trace[(t - 40) * 16 + 10] = clock64();
#pragma unroll 1
            for (int k = 0; k < MT_EXP_GROUPS; ++k) step(cur);
            t += MT_EXP_GROUPS;
            ++n;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if ((warp == MT_ISSUER_WARP || warp == MT_ISSUER2_WARP) && rank == 0) {
        // ================= MMA issuers =================
        // The whole warp runs the loop so that every operand is computed in warp-uniform code (uniform registers, no
        // per-instruction ELECT / R2UR waterfall); one elected lane issues the tcgen05 instructions.
        // TWO issuer warps take alternate tiles (issuer g <-> tiles t = g mod 2 <-> accumulator buffer g): issuing blocks on
        // the short MMA queue for the whole tensor time of a tile, so a single issuer exposes its per-tile barrier waits,
        // commits and loop latency (~200 cycles of a ~1200-cycle tile period) as tensor-pipe idle time; with two, the next
        // tile's MMAs are already queued behind the current one's.
        const int issuer = warp == MT_ISSUER_WARP ? 0 : 1;
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t sbu = __shfl_sync(0xffffffffu, sbase, 0);
        const uint64_t dca = umma_desc_sw128(sbu + MT_SMEM_CA), dcb = umma_desc_sw128(sbu + MT_SMEM_CB);
        int t = 0;
        bool ready_next = false;                             // barriers of tile t already observed complete (probed mid-tile)
        // (the row count of the NEXT set is loaded a whole set ahead: its global-load latency would otherwise stall the
        //  pipeline at every set boundary)
        int cnt_next = match_set_count(train_counts, min((int)blockIdx.z, nsets - 1), nt);
        for (int set = blockIdx.z; set < nsets && ok; set += gridDim.z) {
            const MatchSetRange rg = match_set_range_n(cnt_next, nt, split, rows_per_split);
            if (set + (int)gridDim.z < nsets) cnt_next = match_set_count(train_counts, set + (int)gridDim.z, nt);   // consumed at the next set boundary
            for (int i = 0; i < rg.ntiles; ++i, ++t) {
                if ((t & 1) != issuer) continue;              // the other issuer's tile
                const int s = t % MT_STAGES;
                const uint32_t ph = (uint32_t)(t / MT_STAGES) & 1u;
                const int b = t & 1;
                const uint32_t bph = (uint32_t)(t >> 1) & 1u;
                const bool tr_on = trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t >= 40 && t < 56 && lane == 0;
                if (tr_on) trace[(t - 40) * 16 + 0] = clock64();
                if (!ready_next) {
                    if (!mbar_wait(bar_tempty + 8 * b, bph ^ 1u)) { ok = false; break; }
                    if (tr_on) trace[(t - 40) * 16 + 1] = clock64();
                    if (!mbar_wait(bar_full + 8 * s, ph)) { ok = false; break; }
                }
                if (tr_on) trace[(t - 40) * 16 + 2] = clock64();
                tc_fence_after();
                const uint32_t sb = sbu + MT_SMEM_B + s * MT_B_BYTES;
                const uint64_t db0 = umma_desc_sw128(sb);
                // A set's last tile usually holds fewer than MT_BN rows: issue it with the smallest UMMA N (multiple of 16)
                // that covers them -- tensor time is proportional to N.  Columns beyond keep stale values the epilogue
                // never reads (it bounds partial tiles by the row count).
                const int rows_here = min(MT_BN, rg.n1 - (rg.n0 + i * MT_BN));
                const uint32_t idesc = (MT2_IDESC & ~(0x3Fu << 17)) | ((uint32_t)(((rows_here + 31) & ~31) >> 3) << 17);
                // barriers of this issuer's NEXT tile, t + 2 (same pipelines, consecutive tile numbers even across set boundaries)
                const int s1 = (t + 2) % MT_STAGES, b1 = b;
                const uint32_t ph1 = (uint32_t)((t + 2) / MT_STAGES) & 1u, bph1 = (uint32_t)((t + 2) >> 1) & 1u;
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    if (elect_one() && !(dbg & 2)) {
                        const uint32_t d = tb + (uint32_t)((b * 2 + a) * MT_BN);
                        const uint32_t ta = tb + MT_TMEM_A + (uint32_t)a * 64u;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {      // K = 256 = 8 x UMMA_K(32 int8): 8 TMEM columns of A, 4 k-steps per 128-B swizzle atom of B
                            const uint64_t db = db0 + (uint64_t)(((ks >> 2) * (MT2_HALF * 128) + (ks & 3) * 32) >> 4);   // start-address field, 16-byte units
                            tc_mma2_i8_ts(d, ta + (uint32_t)ks * 8u, db, idesc, ks > 0 ? 1u : 0u);
                        }
                        tc_mma2_i8(d, dca, dcb, idesc, 1u); // bias k-step: + row code in every row
                    }
                    // While the queued MMAs execute, probe the next tile's barriers so its issue can start without a wait.
                    // The probe sits AFTER the last MMA of the tile: issuing blocks on the MMA queue for most of the tile's
                    // tensor time, so by now the epilogue of tile t - 1 has usually handed its accumulators back.
                    if (a == 1) ready_next = mbar_test(bar_tempty + 8 * b1, bph1 ^ 1u) && mbar_test(bar_full + 8 * s1, ph1);
                }
                if (elect_one()) {
                    tc_commit2(bar_empty + 8 * s);           // smem stage reusable once these MMAs retire (both CTAs' barriers)
                    tc_commit2(bar_tfull + 8 * b);           // accumulators of this tile complete (both CTAs' barriers)
                }
                __syncwarp();
                if (tr_on) trace[(t - 40) * 16 + 3] = clock64();
            }
        }
    } else if (warp >= MT_EPI_WARP0 && warp < MT_EPI_WARP0 + 8) {
        // ================= epilogue =================
        const int a = (warp - MT_EPI_WARP0) >> 2;            // A tile; the TMEM lane quarter is fixed by hardware to warp % 4
        const int qrow = q0 + a * 128 + (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        int t = 0;
        int cnt_next = match_set_count(train_counts, min((int)blockIdx.z, nsets - 1), nt);
        for (int set = blockIdx.z; set < nsets && ok; set += gridDim.z) {
            const MatchSetRange rg = match_set_range_n(cnt_next, nt, split, rows_per_split);   // (next set's row count loaded a set ahead, as in the issuer)
            if (set + (int)gridDim.z < nsets) cnt_next = match_set_count(train_counts, set + (int)gridDim.z, nt);
            int m1 = INT_MIN, m2 = INT_MIN;
            for (int i = 0; i < rg.ntiles; ++i, ++t) {
                const int b = t & 1;
                const uint32_t bph = (uint32_t)(t >> 1) & 1u;
                const bool tr_on = trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t >= 40 && t < 56 && warp == MT_EPI_WARP0 + 1 && lane == 0;
                if (tr_on) trace[(t - 40) * 16 + 4] = clock64();
                if (!mbar_wait(bar_tfull + 8 * b, bph)) { ok = false; break; }
                if (tr_on) trace[(t - 40) * 16 + 5] = clock64();
                tc_fence_after();
                const int j0 = rg.n0 + i * MT_BN;
                const bool full = (j0 + MT_BN <= rg.n1);
                constexpr int NCH = MT_BN / 32;
                uint32_t r[NCH][32];
                if (!(dbg & 1)) {
                    const uint32_t tacc = tmem_base + lane_base + (uint32_t)((b * 2 + a) * MT_BN);
                    int k1 = INT_MIN, k2 = INT_MIN;
                    if (!KNN2 && full) {
                        // accumulator = 127 * dot + code lies in [-32512, 32607]: it fits in 16 bits, so the tile comes out of
                        // tensor memory packed two columns per register (half the registers, half the max operations), and every
                        // value of a row is distinct (distinct codes), so the packed maximum loses nothing
                        uint32_t rp[MT_BN / 2];
                        tc_ld32_pack16(tacc, rp);
                        tc_ld16_pack16(tacc + 64, rp + 32);
                        tc_wait_ld();
                        if (tr_on) trace[(t - 40) * 16 + 6] = clock64();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * b);           // TMEM buffer back to the issuer BEFORE the reduction
                        unsigned p[8];                           // eight independent max chains (ILP), then a short tree
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            p[u] = __vimax3_s16x2(rp[u], rp[8 + u], rp[16 + u]);
                            p[u] = __vimax3_s16x2(p[u], rp[24 + u], rp[32 + u]);
                            p[u] = __vmaxs2(p[u], rp[40 + u]);
                        }
                        const unsigned pk = __vimax3_s16x2(__vimax3_s16x2(p[0], p[1], p[2]), __vimax3_s16x2(p[3], p[4], p[5]), __vmaxs2(p[6], p[7]));
                        k1 = max((int)(pk << 16) >> 16, (int)pk >> 16);
                    } else if (!KNN2) {
                        // a set's partial last tile was issued with N = 32, 64 (or 96) columns, its rows past the set's end are
                        // copies of the last row (never the maximum): same packed path, 32-column chunks as far as N goes
                        const int nch = (rg.n1 - j0 + 31) >> 5;
                        uint32_t rp[MT_BN / 2];
#pragma unroll
                        for (int ch = 0; ch < NCH; ++ch) {
                            if (ch < nch) tc_ld16_pack16(tacc + (uint32_t)(ch * 32), rp + ch * 16);
                            else {
#pragma unroll
                                for (int u = 0; u < 16; ++u) rp[ch * 16 + u] = 0x80008000u;
                            }
                        }
                        tc_wait_ld();
                        if (tr_on) trace[(t - 40) * 16 + 6] = clock64();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * b);
                        unsigned p[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            p[u] = __vimax3_s16x2(rp[u], rp[8 + u], rp[16 + u]);
                            p[u] = __vimax3_s16x2(p[u], rp[24 + u], rp[32 + u]);
                            p[u] = __vmaxs2(p[u], rp[40 + u]);
                        }
                        const unsigned pk = __vimax3_s16x2(__vimax3_s16x2(p[0], p[1], p[2]), __vimax3_s16x2(p[3], p[4], p[5]), __vmaxs2(p[6], p[7]));
                        k1 = max((int)(pk << 16) >> 16, (int)pk >> 16);
                    } else {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) tc_ld32(tacc + (uint32_t)(ch * 32), r[ch]);
                    tc_wait_ld();
                    if (tr_on) trace[(t - 40) * 16 + 6] = clock64();
                    // The accumulators now live in registers: hand the TMEM buffer back to the issuer BEFORE the max
                    // reduction, so the epilogue's arithmetic overlaps the MMAs of tile t + 2 instead of gating them.
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * b);
#pragma unroll
                        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                            for (int c = 0; c < 32; ++c)
                                if (full || j0 + ch * 32 + c < rg.n1) {
                                    const int k = (int)r[ch][c];
                                    if (KNN2) k2 = max(k2, min(k1, k));
                                    k1 = max(k1, k);
                                }
                    }
                    // decode the tile winner(s) into the global key  dot << 20 | (0xFFFFF - j)
#pragma unroll
                    for (int w = 0; w < (KNN2 ? 2 : 1); ++w) {
                        const int k = w == 0 ? k1 : k2;
                        if (k != INT_MIN) {
                            const unsigned kk = (unsigned)(k + MT_ASCALE * 256);
                            const unsigned q = kk / (unsigned)MT_ASCALE;          // dot + 256
                            const int code = (MT_BN - 1) - (int)(kk - q * (unsigned)MT_ASCALE);   // 48 * (CTA holding the row) + local row
                            const int nh = ((min(MT_BN, rg.n1 - j0) + 31) & ~31) >> 1;
                            const int jl = code >= MT2_HALF ? nh + code - MT2_HALF : code;
                            const int gk = ((int)q - 256) * (1 << MT_KEY_SHIFT) + ((MT_MAX_TRAIN - 1) - (j0 + jl));
                            if (KNN2) m2 = max(m2, min(m1, gk));
                            m1 = max(m1, gk);
                        }
                    }
                }
                else {                                       // (perf-experiment mode without TMEM loads: still release the buffer)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * b);
                }
                if (tr_on) trace[(t - 40) * 16 + 11] = clock64() + (m1 & 1);
                if (tr_on) trace[(t - 40) * 16 + 7] = clock64();
            }
            const bool tr_set = trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t > 40 && t <= 56 && warp == MT_EPI_WARP0 + 1 && lane == 0;
            if (tr_set) trace[(t - 1 - 40) * 16 + 12] = clock64();
            if (ok && qrow < nq) {
                const size_t o = (size_t)set * nq + qrow;
                if (keys) {
                    if (m1 != INT_MIN) atomicMax(&keys[o], m1);
                } else if (m1 == INT_MIN) {                   // empty train set: no match (trainIdx -1)
                    best[o] = make_int4(qrow, -1, 0, 0);
                    if (KNN2) second[o] = make_int4(qrow, -1, 0, 0);
                } else {
                    const int dot = m1 >> MT_KEY_SHIFT, j = (MT_MAX_TRAIN - 1) - (m1 & (MT_MAX_TRAIN - 1));
                    best[o] = make_int4(qrow, j, 0, __float_as_int((float)((256 - dot) >> 1)));
                    if (KNN2) {
                        if (m2 != INT_MIN) {
                            const int dot2 = m2 >> MT_KEY_SHIFT, j2 = (MT_MAX_TRAIN - 1) - (m2 & (MT_MAX_TRAIN - 1));
                            second[o] = make_int4(qrow, j2, 0, __float_as_int((float)((256 - dot2) >> 1)));
                        } else {
                            second[o] = make_int4(qrow, -1, 0, 0);
                        }
                    }
                }
            }
            if (tr_set) trace[(t - 1 - 40) * 16 + 13] = clock64();
        }
    }
    if (!ok) atomicOr(status, 2);
    tc_fence_before();
    cluster_sync_all();                                      // neither CTA may free tensor memory or exit while the pair still works
    if (warp == MT_ISSUER_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(MT_TMEM_COLS) : "memory");
    }
}

}  // namespace orbx
