// EXPERIMENT (not on the product path): the tcgen05 int8-sliced GEMM with the A operand in TENSOR MEMORY.
//
// Question: i8gemm_kernel's tensor pipe is ~50 % active at the kernel-matrix shape; is it limited by re-reading
// the A slices from shared memory for every slice product?  Here four loader warps bring every A chunk ONCE
// global -(cp.async)-> shared -> registers -(tcgen05.st)-> TMEM (one row per lane, 32 int8 of K per 8 columns;
// the layout was confirmed numerically: results identical to i8gemm_kernel), the MMAs take A from TMEM
// (tcgen05.mma ... [d_tmem], [a_tmem], b_desc, ...) and only B from shared memory.
//
// Measured on B200, M x N x K = 97556 x 500 x 544, 21 slice products (tools/i8gemm_test.cu):
//   i8gemm_kernel (A and B by TMA, both from shared memory)        0.64 ms  1850 int8 TOP/s
//   this kernel, A by cp.async (or by TMA) -> TMEM                 0.60-0.63 ms
//   this kernel, A straight from global through registers          0.88 ms  (latency-bound)
//   this kernel, A preloaded (no A traffic at all, wrong results)  0.40 ms  2970 TOP/s = 90 % of the int8 peak
// i.e. the MMA side runs at 90 % of peak once A sits in TMEM, but bringing the 48 KB A chunk + 24 KB B chunk per
// 42 MMAs into the SM costs the same ~3.1k cycles whichever path it takes, and time scales 1/grid (a per-SM
// ingest limit of ~22 B/clk, not L2 bandwidth, not shared-memory bandwidth).  Raising the tile's arithmetic
// intensity needs N > 64, which the 512 TMEM columns do not allow with 6 accumulator groups.  Kept for the record.
#pragma once
#include "i8gemm_kernel.cuh"

namespace sgpr {
namespace i8g {

struct ProblemTA {
    CUtensorMap mapA;      // unused here
    CUtensorMap mapB;      // int8 [NS][rowsB][Kpad], box {64, 64, NS}
    int M, N, Kpad;
    const signed char* A;  // slice 0, row 0 of the A digits; row r of slice t at A + t*a_slice + r*a_ld
    long long a_slice;
    int a_ld;
};

constexpr int NTHREADS_TA = 448;

__device__ __forceinline__ void mma_i8_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint4 lo, uint4 hi) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(lo.x),
                 "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// STAGES = TMA stages of B (24 KB each); the A chunks travel global -> shared by cp.async, issued by the loader
// warp that later consumes them (ADEPTH chunks in flight, 12 KB per warp and chunk, private to the warp: no
// barrier), then ld.shared -> tcgen05.st -> TMEM.
template <int NS, int TR, int STAGES, class Epi, int ADEPTH = 3>
__global__ void __launch_bounds__(NTHREADS_TA, 1) i8gemm_ta_kernel(const __grid_constant__ Common cm,
                                                                   const ProblemTA* __restrict__ probs, Epi epi) {
    using SC = Scheme<NS, TR>;
    constexpr int B_BYTES = SC::B_BYTES;
    constexpr int A_WARP_BYTES = NS * 32 * BKB;           // one warp's 32 rows of one chunk, all slices
    constexpr int ACOL0 = SC::NG * BN, AHALF = NS * 8;
    static_assert(ACOL0 + 2 * AHALF <= 512, "accumulators + A buffers exceed TMEM");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem + (size_t)STAGES * B_BYTES;    // [4 warps][ADEPTH][NS][32 rows][64 B], 16-byte chunks swizzled
    uint64_t* full_bar = (uint64_t*)(smem_a + (size_t)4 * ADEPTH * A_WARP_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* a_full = empty_bar + STAGES;     // [2]
    uint64_t* a_empty = a_full + 2;            // [2]
    uint64_t* tmem_full = a_empty + 2;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t tmem_cols = 512;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int h = 0; h < 2; ++h) {
            mbar_init(&a_full[h], 4);            // one arrive per loader warp, after its tcgen05.st completed
            mbar_init(&a_empty[h], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 8);                // one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = cm.tile_start[cm.n_prob];

    if (warp == 0) {
        // ===================== TMA producer: B =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int gt = blockIdx.x; gt < n_tiles; gt += gridDim.x) {
            int pi = 0;
            for (int q = 1; q < cm.n_prob; ++q)
                if (gt >= cm.tile_start[q]) pi = q;
            const ProblemTA& P = probs[pi];
            const int tile = gt - cm.tile_start[pi];
            const int tiles_n = (P.N + BN - 1) / BN;
            const int tn = tile % tiles_n;
            const int nk = P.Kpad / BKB;
            for (int kt = 0; kt < nk; ++kt) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full_bar[stage], B_BYTES);
                    tma_load_3d(smem + (size_t)stage * B_BYTES, &P.mapB, &full_bar[stage], kt * BKB, tn * BN, 0);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_i8(BM, BN);
        const uint64_t desc_hi = make_desc_sw64(0);
        int stage = 0;
        uint32_t phase = 0, tphase = 0, aphase = 0;
        for (int gt = blockIdx.x; gt < n_tiles; gt += gridDim.x) {
            int pi = 0;
            for (int q = 1; q < cm.n_prob; ++q)
                if (gt >= cm.tile_start[q]) pi = q;
            const int nk = probs[pi].Kpad / BKB;
            mbar_wait(tmem_empty, tphase ^ 1);
            tc_fence_after();
            for (int kt = 0; kt < nk; ++kt) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sb = smem_u32(smem + (size_t)stage * B_BYTES);
                const uint64_t bdesc = desc_hi | (uint64_t)((sb >> 4) & 0x3FFF);
                const uint32_t later = kt > 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    mbar_wait(&a_full[ks], aphase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int t = 1; t <= NS; ++t) {
#pragma unroll
                            for (int u = 1; u <= NS; ++u) {
                                if (t + u <= TR) {
                                    const uint32_t accum = (ks == 0 && (t == 1 || u == NS)) ? later : 1u;
                                    mma_i8_ta(tmem_base + (t + u - 2) * BN, tmem_base + ACOL0 + ks * AHALF + (t - 1) * 8,
                                              bdesc + (((u - 1) * (BN * BKB) + ks * 32) >> 4), idesc, accum);
                                }
                            }
                        }
                        tc_commit(&a_empty[ks]);                      // this half of the A buffer may be overwritten
                        if (ks == 1) {
                            tc_commit(&empty_bar[stage]);             // B stage free once these MMAs retire
                            if (kt == nk - 1) tc_commit(tmem_full);   // accumulators of this tile complete
                        }
                    }
                    __syncwarp();
                }
                aphase ^= 1;
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            tphase ^= 1;
        }
    } else if (warp >= 10) {
        // ===================== A loaders: global -(cp.async)-> shared -> registers -> TMEM =====================
        const int q = warp & 3;                                   // TMEM lane quarter = rows 32q .. 32q+31 of the tile
        uint8_t* abuf = smem_a + (size_t)q * ADEPTH * A_WARP_BYTES;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + ACOL0;
        const int sw = (lane >> 1) & 3;                           // consumer: 16-byte chunk c of row `lane` at c ^ sw
        const int cl = lane >> 2, cc = lane & 3;                  // copier: row cl (+8 per step), chunk cc of 8 rows x 64 B
        struct Cursor {
            int gt, kt, nk, row0, M, a_ld;
            long long a_slice;
            const signed char* A;
        };
        auto open_tile = [&](Cursor& c) {
            int pi = 0;
            for (int qq = 1; qq < cm.n_prob; ++qq)
                if (c.gt >= cm.tile_start[qq]) pi = qq;
            const ProblemTA& P = probs[pi];
            const int tile = c.gt - cm.tile_start[pi];
            const int tiles_n = (P.N + BN - 1) / BN;
            c.row0 = (tile / tiles_n) * BM + q * 32;
            c.nk = P.Kpad / BKB;
            c.M = P.M;
            c.A = P.A;
            c.a_slice = P.a_slice;
            c.a_ld = P.a_ld;
        };
        auto advance = [&](Cursor& c) -> bool {   // next chunk; false when the CTA's tiles are exhausted
            if (++c.kt < c.nk) return true;
            c.kt = 0;
            c.gt += gridDim.x;
            if (c.gt >= n_tiles) return false;
            open_tile(c);
            return true;
        };
        auto issue = [&](const Cursor& c, int slot) {
            uint8_t* dst0 = abuf + (size_t)slot * A_WARP_BYTES;
#pragma unroll
            for (int i = 0; i < NS * 4; ++i) {
                const int t = i >> 2, l = (i & 3) * 8 + cl;
                const int row = c.row0 + l;
                const bool ok = row < c.M;
                const signed char* src = c.A + t * c.a_slice + (long long)(ok ? row : 0) * c.a_ld + c.kt * BKB + cc * 16;
                cp_async16_zfill(dst0 + (t * 32 + l) * BKB + ((cc ^ ((l >> 1) & 3)) << 4), src, ok);
            }
        };
        Cursor ci{}, cc_{};
        ci.gt = cc_.gt = blockIdx.x;
        bool more_issue = ci.gt < n_tiles, more = more_issue;
        if (more) {
            open_tile(ci);
            open_tile(cc_);
        }
        int islot = 0, cslot = 0;
        for (int d = 0; d < ADEPTH - 1; ++d) {   // prologue: ADEPTH-1 chunks in flight
            if (more_issue) {
                issue(ci, islot);
                more_issue = advance(ci);
            }
            cp_async_commit();
            if (++islot == ADEPTH) islot = 0;
        }
        uint32_t aphase = 0;
        while (more) {
            if (more_issue) {
                issue(ci, islot);
                more_issue = advance(ci);
            }
            cp_async_commit();
            if (++islot == ADEPTH) islot = 0;
            cp_async_wait<ADEPTH - 1>();          // the chunk to consume has landed (this thread's copies)
            __syncwarp();                         // ... and every lane's
            const uint8_t* sa = abuf + (size_t)cslot * A_WARP_BYTES + lane * BKB;
            uint4 v[NS][4];
#pragma unroll
            for (int t = 0; t < NS; ++t)
#pragma unroll
                for (int c = 0; c < 4; ++c) v[t][c] = *reinterpret_cast<const uint4*>(sa + t * (32 * BKB) + ((c ^ sw) << 4));
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                mbar_wait(&a_empty[ks], aphase ^ 1);
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < NS; ++t) tmem_st8(lane_addr + ks * AHALF + t * 8, v[t][2 * ks], v[t][2 * ks + 1]);
                tmem_st_wait();                   // warp-collective: all 32 rows of this warp are in TMEM
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[ks]);
            }
            aphase ^= 1;
            if (++cslot == ADEPTH) cslot = 0;
            more = advance(cc_);
        }
        cp_async_wait<0>();
    } else {
        // ===================== epilogue: 8 warps, 32 lanes x 32 columns each =====================
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row_in_tile = q * 32 + lane;
        uint32_t tphase = 0;
        for (int gt = blockIdx.x; gt < n_tiles; gt += gridDim.x) {
            int pi = 0;
            for (int qq = 1; qq < cm.n_prob; ++qq)
                if (gt >= cm.tile_start[qq]) pi = qq;
            const ProblemTA& P = probs[pi];
            const int tile = gt - cm.tile_start[pi];
            const int tiles_n = (P.N + BN - 1) / BN;
            const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
            mbar_wait(tmem_full, tphase);
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + half * 32;
            double v[32];
#pragma unroll
            for (int cc = 0; cc < 32; cc += 8) combine8<SC::NG>(lane_addr + cc, v + cc);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            tphase ^= 1;
#pragma unroll
            for (int cc = 0; cc < 32; cc += 16) epi(pi, tm * BM + row_in_tile, tn * BN + half * 32 + cc, v + cc, P.M, P.N);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

template <int NS, int STAGES, int ADEPTH = 3>
constexpr size_t smem_bytes_ta() { return (size_t)STAGES * NS * BN * BKB + (size_t)ADEPTH * NS * BM * BKB + 1024 + 256; }

}  // namespace i8g
}  // namespace sgpr
