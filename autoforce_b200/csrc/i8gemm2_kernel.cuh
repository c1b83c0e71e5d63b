// CTA-pair variant of the int8-sliced tcgen05 GEMM (i8gemm_kernel.cuh): `tcgen05.mma.cta_group::2`.
//
// Why: the single-CTA kernel is bound by the bytes every SM pulls from L2 into shared memory (DESIGN.md section 4.1:
// 72 KB per K-chunk for 42 MMAs, ~22 B/clk/SM).  Two CTAs of a cluster (the two SMs of a TPC) compute ONE 256 x 64 tile:
// each CTA stages the 128 A rows whose accumulators live in its own tensor memory, but only HALF of the B chunk
// (32 of the 64 rows) -- the MMA reads the other half from the peer's shared memory.  Per CTA and K-chunk:
// 48 KB (A) + 12 KB (B/2) = 60 KB instead of 72 KB for the same number of MMAs: -17 % operand feed.
//
// Roles per CTA (320 threads): warp 0 = TMA producer (both CTAs; every load signals the LEADER's full barrier),
// warp 1 = MMA issuer (leader CTA only; one elected lane issues the cta_group::2 MMAs and multicasts the commits to
// both CTAs' barriers), warps 2..9 = epilogue (each CTA drains its own 128 TMEM lanes; the peer's epilogue warps arrive
// on the leader's tmem_empty barrier through shared::cluster).
#pragma once
#include "i8gemm_kernel.cuh"

namespace sgpr {
namespace i8g {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even CTA of the pair
constexpr int BNH = BN / 2;                      // B rows staged per CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// executed by both CTAs; the transaction bytes are credited to the leader CTA's barrier
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
// arrive on the barrier at the same offset in BOTH CTAs once all MMAs issued so far have retired
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void mma_i8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct Problem2 {
    CUtensorMap mapA;     // as Problem::mapA (box 128 rows)
    CUtensorMap mapBh;    // B operand with a box of BN/2 = 32 rows
    int N, Kpad;
    const double* aux;    // as Problem::aux
};

template <int NS, int TR>
struct Scheme2 {
    static constexpr int NG = TR - 1;
    static constexpr int A_BYTES = NS * BM * BKB, B_BYTES = NS * BNH * BKB, STAGE_BYTES = A_BYTES + B_BYTES;
    static_assert(NG * BN <= 512, "accumulators exceed TMEM");
};

template <int NS, int TR, int STAGES, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
    i8gemm2_kernel(const Common* __restrict__ cmp, const Problem2* __restrict__ probs, Epi epi) {
    using SC = Scheme2<NS, TR>;
    __shared__ Common cm;
    __shared__ int tstart[9];      // first 256-row tile of each problem
    if (threadIdx.x < sizeof(Common) / 4) reinterpret_cast<int*>(&cm)[threadIdx.x] = reinterpret_cast<const int*>(cmp)[threadIdx.x];
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + (size_t)STAGES * SC::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    constexpr uint32_t tmem_cols = 512;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);        // the leader's producer arrives (with the bytes of BOTH CTAs expected)
            mbar_init(&empty_bar[s], 1);       // multicast commit of the leader's MMA thread
        }
        mbar_init(tmem_full, 1);               // multicast commit
        mbar_init(tmem_empty, 16);             // 8 epilogue warps of each CTA (used in the leader only)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc2(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int p = 0; p < 8; ++p) {
            tstart[p] = t;
            if (p < cm.n_prob) t += ((cm.M[p] + 2 * BM - 1) / (2 * BM)) * ((probs[p].N + BN - 1) / BN);
        }
        tstart[8] = t;
        for (int p = cm.n_prob; p < 8; ++p) tstart[p] = t;
    }
    __syncthreads();
    cluster_sync_all();                        // both CTAs' barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = tstart[8];

    auto locate = [&](int gt, int& pi, int& tm, int& tn) {
        pi = 0;
        for (int q = 1; q < cm.n_prob; ++q)
            if (gt >= tstart[q]) pi = q;
        const int tile = gt - tstart[pi];
        const int tiles_n = (probs[pi].N + BN - 1) / BN;
        tm = tile / tiles_n;
        tn = tile - tm * tiles_n;
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int gt = pair; gt < n_tiles; gt += n_pairs) {
            int pi, tm, tn;
            locate(gt, pi, tm, tn);
            const Problem2& P = probs[pi];
            const int nk = P.Kpad / BKB;
            for (int kt = 0; kt < nk; ++kt) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sa = smem + (size_t)stage * SC::STAGE_BYTES;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * SC::STAGE_BYTES);
                    tma_load_4d_2sm(sa, &P.mapA, &full_bar[stage], 0, cm.row0[pi] + tm * 2 * BM + (int)rank * BM, kt, 0);
                    tma_load_4d_2sm(sa + SC::A_BYTES, &P.mapBh, &full_bar[stage], 0, tn * BN + (int)rank * BNH, kt, 0);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            // instruction descriptor: M = 256 (the pair), N = 64
            const uint32_t idesc = make_idesc_i8(2 * BM, BN);
            const uint64_t desc_hi = make_desc_sw64(0);
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int gt = pair; gt < n_tiles; gt += n_pairs) {
                int pi, tm, tn;
                locate(gt, pi, tm, tn);
                const int nk = probs[pi].Kpad / BKB;
                mbar_wait(tmem_empty, tphase ^ 1);      // both CTAs' epilogues have drained the accumulators
                tc_fence_after();
                for (int kt = 0; kt < nk; ++kt) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * SC::STAGE_BYTES);
                    const uint64_t adesc = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF);
                    const uint64_t bdesc = desc_hi | (uint64_t)(((sa + SC::A_BYTES) >> 4) & 0x3FFF);
                    const uint32_t later = kt > 0 ? 1u : 0u;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < BKB / 32; ++ks) {
#pragma unroll
                            for (int t = 1; t <= NS; ++t) {
#pragma unroll
                                for (int u = 1; u <= NS; ++u) {
                                    if (t + u <= TR) {
                                        const uint32_t accum = (ks == 0 && (t == 1 || u == NS)) ? later : 1u;
                                        mma_i8_2cta(tmem_base + (t + u - 2) * BN, adesc + (((t - 1) * (BM * BKB) + ks * 32) >> 4),
                                                    bdesc + (((u - 1) * (BNH * BKB) + ks * 32) >> 4), idesc, accum);
                                    }
                                }
                            }
                        }
                        tc_commit2(&empty_bar[stage]);              // both CTAs' stage is free once these MMAs retire
                        if (kt == nk - 1) tc_commit2(tmem_full);    // both CTAs' accumulators complete
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tphase ^= 1;
            }
        }
    } else {
        // ===================== epilogue (both CTAs): this CTA's 128 rows of the 256-row tile =====================
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row_in_tile = q * 32 + lane;
        constexpr int HC = BN / 2;
        uint32_t tphase = 0;
        for (int gt = pair; gt < n_tiles; gt += n_pairs) {
            int pi, tm, tn;
            locate(gt, pi, tm, tn);
            double v[HC];
            mbar_wait(tmem_full, tphase);
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + half * HC;
#pragma unroll
            for (int cc = 0; cc < HC; cc += 8) combine8<SC::NG>(lane_addr + cc, v + cc);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0)
                    mbar_arrive(tmem_empty);
                else
                    mbar_arrive_leader(tmem_empty);
            }
            tphase ^= 1;
            const int row = tm * 2 * BM + (int)rank * BM + row_in_tile;
#pragma unroll
            for (int cc = 0; cc < HC; cc += 16) epi(pi, cm.row0[pi], row, tn * BN + half * HC + cc, v + cc, cm.M[pi], probs[pi].N, probs[pi].aux);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                        // the peer may still read this CTA's shared memory / signal its barriers
    if (warp == 1) tmem_dealloc2(tmem_base, tmem_cols);
}

template <int NS, int STAGES>
constexpr size_t smem_bytes2() { return (size_t)STAGES * NS * (BM + BNH) * BKB + 1024 + 256; }

}  // namespace i8g
}  // namespace sgpr
