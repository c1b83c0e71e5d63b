// FP64-accurate GEMM on the 5th-generation tensor cores (tcgen05, sm_100a) by integer slicing.
//
//   C[M x N] = A[M x K] . B[N x K]^T ,   |A|,|B| <= 1 (rows are normalised descriptors / k^(xi-1))
//
// Each operand is held as NS balanced signed digits in base 256 ("Ozaki scheme"):
//   x ~= 2^-(8 NS - 2) * sum_t a_t 256^(NS - t),   a_t in [-128, 127]   (int8; one bit of headroom)
// so that  A.B^T = 2^-2(8NS-2) * sum_{t,u} 256^(2NS - t - u)  (a_t . b_u^T),  every slice product an
// int8 x int8 -> int32 GEMM that tcgen05.mma kind::i8 evaluates EXACTLY (|sum| <= K 128^2 pairs < 2^31
// for K <= 16384).  Pairs with t + u > TR are dropped (they are below the rounding of the retained
// ones); all pairs with the same t + u share one TMEM accumulator, so the epilogue converts TR - 1
// int32 tiles to float64.  For unit-norm rows: NS = 6, TR = 7 (21 slice products) agrees with the
// float64 product to ~1e-12 absolute, NS = 6, TR = 8 (26 products) to ~3e-14.
//
// Kernel structure (one CTA per SM, persistent over tiles): warp 0 = TMA producer (one 3-D bulk
// tensor copy per operand per stage brings all NS slices of a 128 x 64 B / 64 x 64 B K-chunk,
// SWIZZLE_64B), warp 1 = single-thread tcgen05.mma issuer (2 k-steps x pairs per stage, accumulators
// in TMEM: (TR-1) x 64 columns), warps 2-9 = epilogue (tcgen05.ld -> float64 -> fused epilogue; each warp
// owns 32 TMEM lanes x 32 columns).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgpr {
namespace i8g {

constexpr int BM = 128, BN = 64, BKB = 64;     // tile; BKB = K bytes (= int8 elements) per stage
constexpr int MAXS = 7;                        // max slices
constexpr int NTHREADS = 320;                  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (2 per TMEM lane quarter)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one elected lane of a converged warp (keeps the surrounding code warp-uniform, so descriptors and
// TMEM addresses stay in uniform registers and UTCIMMA issues without per-instruction waterfall loops)
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t"
        "}"
        : "=r"(pred));
    return pred;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns of TMEM -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_64B, rows of 64 B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;               // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                        // SWIZZLE_64B
    return d;
}
// instruction descriptor: int8 x int8 -> int32, both K-major, M = 128, N = BN
__device__ __forceinline__ uint32_t make_idesc_i8(int M, int N) {
    uint32_t d = 0;
    d |= 2u << 4;                  // c_format = S32
    d |= 1u << 7;                  // a_format = signed int8
    d |= 1u << 10;                 // b_format = signed int8
    d |= (uint32_t)(N >> 3) << 17; // n_dim
    d |= (uint32_t)(M >> 4) << 24; // m_dim
    return d;
}

struct Problem {
    CUtensorMap mapA;    // int8 [NS][Kpad/64][rowsA][64], box {64, 128, 1, NS}; covers the WHOLE operand buffer (all species' rows):
                         // the problem's first row comes from Common::row0
    CUtensorMap mapB;    // int8 [NS][Kpad/64][rowsB][64], box {64,  64, 1, NS}
    int M, N, Kpad;      // N columns; Kpad multiple of 64 (zero padded); M unused (Common::M)
    const double* aux;   // per-problem vector handed to the epilogue functor (mu block / choli row scales); may be null
    const int* nk_tn;    // optional [ceil(N/64)]: K chunks (of 64) that can be non-zero for column tile tn; nullptr =
                         // Kpad / 64 for every tile.  Lets a triangular B (choli) skip its zero part; 0 = the tile
                         // is identically zero: no loads, no MMAs, the epilogue runs on zeros.
};

// Work list of one grouped launch.  It lives in DEVICE memory and is written by a one-warp set-up kernel from the
// species row ranges the neighbour stage left on the device (i8gemm.cu: i8_setup_kernel), so the host never needs to
// know how many rows each central species has this step: no device-to-host copy, no synchronisation, and the launch
// sequence is identical from step to step (CUDA-graph capturable).
struct Common {
    int n_prob;
    int tile_start[9];   // first tile of each problem (problem-major)
    int row0[8];         // first row of problem p in the row-major operand / output buffers (species block)
    int M[8];            // rows of problem p
};

template <int NS, int TR>
struct Scheme {
    static constexpr int NG = TR - 1;                       // accumulator groups (t + u = 2 .. TR)
    static constexpr int A_BYTES = NS * BM * BKB, B_BYTES = NS * BN * BKB, STAGE_BYTES = A_BYTES + B_BYTES;
    static_assert(NG * BN <= 512, "accumulators exceed TMEM");
    static_assert(NS <= MAXS && TR <= 2 * NS && TR >= NS + 1, "bad slicing scheme");
    static constexpr int pairs() {
        int n = 0;
        for (int t = 1; t <= NS; ++t)
            for (int u = 1; u <= NS; ++u) n += (t + u <= TR);
        return n;
    }
};

// MMA schedule of one 32-byte k-step.  The accumulator of slice pair (t, u) is group t + u - 2 at TMEM columns
// (t + u - 2) * BN, and the B slices sit one after the other in shared memory ([slice][64 rows][64 B]).  For a fixed A
// slice t the pairs u = 1 .. U_t therefore read CONSECUTIVE B rows and write CONSECUTIVE accumulator columns: they are
// issued as ONE wide MMA (N = 64 * nu <= 256) instead of nu narrow ones.  A 128 x 64 x 32 MMA reads 4 KB of A and 2 KB of
// B from shared memory in 32 cycles -- more than the 128 B/clk the SM can deliver -- so the narrow schedule (21 MMAs,
// 126 KB of operand reads per k-step) is bound by shared-memory bandwidth, not by the tensor pipe; the wide schedule
// (8 MMAs for NS = 6, TR = 7) reads each A slice once or twice: 74 KB per k-step.
// Segments never mix accumulator groups that are written for the first time in a tile (overwrite) with groups that
// already hold partial sums (accumulate).
struct MmaSeg {
    int t, u0, nu, fresh;
};
template <int NS, int TR>
struct MmaPlan {
    static constexpr int MAXSEG = 4 * NS;
    MmaSeg seg[MAXSEG];
    int n;
    constexpr MmaPlan() : seg(), n(0) {
        for (int t = 1; t <= NS; ++t) {
            const int U = (TR - t < NS) ? (TR - t) : NS;
            // highest group touched by the slices before t (nondecreasing in t)
            const int prevmax = (t == 1) ? -1 : ((t + NS - 3 < TR - 2) ? (t + NS - 3) : (TR - 2));
            int u0 = 1;
            while (u0 <= U) {
                const int g0 = t + u0 - 2;
                const int fresh = g0 > prevmax;
                int nu = U - u0 + 1;
                if (nu > 4) nu = 4;
                if (!fresh && g0 + nu - 1 > prevmax) nu = prevmax - g0 + 1;
                seg[n].t = t;
                seg[n].u0 = u0;
                seg[n].nu = nu;
                seg[n].fresh = fresh;
                ++n;
                u0 += nu;
            }
        }
    }
};

// 8 / 16 TMEM columns of every accumulator group -> float64:  T = sum_g acc_g 256^(NG-1-g), combined exactly in
// two halves (|acc_g| < 2^27: hi < 2^51, lo < 2^43, see combine_groups), one rounding; result = T * 2^-(8 (NG-1) + 12).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// int32 -> float64, exact, on the ALU + FP64 pipes (the I2F/F2I conversion unit has a fraction of their throughput and
// was the limiter of the single-K-chunk kernel-matrix GEMM): 2^52 + 2^31 + x has the mantissa word x ^ 0x80000000.
__device__ __forceinline__ double i32_to_f64(int32_t x) {
    return __hiloint2double(0x43300000, (int)((uint32_t)x ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
}

// T = sum_g acc_g 256^(NG-1-g) in float64 with ONE rounding: Horner over the (up to) four most significant groups is
// exact (|.| < 2^52), so is the rest (< 2^43); the final fma rounds once -- same value as an int64 evaluation.
__host__ __device__ constexpr double pow2_neg(int e) { return e <= 0 ? 1.0 : 0.5 * pow2_neg(e - 1); }   // 2^-e

template <int NG, bool SMALLK>
__device__ __forceinline__ double combine_groups(const int32_t (*r)[16], int j) {
    constexpr double sc = pow2_neg(8 * (NG - 1) + 12);
    if (SMALLK && NG == 6) {
        // one K chunk: |acc_g| <= 64 . 127^2 . (g + 1) < 2^23, so neighbouring groups pair up exactly in int32
        const double p01 = i32_to_f64(r[0][j] * 256 + r[1][j]), p23 = i32_to_f64(r[2][j] * 256 + r[3][j]);
        const double p45 = i32_to_f64(r[4][j] * 256 + r[5][j]);
        return fma(fma(p01, 65536.0, p23), 65536.0, p45) * sc;   // inner fma exact (< 2^46), outer rounds once
    }
    constexpr int NH = NG < 4 ? NG : 4;
    double hi = i32_to_f64(r[0][j]);
#pragma unroll
    for (int g = 1; g < NH; ++g) hi = fma(hi, 256.0, i32_to_f64(r[g][j]));
    if (NG == NH) return hi * sc;
    double lo = i32_to_f64(r[NH][j]);
#pragma unroll
    for (int g = NH + 1; g < NG; ++g) lo = fma(lo, 256.0, i32_to_f64(r[g][j]));
    return fma(hi, (double)(1ll << (8 * (NG - NH))), lo) * sc;
}

template <int NG, bool SMALLK = false>
__device__ __forceinline__ void combine8(uint32_t taddr_lane_col, double* v) {
    int32_t r[NG][16];
#pragma unroll
    for (int g = 0; g < NG; ++g) tmem_ld8(taddr_lane_col + g * BN, r[g]);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = combine_groups<NG, SMALLK>(r, j);
}

template <int NG, bool SMALLK = false>
__device__ __forceinline__ void combine16(uint32_t taddr_lane_col, double* v) {
    int32_t r[NG][16];
#pragma unroll
    for (int g = 0; g < NG; ++g) tmem_ld16(taddr_lane_col + g * BN, r[g]);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = combine_groups<NG, SMALLK>(r, j);
}

// Epilogue functor:  epi(prob, row0, row, col0, v[16], M, N, aux)  (aux = Problem::aux)  called by each of the 128 epilogue threads (one
// output row each: row0 + row in the buffers, row < M valid) for every 16-column chunk of its row.
// EPW = epilogue warps per TMEM lane quarter (each takes BN / EPW columns of its 32 rows): 2 when the main loop is long
// (the epilogue hides behind the next tile's MMAs), 4 for single-chunk K, where the kernel IS its epilogue.  EPW = 4
// REQUIRES Kpad <= 64 for every problem of the launch (its accumulator read-out relies on the bound, combine_groups).
template <int NS, int TR, int STAGES, class Epi, int DEBUG_SKIP = 0, int EPW = 2>
__global__ void __launch_bounds__(64 + 128 * EPW, 1) i8gemm_kernel(const Common* __restrict__ cmp,
                                                             const Problem* __restrict__ probs, Epi epi) {
    // the work list is read once per CTA into shared memory: every role indexes it with run-time indices, which
    // would otherwise put a per-thread copy into local memory on the tile-scheduling path
    __shared__ Common cm;
    static_assert(sizeof(Common) % 4 == 0, "Common is copied word-wise");
    if (threadIdx.x < sizeof(Common) / 4) reinterpret_cast<int*>(&cm)[threadIdx.x] = reinterpret_cast<const int*>(cmp)[threadIdx.x];
    using SC = Scheme<NS, TR>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + (size_t)STAGES * SC::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t tmem_cols = 512;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4 * EPW);         // one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = cm.tile_start[cm.n_prob];

    if (warp == 0) {
        // ===================== TMA producer (converged warp, elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int gt = blockIdx.x; gt < n_tiles; gt += gridDim.x) {
            int pi = 0;
            for (int q = 1; q < cm.n_prob; ++q)
                if (gt >= cm.tile_start[q]) pi = q;
            const Problem& P = probs[pi];
            const int tile = gt - cm.tile_start[pi];
            const int tiles_n = (P.N + BN - 1) / BN;
            const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
            const int nk = P.nk_tn ? P.nk_tn[tn] : P.Kpad / BKB;
            for (int kt = 0; kt < nk; ++kt) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sa = smem + (size_t)stage * SC::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], SC::STAGE_BYTES);
                    // operands are stored K-chunk-major ([slice][chunk][row][64 B]): the 128 (64) rows of one chunk are
                    // ONE contiguous 8 (4) KB block per slice, i.e. whole cache lines on the L2 -> SM path
                    tma_load_4d(sa, &P.mapA, &full_bar[stage], 0, cm.row0[pi] + tm * BM, kt, 0);
                    tma_load_4d(sa + SC::A_BYTES, &P.mapB, &full_bar[stage], 0, tn * BN, kt, 0);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, one elected lane issues; fully unrolled) ==========
        {
            constexpr MmaPlan<NS, TR> plan{};
            const uint32_t idesc_n[4] = {make_idesc_i8(BM, BN), make_idesc_i8(BM, 2 * BN), make_idesc_i8(BM, 3 * BN),
                                         make_idesc_i8(BM, 4 * BN)};
            const uint64_t desc_hi = make_desc_sw64(0);      // everything but the start address
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int gt = blockIdx.x; gt < n_tiles; gt += gridDim.x) {
                int pi = 0;
                for (int q = 1; q < cm.n_prob; ++q)
                    if (gt >= cm.tile_start[q]) pi = q;
                const Problem& P = probs[pi];
                const int nk = P.nk_tn ? P.nk_tn[(gt - cm.tile_start[pi]) % ((P.N + BN - 1) / BN)] : P.Kpad / BKB;
                if (nk == 0) continue;                  // identically zero tile: nothing to accumulate
                mbar_wait(tmem_empty, tphase ^ 1);      // epilogue has drained the accumulators
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
                        for (int i = 0; i < plan.n; ++i) {
                            const int t = plan.seg[i].t, u0 = plan.seg[i].u0, nu = plan.seg[i].nu;
                            // the first product of a tile into a group's accumulator overwrites it
                            const uint32_t accum = (ks == 0 && plan.seg[i].fresh) ? later : 1u;
                            mma_i8(tmem_base + (t + u0 - 2) * BN, adesc + (((t - 1) * (BM * BKB) + ks * 32) >> 4),
                                   bdesc + (((u0 - 1) * (BN * BKB) + ks * 32) >> 4), idesc_n[nu - 1], accum);
                        }
                    }
                    tc_commit(&empty_bar[stage]);       // smem stage is free once these MMAs retire
                    if (kt == nk - 1) tc_commit(tmem_full);   // accumulators of this tile complete
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
        // ===================== epilogue: warps 2..; TMEM lanes 32*(warp%4)..+31, column part (warp-2)/4 ==========
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row_in_tile = q * 32 + lane;
        constexpr int HC = BN / EPW;                   // columns per epilogue warp
        static_assert(HC % 16 == 0, "the epilogue functor takes 16-column chunks");
        uint32_t tphase = 0;
        for (int gt = blockIdx.x; gt < n_tiles; gt += gridDim.x) {
            int pi = 0;
            for (int qq = 1; qq < cm.n_prob; ++qq)
                if (gt >= cm.tile_start[qq]) pi = qq;
            const Problem& P = probs[pi];
            const int tile = gt - cm.tile_start[pi];
            const int tiles_n = (P.N + BN - 1) / BN;
            const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
            double v[HC];
            if (P.nk_tn && P.nk_tn[tn] == 0) {         // identically zero tile (see Problem::nk_tn)
#pragma unroll
                for (int j = 0; j < HC; ++j) v[j] = 0.0;
#pragma unroll
                for (int cc = 0; cc < HC; cc += 16)
                    epi(pi, cm.row0[pi], tm * BM + row_in_tile, tn * BN + half * HC + cc, v + cc, cm.M[pi], P.N, P.aux);
                continue;
            }
            mbar_wait(tmem_full, tphase);
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + half * HC;
            // drain the accumulators into float64 registers, hand TMEM back to the MMA warp, and only
            // then run the (expensive) fused epilogue: it overlaps with the next tile's main loop
            if (DEBUG_SKIP != 2) {
#pragma unroll
                for (int cc = 0; cc < HC; cc += 8) combine8<SC::NG, EPW == 4>(lane_addr + cc, v + cc);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            tphase ^= 1;
            if (DEBUG_SKIP == 0) {
#pragma unroll
                for (int cc = 0; cc < HC; cc += 16)
                    epi(pi, cm.row0[pi], tm * BM + row_in_tile, tn * BN + half * HC + cc, v + cc, cm.M[pi], P.N, P.aux);
            } else if (DEBUG_SKIP == 1) {
                double s = 0;
                for (int j = 0; j < HC; ++j) s += v[j];
                if (s == 123.456) epi(pi, cm.row0[pi], tm * BM + row_in_tile, tn * BN + half * HC, v, cm.M[pi], P.N, P.aux);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

template <int NS, int STAGES>
constexpr size_t smem_bytes() { return (size_t)STAGES * NS * (BM + BN) * BKB + 1024 + 256; }

}  // namespace i8g
}  // namespace sgpr
