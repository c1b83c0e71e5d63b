// Templated FP64 DMMA "TN" GEMM kernel  C[MxN] = A[MxK] . B[NxK]^T  (K contiguous in both
// operands), shared by gemm.cu (product) and tools/gemm_tune.cu (tuning harness).
// See gemm.cu for what the two epilogues compute and which reference lines they replace.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgpr {
namespace gemm {

struct GemmArgs {
    const double* A;   // [M, lda]
    const double* B;   // [N, ldb]
    int lda, ldb;
    int M, N, K;       // K rounded up to even by the caller (pad entries are zero)
    // epilogue 1 (kernel matrix)
    const double* mu;  // [N]
    double* G;         // [M, ldg]
    int ldg;
    int n_store;       // columns of G written (>= N, even); columns >= N get 0
    double* Kmat;      // optional [*, ldk] <- k^xi (caller's row/column order via the maps)
    int ldk;
    const int* col_map;  // [N]
    const int* row_map;  // [M] or nullptr
    double xi;
    int xi_int;        // xi if it is a small positive integer, else -1
    double* erow_part; // per-row energy partials: erow_part[(tile_n * WN + warp_n) * erow_ld + row]
    int erow_ld;
    double* Kc;        // optional [M, ldg] <- k^xi in GEMM row/column order (input of the covloss GEMM)
    const double* wmat;  // optional per-element weights dL/dK [*, ldk] (caller's order via the maps) used
                         // INSTEAD of mu[col]: vector-Jacobian product of the kernel matrix
    // epilogue 3 (covloss): per-row partial sums of squares, part[(tile_n * WN + warp_n) * part_ld + row]
    double* part;
    int part_ld;
    // epilogue 2 (back projection)
    double* C;         // [M, ldc]
    int ldc;
};

constexpr int kMaxProb = 8;
struct GemmBatch {
    int n_prob;
    int tile_start[kMaxProb + 1];   // first tile of each problem (problem-major tile order)
    GemmArgs p[kMaxProb];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int bytes = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double powm1(double k, double xi, int xi_int) {
    if (xi_int >= 1) {  // k^(xi-1) by repeated multiplication
        double r = 1.0;
        for (int t = 1; t < xi_int; ++t) r *= k;
        return r;
    }
    return pow(k, xi - 1.0);
}

// Tile configuration: WM x WN warps, each owning TM x TN DMMA tiles of 8x8; BK doubles of K per
// pipeline stage.  Shared-memory rows hold BK doubles without padding; the 16-byte chunk c of row
// r is stored at chunk position c ^ ((r & 3) << 1), so a fragment load (4 rows x 4 consecutive
// doubles per half-warp) touches 8 distinct 16 B chunks -> conflict-free.
template <int WM_, int WN_, int TM_, int TN_, int BK_, int STAGES_, int MINB_>
struct Cfg {
    static constexpr int WM = WM_, WN = WN_, TM = TM_, TN = TN_, BK = BK_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int BM = WM * TM * 8, BN = WN * TN * 8;
    static constexpr int NT = 32 * WM * WN;
    static constexpr int CH = BK / 2;  // 16 B chunks per row
    static constexpr int SMEM = STAGES * (BM + BN) * BK * (int)sizeof(double);
};

__device__ __forceinline__ int swz(int row, int chunk) { return chunk ^ ((row & 3) << 1); }

template <class C, int EPI>
__global__ void __launch_bounds__(C::NT, C::MINB) gemm_tn_kernel(const __grid_constant__ GemmBatch batch) {
    constexpr int BM = C::BM, BN = C::BN, BK = C::BK, STAGES = C::STAGES, NT = C::NT, TM = C::TM, TN = C::TN, CH = C::CH;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                           // [STAGES][BM][BK]
    double* Bs = smem + STAGES * BM * BK;        // [STAGES][BN][BK]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / C::WN, wn = warp % C::WN;
    const int gid = lane >> 2, tig = lane & 3;
    int koff[BK / 4];   // swizzled fragment column offsets of the k4 steps of a stage
#pragma unroll
    for (int q = 0; q < BK / 4; ++q) koff[q] = swz(gid, 2 * q + (tig >> 1)) * 2 + (tig & 1);
    const int n_tiles = batch.tile_start[batch.n_prob];

    for (int gtile = blockIdx.x; gtile < n_tiles; gtile += gridDim.x) {
        // problem (central species) of this tile; static unroll keeps the parameters in the constant bank
        GemmArgs g = batch.p[0];
        int tile = gtile;
#pragma unroll
        for (int q = 1; q < kMaxProb; ++q)
            if (q < batch.n_prob && gtile >= batch.tile_start[q]) {
                g = batch.p[q];
                tile = gtile - batch.tile_start[q];
            }
        const int tiles_n = (g.N + BN - 1) / BN;
        const int nk = (g.K + BK - 1) / BK;
        const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
        const int row0 = tm * BM, col0 = tn * BN;
        double acc[TM][TN][2];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        auto load_stage = [&](int stage, int kt) {
            const int k0 = kt * BK;
#pragma unroll
            for (int it = 0; it < (BM * CH + NT - 1) / NT; ++it) {
                const int ch = tid + it * NT;
                if ((BM * CH) % NT == 0 || ch < BM * CH) {
                    const int r = ch / CH, c = ch % CH;
                    const bool ok = (row0 + r < g.M) && (k0 + 2 * c + 2 <= g.K);
                    const double* src = ok ? g.A + (size_t)(row0 + r) * g.lda + k0 + 2 * c : g.A;
                    cp_async16(As + (stage * BM + r) * BK + swz(r, c) * 2, src, ok);
                }
            }
#pragma unroll
            for (int it = 0; it < (BN * CH + NT - 1) / NT; ++it) {
                const int ch = tid + it * NT;
                if ((BN * CH) % NT == 0 || ch < BN * CH) {
                    const int r = ch / CH, c = ch % CH;
                    const bool ok = (col0 + r < g.N) && (k0 + 2 * c + 2 <= g.K);
                    const double* src = ok ? g.B + (size_t)(col0 + r) * g.ldb + k0 + 2 * c : g.B;
                    cp_async16(Bs + (stage * BN + r) * BK + swz(r, c) * 2, src, ok);
                }
            }
        };

        // warp-uniform: does this warp own any real row / column of the tile?
        const bool warp_live = (row0 + wm * TM * 8 < g.M) && (col0 + wn * TN * 8 < g.N);

#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nk) load_stage(s, s);
            cp_async_commit();
        }
        for (int kt = 0; kt < nk; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int kn = kt + STAGES - 1;   // refill the stage consumed at iteration kt-1
            if (kn < nk) load_stage(kn % STAGES, kn);
            cp_async_commit();
            if (warp_live) {
                const double* a_s = As + ((kt % STAGES) * BM + wm * TM * 8 + gid) * BK;
                const double* b_s = Bs + ((kt % STAGES) * BN + wn * TN * 8 + gid) * BK;
#pragma unroll
                for (int q = 0; q < BK / 4; ++q) {
                    double a[TM], b[TN];
#pragma unroll
                    for (int i = 0; i < TM; ++i) a[i] = a_s[i * 8 * BK + koff[q]];
#pragma unroll
                    for (int j = 0; j < TN; ++j) b[j] = b_s[j * 8 * BK + koff[q]];
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // all warps done with smem before the next tile's prologue

        // ---- epilogue
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = row0 + wm * TM * 8 + i * 8 + gid;
            const bool rv = r < g.M;   // no early exit: the quad shuffles below need every lane
            double e_row = 0.0;   // EPI 1: sum_m mu_m k^xi over this warp's columns
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int c = col0 + wn * TN * 8 + j * 8 + 2 * tig;
                if (!rv) continue;
                if (EPI == 1) {
                    double v[2];
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int cc = c + t;
                        double gv = 0.0;
                        if (cc < g.N) {
                            const double k = acc[i][j][t];
                            const double pw = powm1(k, g.xi, g.xi_int);
                            const double m = g.wmat ? g.wmat[(g.row_map ? (size_t)g.row_map[r] : (size_t)r) * g.ldk + g.col_map[cc]]
                                                    : g.mu[cc];
                            gv = g.xi * m * pw;
                            e_row += m * pw * k;
                            if (g.Kmat) {
                                const size_t kr = g.row_map ? (size_t)g.row_map[r] : (size_t)r;
                                g.Kmat[kr * g.ldk + g.col_map[cc]] = pw * k;
                            }
                            if (g.Kc) g.Kc[(size_t)r * g.ldg + cc] = pw * k;
                        } else if (g.Kc && cc < g.n_store) {
                            g.Kc[(size_t)r * g.ldg + cc] = 0.0;
                        }
                        v[t] = gv;
                    }
                    if (c + 1 < g.n_store) {
                        *reinterpret_cast<double2*>(g.G + (size_t)r * g.ldg + c) = make_double2(v[0], v[1]);
                    } else if (c < g.n_store) {
                        g.G[(size_t)r * g.ldg + c] = v[0];
                    }
                } else if (EPI == 2) {
                    if (c + 1 < g.N) {
                        *reinterpret_cast<double2*>(g.C + (size_t)r * g.ldc + c) = make_double2(acc[i][j][0], acc[i][j][1]);
                    } else if (c < g.N) {
                        g.C[(size_t)r * g.ldc + c] = acc[i][j][0];
                    }
                }
            }
            if (EPI == 1) {
                e_row += __shfl_xor_sync(0xffffffffu, e_row, 1);
                e_row += __shfl_xor_sync(0xffffffffu, e_row, 2);
                if (tig == 0 && rv) g.erow_part[(size_t)(tn * C::WN + wn) * g.erow_ld + r] = e_row;
            }
        }
        if (EPI == 3) {
            // row-wise sum of squares of this warp's 8*TN columns, reduced over the quad
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const int r = row0 + wm * TM * 8 + i * 8 + gid;
                double ss = 0.0;
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int c = col0 + wn * TN * 8 + j * 8 + 2 * tig;
                    if (c < g.N) ss += acc[i][j][0] * acc[i][j][0];
                    if (c + 1 < g.N) ss += acc[i][j][1] * acc[i][j][1];
                }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                if (tig == 0 && r < g.M) g.part[(size_t)(tn * C::WN + wn) * g.part_ld + r] = ss;
            }
        }
    }
}

template <class C>
inline void add_problem(GemmBatch& b, const GemmArgs& a) {
    const int tiles = ((a.M + C::BM - 1) / C::BM) * ((a.N + C::BN - 1) / C::BN);
    b.p[b.n_prob] = a;
    b.tile_start[b.n_prob + 1] = b.tile_start[b.n_prob] + tiles;
    b.n_prob++;
}

}  // namespace gemm
}  // namespace sgpr
