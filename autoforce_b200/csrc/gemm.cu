// FP64 tensor-core GEMMs of the SGPR prediction path:
//
//   (1) kernel matrix   k[i,m] = p_hat_i . z_hat_m        (similarity/universal.py:109-122:
//       torch.sparse.sum(d*dd) per (atom, inducing) pair in a Python double loop,
//       similarity/similarity.py:17-31), with the epilogue fused:
//           e_i   = sum_m mu_m k^xi           (calculator/active.py:548-550, cov @ mu)
//           G[i,m] = xi mu_m k^(xi-1)         (what autograd produces for d(cov@mu)/dk)
//   (2) back projection g_i = dE_i/dp_hat_i = sum_m G[i,m] z_hat_m   (the transposed GEMM,
//       autograd through universal.py:121 in calculator/active.py:587-599).
//
// Both are C[MxN] = A[MxK] . B[NxK]^T with K contiguous in both operands ("TN").  Operands
// are float64: the energy tolerance (1e-6 eV/atom with sum|mu| ~ 1e2..1e4) needs ~1e-10
// relative accuracy on k, so the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) is used;
// tcgen05 has no f64 kind.  cp.async multi-stage pipeline, padded shared-memory rows
// (conflict-free fragment loads), persistent tile loop sized to the SM count.
#include "sgpr_internal.cuh"

namespace sgpr {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 3, PADK = 4;
constexpr int LDS = BK + PADK;  // 20 doubles = 160 B: 4 consecutive rows hit 4 distinct 32B bank groups
constexpr int NTHREADS = 256;
constexpr int SMEM_BYTES = STAGES * (BM + BN) * LDS * (int)sizeof(double);

struct GemmArgs {
    const double* A;   // [M, lda]
    const double* B;   // [N, ldb]
    int lda, ldb;
    int M, N, K;       // K is rounded up to even by the caller (pad entries are zero)
    // epilogue 1 (kernel matrix)
    const double* mu;  // [N]
    double* G;         // [M, ldg]
    int ldg;
    int n_store;       // columns of G written (>= N, even); columns >= N get 0
    double* Kmat;      // optional [M, ldk] <- k^xi  (original column order via col_map)
    int ldk;
    const int* col_map;  // [N] column of Kmat for inducing column n
    const int* row_map;  // [M] row of Kmat for GEMM row m (or nullptr: identity + row0)
    double xi;
    int xi_int;        // xi if it is a small positive integer, else -1
    double* epart;     // [gridDim.x] per-CTA energy partial
    const unsigned char* row_owned;  // [M] count this row's energy (atom sharding), nullptr = all
    // epilogue 2 (back projection)
    double* C;         // [M, ldc]
    int ldc;
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
    // k^(xi-1)
    if (xi_int >= 1) {
        double r = 1.0;
        for (int t = 1; t < xi_int; ++t) r *= k;
        return r;
    }
    return pow(k, xi - 1.0);
}

template <int EPI>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tn_kernel(GemmArgs g) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                            // [STAGES][BM][LDS]
    double* Bs = smem + STAGES * BM * LDS;        // [STAGES][BN][LDS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;      // 4 x 2 warps, warp tile 32 x 32
    const int gid = lane >> 2, tig = lane & 3;
    const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
    const int n_tiles = tiles_m * tiles_n;
    const int nk = (g.K + BK - 1) / BK;
    double e_acc = 0.0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
        const int row0 = tm * BM, col0 = tn * BN;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        auto load_stage = [&](int stage, int kt) {
            const int k0 = kt * BK;
            // A: 128 rows x 8 chunks(16B)
#pragma unroll
            for (int it = 0; it < (BM * (BK / 2)) / NTHREADS; ++it) {
                const int ch = tid + it * NTHREADS;
                const int r = ch >> 3, c2 = (ch & 7) * 2;
                const bool ok = (row0 + r < g.M) && (k0 + c2 + 2 <= g.K);
                const double* src = ok ? g.A + (size_t)(row0 + r) * g.lda + k0 + c2 : g.A;
                cp_async16(As + (stage * BM + r) * LDS + c2, src, ok);
            }
#pragma unroll
            for (int it = 0; it < (BN * (BK / 2)) / NTHREADS; ++it) {
                const int ch = tid + it * NTHREADS;
                const int r = ch >> 3, c2 = (ch & 7) * 2;
                const bool ok = (col0 + r < g.N) && (k0 + c2 + 2 <= g.K);
                const double* src = ok ? g.B + (size_t)(col0 + r) * g.ldb + k0 + c2 : g.B;
                cp_async16(Bs + (stage * BN + r) * LDS + c2, src, ok);
            }
        };

        // prologue
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nk) load_stage(s, s);
            cp_async_commit();
        }
        for (int kt = 0; kt < nk; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            // prefetch tile kt+STAGES-1 into the stage consumed at iteration kt-1
            const int kn = kt + STAGES - 1;
            if (kn < nk) load_stage(kn % STAGES, kn);
            cp_async_commit();
            const double* a_s = As + ((kt % STAGES) * BM + wm * 32 + gid) * LDS + tig;
            const double* b_s = Bs + ((kt % STAGES) * BN + wn * 32 + gid) * LDS + tig;
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = a_s[i * 8 * LDS + kk];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = b_s[j * 8 * LDS + kk];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // all warps done with smem before the next tile's prologue

        // ---- epilogue
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = row0 + wm * 32 + i * 8 + gid;
            if (r >= g.M) continue;
            const double ew = (EPI == 1 && g.row_owned) ? (g.row_owned[r] ? 1.0 : 0.0) : 1.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = col0 + wn * 32 + j * 8 + 2 * tig;
                if (EPI == 1) {
                    double v[2];
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int cc = c + t;
                        double gv = 0.0;
                        if (cc < g.N) {
                            const double k = acc[i][j][t];
                            const double pw = powm1(k, g.xi, g.xi_int);
                            const double m = g.mu[cc];
                            gv = g.xi * m * pw;
                            e_acc += ew * (m * pw * k);
                            if (g.Kmat) {
                                const size_t kr = g.row_map ? (size_t)g.row_map[r] : (size_t)r;
                                g.Kmat[kr * g.ldk + g.col_map[cc]] = pw * k;
                            }
                        }
                        v[t] = gv;
                    }
                    if (c + 1 < g.n_store) {
                        *reinterpret_cast<double2*>(g.G + (size_t)r * g.ldg + c) = make_double2(v[0], v[1]);
                    } else if (c < g.n_store) {
                        g.G[(size_t)r * g.ldg + c] = v[0];
                    }
                } else {
                    if (c + 1 < g.N) {
                        *reinterpret_cast<double2*>(g.C + (size_t)r * g.ldc + c) = make_double2(acc[i][j][0], acc[i][j][1]);
                    } else if (c < g.N) {
                        g.C[(size_t)r * g.ldc + c] = acc[i][j][0];
                    }
                }
            }
        }
    }
    if (EPI == 1) {
        // deterministic per-CTA reduction of the energy partials
        __shared__ double red[NTHREADS / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e_acc += __shfl_xor_sync(0xffffffffu, e_acc, o);
        if (lane == 0) red[warp] = e_acc;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < NTHREADS / 32; ++w) s += red[w];
            g.epart[blockIdx.x] = s;
        }
    }
}

bool g_attr_set = false;

template <int EPI>
int launch_gemm(sgpr_context* h, const GemmArgs& a, cudaStream_t st, int grid) {
    if (!g_attr_set) {
        SGPR_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SGPR_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        g_attr_set = true;
    }
    gemm_tn_kernel<EPI><<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
    SGPR_CUDA(cudaGetLastError());
    h->stats.kernel_launches += 1;
    return SGPR_OK;
}

}  // namespace

int gemm_grid_size(sgpr_context* h) { return 2 * h->sm_count; }

// Kernel-matrix GEMM of every central species with rows (h->row_first) and inducing
// points (h->m_first).  Writes G (h->gmat, [n_rows, ldg]) and per-CTA energy partials
// into h->epart [S * grid]; optionally K^xi into Kmat (caller's row/column order).
int gemm_kernel_matrix(sgpr_context* h, double* Kmat, int ldk, const int* row_map_d, cudaStream_t st) {
    const int grid = gemm_grid_size(h);
    for (int s = 0; s < h->S; ++s) {
        const int r0 = h->row_first[s], r1 = h->row_first[s + 1];
        const int m0 = h->m_first[s], m1 = h->m_first[s + 1];
        if (r1 == r0 || m1 == m0 || !h->dp.central_enabled[s]) continue;
        GemmArgs a{};
        a.A = h->phat.as<double>() + (size_t)r0 * h->dp.ldp;
        a.lda = h->dp.ldp;
        a.B = h->zhat.as<double>() + (size_t)m0 * h->dp.ldp;
        a.ldb = h->dp.ldp;
        a.M = r1 - r0;
        a.N = m1 - m0;
        a.K = (h->dp.D + 1) & ~1;
        a.mu = h->mu.as<double>() + m0;
        a.G = h->gmat.as<double>() + (size_t)r0 * h->ldg;
        a.ldg = h->ldg;
        a.n_store = (a.N + 1) & ~1;
        a.Kmat = Kmat;
        a.ldk = ldk;
        a.col_map = h->ind_perm_d.as<int>() + m0;
        a.row_map = row_map_d ? row_map_d + r0 : nullptr;
        a.xi = h->xi;
        a.xi_int = h->xi_int;
        a.epart = h->epart.as<double>() + (size_t)s * grid;
        a.row_owned = h->active_all ? nullptr : h->row_owned.as<unsigned char>() + r0;
        SGPR_TRY(launch_gemm<1>(h, a, st, grid));
        h->stats.gemm_flops += 2.0 * a.M * (double)a.N * a.K;
    }
    return SGPR_OK;
}

// Back projection g = G . Zhat per species: C[rows, D] = G[rows, M_s] . ZhatT_s[D, M_s]^T
int gemm_back_projection(sgpr_context* h, cudaStream_t st) {
    const int grid = gemm_grid_size(h);
    for (int s = 0; s < h->S; ++s) {
        const int r0 = h->row_first[s], r1 = h->row_first[s + 1];
        const int m0 = h->m_first[s], m1 = h->m_first[s + 1];
        if (r1 == r0 || m1 == m0 || !h->dp.central_enabled[s]) continue;
        GemmArgs a{};
        a.A = h->gmat.as<double>() + (size_t)r0 * h->ldg;
        a.lda = h->ldg;
        a.B = h->zhat_t.as<double>() + (size_t)h->zt_off[s];
        a.ldb = h->ld_zt;
        a.M = r1 - r0;
        a.N = h->dp.D;
        a.K = ((m1 - m0) + 1) & ~1;
        a.C = h->gvec.as<double>() + (size_t)r0 * h->dp.ldp;
        a.ldc = h->dp.ldp;
        SGPR_TRY(launch_gemm<2>(h, a, st, grid));
        h->stats.gemm_flops += 2.0 * a.M * (double)a.N * a.K;
    }
    return SGPR_OK;
}

}  // namespace sgpr
