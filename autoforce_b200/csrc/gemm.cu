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
// tcgen05 has no f64 kind.  Kernel: gemm_kernel.cuh -- cp.async 3-stage pipeline, XOR-swizzled
// shared-memory rows (conflict-free fragment loads without padding), ONE grouped persistent
// launch over all central species per GEMM.  Tile configurations were picked from a sweep on
// the c3 shapes (tools/gemm_tune.cu, profiles/r01_gemm_tune.log): 64x64 tiles / 4 CTAs per SM for
// the kernel matrix (pow/mu epilogue), 128x64 tiles with 64x32 warp tiles for the back projection.
#include "sgpr_internal.cuh"
#include "gemm_kernel.cuh"

namespace sgpr {

using namespace gemm;

namespace {

using CfgK = Cfg<2, 2, 4, 4, 16, 3, 4>;   // kernel matrix   : 64 x 64 tile, 4 warps, 4 CTAs / SM
using CfgB = Cfg<2, 2, 8, 4, 16, 3, 2>;   // back projection : 128 x 64 tile, 4 warps (64 x 32 each), 2 CTAs / SM

int g_grid[4] = {0, 0, 0, 0};

template <class C, int EPI>
int launch_gemm(sgpr_context* h, GemmBatch& b, cudaStream_t st) {
    if (b.n_prob == 0) return SGPR_OK;
    auto kern = gemm_tn_kernel<C, EPI>;
    if (g_grid[EPI] == 0) {
        SGPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int occ = 0;
        SGPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, C::SMEM));
        if (occ < 1) occ = 1;
        if (occ > C::MINB) occ = C::MINB;
        g_grid[EPI] = occ * h->sm_count;
    }
    int grid = g_grid[EPI];
    if (b.tile_start[b.n_prob] < grid) grid = b.tile_start[b.n_prob];
    kern<<<grid, C::NT, C::SMEM, st>>>(b);
    SGPR_CUDA(cudaGetLastError());
    h->stats.kernel_launches += 1;
    return SGPR_OK;
}

}  // namespace

// number of per-row energy partials written by the kernel-matrix GEMM for a species with Ms inducing points
int gemm_energy_parts(int Ms) { return ((Ms + CfgK::BN - 1) / CfgK::BN) * CfgK::WN; }

// Kernel-matrix GEMM, all central species in ONE grouped launch (rows h->row_first, inducing
// points h->m_first).  Writes G (h->gmat, [n_rows, ldg]) and per-CTA energy partials into
// h->epart [grid]; optionally K^xi into Kmat (caller's row/column order).
int gemm_kernel_matrix(sgpr_context* h, double* Kmat, int ldk, const int* row_map_d, bool store_kc, cudaStream_t st,
                       const double* wmat) {
    GemmBatch b{};
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
        a.wmat = wmat;
        a.ldk = ldk;
        a.col_map = h->ind_perm_d.as<int>() + m0;
        a.row_map = row_map_d ? row_map_d + r0 : nullptr;
        a.xi = h->xi;
        a.xi_int = h->xi_int;
        a.erow_part = h->erow_part.as<double>() + r0;
        a.erow_ld = (int)h->n_active + 1;
        a.Kc = store_kc ? h->kcmat.as<double>() + (size_t)r0 * h->ldg : nullptr;
        add_problem<CfgK>(b, a);
        h->stats.gemm_flops += 2.0 * a.M * (double)a.N * a.K;
    }
    return launch_gemm<CfgK, 1>(h, b, st);
}

// Back projection g = G . Zhat per species: C[rows, D] = G[rows, M_s] . ZhatT_s[D, M_s]^T
int gemm_back_projection(sgpr_context* h, cudaStream_t st) {
    GemmBatch b{};
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
        add_problem<CfgB>(b, a);
        h->stats.gemm_flops += 2.0 * a.M * (double)a.N * a.K;
    }
    return launch_gemm<CfgB, 2>(h, b, st);
}

// Covloss GEMM (calculator/active.py:781-783: b = choli @ cov.T; c = (b*b).sum(0)):
// B_s[rows_s, M] = K_s[rows_s, M_s] . choli[:, cols_s]^T, reduced on the fly to per-row partial
// sums of squares (h->cpart [n_part, n_rows]); K rows are zero outside the central species' block.
int gemm_covloss_parts(sgpr_context* h) { return ((h->M + CfgB::BN - 1) / CfgB::BN) * CfgB::WN; }

int gemm_covloss(sgpr_context* h, int64_t n_rows, cudaStream_t st) {
    GemmBatch b{};
    for (int s = 0; s < h->S; ++s) {
        const int r0 = h->row_first[s], r1 = h->row_first[s + 1];
        const int m0 = h->m_first[s], m1 = h->m_first[s + 1];
        if (r1 == r0 || m1 == m0 || !h->dp.central_enabled[s]) continue;
        GemmArgs a{};
        a.A = h->kcmat.as<double>() + (size_t)r0 * h->ldg;
        a.lda = h->ldg;
        a.B = h->choli_t.as<double>() + (size_t)s * h->M * h->ld_zt;
        a.ldb = h->ld_zt;
        a.M = r1 - r0;
        a.N = h->M;
        a.K = ((m1 - m0) + 1) & ~1;
        a.part = h->cpart.as<double>() + r0;
        a.part_ld = (int)n_rows;
        add_problem<CfgB>(b, a);
        h->stats.covloss_flops += 2.0 * a.M * (double)a.N * a.K;
    }
    return launch_gemm<CfgB, 3>(h, b, st);
}

}  // namespace sgpr
