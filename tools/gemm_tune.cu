// Developer tool (GPU box): time tile configurations of the FP64 DMMA GEMM kernel on the c3 shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I autoforce_b200/csrc tools/gemm_tune.cu -o /tmp/gemm_tune
#include <cstdio>
#include <vector>
#include "gemm_kernel.cuh"
using namespace sgpr::gemm;

template <class C, int EPI>
float run(const char* name, int nprob, int M, int N, int K, int lda, int ldb, int ldo, double* A, double* B, double* O,
          double* mu, double* epart, int ctas_per_sm, int sms) {
    GemmBatch b{};
    for (int p = 0; p < nprob; ++p) {
        GemmArgs a{};
        a.A = A + (size_t)p * M * lda; a.lda = lda; a.B = B; a.ldb = ldb; a.M = M; a.N = N; a.K = K;
        a.mu = mu; a.G = O + (size_t)p * M * ldo; a.ldg = ldo; a.n_store = (N + 1) & ~1; a.xi = 4.0; a.xi_int = 4; a.epart = epart;
        a.C = O + (size_t)p * M * ldo; a.ldc = ldo;
        add_problem<C>(b, a);
    }
    auto kern = gemm_tn_kernel<C, EPI>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, C::SMEM);
    int grid = (ctas_per_sm > 0 ? ctas_per_sm : occ) * sms;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) kern<<<grid, C::NT, C::SMEM>>>(b);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) kern<<<grid, C::NT, C::SMEM>>>(b);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    cudaError_t err = cudaGetLastError();
    double tf = 2.0 * nprob * (double)M * N * K / (ms * 1e-3) / 1e12;
    printf("%-34s EPI%d M=%d N=%d K=%d  occ=%d grid=%d smem=%d  %.3f ms  %.2f TFLOP/s %s\n", name, EPI, M, N, K, occ, grid, C::SMEM, ms, tf,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    return ms;
}

#define RUN(WM, WN, TM, TN, BK, ST, MINB)                                                                         \
    {                                                                                                             \
        using C = Cfg<WM, WN, TM, TN, BK, ST, MINB>;                                                              \
        run<C, 1>(#WM "x" #WN " warps, " #TM "x" #TN " tiles, BK" #BK " st" #ST " mb" #MINB, 4, 24389, 500, 544, 544, 544, 512, A, B, O, mu, ep, 0, sms); \
        run<C, 2>(#WM "x" #WN " warps, " #TM "x" #TN " tiles, BK" #BK " st" #ST " mb" #MINB, 4, 24389, 544, 500, 512, 512, 544, O2, Bt, A, mu, ep, 0, sms); \
    }

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    size_t rows = 4 * 24389 + 256;
    double *A, *B, *Bt, *O, *O2, *mu, *ep;
    cudaMalloc(&A, rows * 544 * 8); cudaMalloc(&O, rows * 544 * 8); cudaMalloc(&O2, rows * 512 * 8);
    cudaMalloc(&B, 512 * 544 * 8); cudaMalloc(&Bt, 544 * 512 * 8); cudaMalloc(&mu, 4096 * 8); cudaMalloc(&ep, 4096 * 8);
    std::vector<double> h(rows * 544);
    for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 2654435761u) % 1000) * 1e-3 - 0.5;
    cudaMemcpy(A, h.data(), rows * 544 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(O2, h.data(), rows * 512 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(B, h.data(), 512 * 544 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(Bt, h.data(), 544 * 512 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(mu, h.data(), 4096 * 8, cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", prop.name, sms);
    RUN(4, 2, 4, 4, 16, 3, 2)
    RUN(4, 2, 4, 4, 16, 4, 2)
    RUN(4, 2, 4, 4, 32, 2, 2)
    RUN(4, 2, 4, 4, 32, 3, 1)
    RUN(2, 4, 4, 4, 16, 3, 2)
    RUN(2, 2, 8, 4, 16, 3, 3)
    RUN(2, 2, 4, 8, 16, 3, 3)
    RUN(2, 2, 8, 4, 16, 3, 2)
    RUN(2, 2, 4, 8, 16, 4, 2)
    RUN(2, 2, 4, 4, 16, 3, 4)
    RUN(2, 2, 4, 4, 16, 4, 6)
    RUN(4, 4, 4, 4, 16, 3, 1)
    RUN(4, 2, 4, 8, 16, 3, 1)
    RUN(2, 4, 8, 4, 16, 3, 1)
    RUN(4, 2, 2, 4, 16, 4, 3)
    RUN(4, 2, 4, 2, 16, 4, 3)
    return 0;
}
