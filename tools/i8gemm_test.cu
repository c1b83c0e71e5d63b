// Developer tool (GPU box): correctness + speed of the tcgen05 int8-sliced GEMM vs float64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I autoforce_b200/csrc -I tools tools/i8gemm_test.cu -o tools/i8gemm_test.bin -lcuda
//   stages 63 / 62 select the experimental A-in-TMEM kernel (tools/i8gemm_ta_kernel.cuh)
//   stages 72 / 74 select the cluster kernel with multicast A (i8gemm_mc_kernel.cuh), 2 / 4 CTAs
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "i8gemm_kernel.cuh"
#include "i8gemm_ta_kernel.cuh"
#include "i8gemm_mc_kernel.cuh"
using namespace sgpr::i8g;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    return (EncodeFn)fn;
}
static int make_map(CUtensorMap* m, void* base, int ns, int rows, int Kpad, int box_rows) {
    static EncodeFn enc = get_encode();
    cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)rows, (cuuint64_t)ns};
    cuuint64_t strides[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows * Kpad};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)ns};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, getenv("PROMO") ? (CUtensorMapL2promotion)atoi(getenv("PROMO")) : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return (int)r;
}

// production layout: K-chunk-major [slice][chunk of 64][row][64]  (4-D map, box {64, rows, 1, ns})
static int make_map_cm(CUtensorMap* m, void* base, int ns, int rows, int Kpad, int box_rows) {
    static EncodeFn enc = get_encode();
    cuuint64_t dims[4] = {64, (cuuint64_t)rows, (cuuint64_t)(Kpad / 64), (cuuint64_t)ns};
    cuuint64_t strides[3] = {64, (cuuint64_t)rows * 64, (cuuint64_t)rows * Kpad};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, (cuuint32_t)ns};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, getenv("PROMO") ? (CUtensorMapL2promotion)atoi(getenv("PROMO")) : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled (chunk-major) failed: %d\n", (int)r);
    return (int)r;
}
static int make_map_cm_part(CUtensorMap* m, void* base, int ns, int rows, int Kpad, int box_rows) {
    static EncodeFn enc = get_encode();
    cuuint64_t dims[4] = {64, (cuuint64_t)rows, (cuuint64_t)(Kpad / 64), (cuuint64_t)ns};
    cuuint64_t strides[3] = {64, (cuuint64_t)rows * 64, (cuuint64_t)rows * Kpad};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled (part) failed: %d\n", (int)r);
    return (int)r;
}
static void to_chunk_major(const std::vector<int8_t>& in, int rows, int Kpad, int ns, std::vector<int8_t>& out) {
    out.assign(in.size(), 0);
    for (int t = 0; t < ns; ++t)
        for (int r = 0; r < rows; ++r)
            for (int k = 0; k < Kpad; ++k)
                out[(size_t)t * rows * Kpad + (size_t)(k / 64) * rows * 64 + (size_t)r * 64 + k % 64] = in[((size_t)t * rows + r) * Kpad + k];
}

// host slicing: x in [-1,1] -> ns balanced base-128 digits, most significant first
static void slice_rows(const std::vector<double>& X, int rows, int K, int Kpad, int ns, std::vector<int8_t>& out) {
    out.assign((size_t)ns * rows * Kpad, 0);
    const double sc = ldexp(1.0, 8 * ns - 2);
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < K; ++k) {
            long long v = llrint(X[(size_t)r * K + k] * sc);
            for (int t = ns; t >= 1; --t) {
                long long d = ((v + 128) % 256 + 256) % 256 - 128;
                out[((size_t)(t - 1) * rows + r) * Kpad + k] = (int8_t)d;
                v = (v - d) / 256;
            }
        }
}

struct StoreEpi {
    double* C;
    int ldc;
    __device__ void operator()(int prob, int row, int col0, const double* v, int M, int N) const {
        if (row >= M) return;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (col0 + j < N) C[(size_t)row * ldc + col0 + j] = v[j];
    }
    // production kernel: rows are offsets from the problem's first row (Common::row0)
    __device__ void operator()(int prob, int row0, int row, int col0, const double* v, int M, int N, const double* = nullptr) const {
        if (row < M) (*this)(prob, row0 + row, col0, v, row0 + M, N);
    }
};

int main(int argc, char** argv) {
    int M = argc > 1 ? atoi(argv[1]) : 300, N = argc > 2 ? atoi(argv[2]) : 150, K = argc > 3 ? atoi(argv[3]) : 544;
    int ns = argc > 4 ? atoi(argv[4]) : 6, tr = argc > 5 ? atoi(argv[5]) : 8, stages = argc > 6 ? atoi(argv[6]) : 3;
    int check = argc > 7 ? atoi(argv[7]) : 1;
    const int Kpad = (K + 63) / 64 * 64;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("%s: M=%d N=%d K=%d ns=%d tr=%d stages=%d\n", prop.name, M, N, K, ns, tr, stages);
    std::vector<double> A((size_t)M * K), B((size_t)N * K);
    srand(1);
    auto fill = [&](std::vector<double>& X, int rows) {
        for (int r = 0; r < rows; ++r) {
            double nrm = 0;
            for (int k = 0; k < K; ++k) {
                double v = (rand() / (double)RAND_MAX - 0.3) * exp(-6.0 * (rand() / (double)RAND_MAX));
                if (r == 0) v = (k == 3) ? 1.0 : 0.0;   // a row with a single unit entry: extreme digit
                X[(size_t)r * K + k] = v;
                nrm += v * v;
            }
            nrm = sqrt(nrm);
            for (int k = 0; k < K; ++k) X[(size_t)r * K + k] /= nrm;
        }
    };
    fill(A, M);
    fill(B, N);
    std::vector<int8_t> As, Bs;
    slice_rows(A, M, K, Kpad, ns, As);
    slice_rows(B, N, K, Kpad, ns, Bs);
    int8_t *dA, *dB;
    double* dC;
    cudaMalloc(&dA, As.size());
    cudaMalloc(&dB, Bs.size());
    cudaMalloc(&dC, (size_t)M * N * 8);
    cudaMemcpy(dA, As.data(), As.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bs.data(), Bs.size(), cudaMemcpyHostToDevice);
    cudaMemset(dC, 0, (size_t)M * N * 8);
    // the production kernel reads K-chunk-major operands, the experimental A-in-TMEM kernel row-major ones
    std::vector<int8_t> Ac, Bc;
    to_chunk_major(As, M, Kpad, ns, Ac);
    to_chunk_major(Bs, N, Kpad, ns, Bc);
    int8_t *dAc, *dBc;
    cudaMalloc(&dAc, Ac.size());
    cudaMalloc(&dBc, Bc.size());
    cudaMemcpy(dAc, Ac.data(), Ac.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dBc, Bc.data(), Bc.size(), cudaMemcpyHostToDevice);
    Problem P;
    P.M = M; P.N = N; P.Kpad = Kpad;
    P.nk_tn = nullptr;
    P.aux = nullptr;
    if (make_map_cm(&P.mapA, dAc, ns, M, Kpad, BM) || make_map_cm(&P.mapB, dBc, ns, N, Kpad, BN)) return 1;
    ProblemTA PT;
    if (make_map(&PT.mapA, dA, ns, M, Kpad, BM) || make_map(&PT.mapB, dB, ns, N, Kpad, BN)) return 1;
    PT.M = M; PT.N = N; PT.Kpad = Kpad;
    PT.A = (const signed char*)dA;
    PT.a_slice = (long long)M * Kpad;
    PT.a_ld = Kpad;
    ProblemTA* dPT;
    cudaMalloc(&dPT, sizeof(ProblemTA));
    cudaMemcpy(dPT, &PT, sizeof(ProblemTA), cudaMemcpyHostToDevice);
    Problem* dP;
    cudaMalloc(&dP, sizeof(Problem));
    cudaMemcpy(dP, &P, sizeof(Problem), cudaMemcpyHostToDevice);
    Common cm{};
    cm.n_prob = 1;
    cm.tile_start[0] = 0;
    cm.tile_start[1] = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    cm.row0[0] = 0;
    cm.M[0] = M;
    Common* dcm;   // the production kernel reads its work list from device memory (written by i8_setup_kernel there)
    cudaMalloc(&dcm, sizeof(Common));
    cudaMemcpy(dcm, &cm, sizeof(Common), cudaMemcpyHostToDevice);
    StoreEpi epi{dC, N};
    void (*kern_mc)(const Common*, const ProblemMC*, StoreEpi) = nullptr;
    ProblemMC* dPM = nullptr;
    int mc = 0;
    if (ns == 6 && tr == 7 && (stages == 72 || stages == 74)) {
        mc = stages == 72 ? 2 : 4;
        ProblemMC PM;
        PM.N = N; PM.Kpad = Kpad; PM.aux = nullptr;
        PM.mapB = P.mapB;
        if (make_map_cm_part(&PM.mapAs, dAc, ns, M, Kpad, BM / mc)) return 1;
        cudaMalloc(&dPM, sizeof(ProblemMC));
        cudaMemcpy(dPM, &PM, sizeof(ProblemMC), cudaMemcpyHostToDevice);
        if (mc == 2) kern_mc = i8gemm_mc_kernel<6, 7, 3, StoreEpi, 2>;
        else kern_mc = i8gemm_mc_kernel<6, 7, 3, StoreEpi, 4>;
    }
    void (*kern)(const Common*, const Problem*, StoreEpi) = nullptr;
    void (*kern_ta)(const Common, const ProblemTA*, StoreEpi) = nullptr;
    size_t smem = 0;
    int nthreads = NTHREADS;
    if (mc) { smem = smem_bytes<6, 3>(); }
    else if (ns == 6 && tr == 7 && stages == 63) { kern_ta = i8gemm_ta_kernel<6, 7, 3, StoreEpi, 3>; smem = smem_bytes_ta<6, 3, 3>(); nthreads = NTHREADS_TA; }
    else if (ns == 6 && tr == 7 && stages == 62) { kern_ta = i8gemm_ta_kernel<6, 7, 4, StoreEpi, 2>; smem = smem_bytes_ta<6, 4, 2>(); nthreads = NTHREADS_TA; }
    else if (ns == 6 && tr == 8 && stages == 31) { kern = i8gemm_kernel<6, 8, 3, StoreEpi, 1>; smem = smem_bytes<6, 3>(); }
    else if (ns == 6 && tr == 8 && stages == 32) { kern = i8gemm_kernel<6, 8, 3, StoreEpi, 2>; smem = smem_bytes<6, 3>(); }
    else if (ns == 6 && tr == 8 && stages == 2) { kern = i8gemm_kernel<6, 8, 2, StoreEpi>; smem = smem_bytes<6, 2>(); }
    else if (ns == 6 && tr == 8) { kern = i8gemm_kernel<6, 8, 3, StoreEpi>; smem = smem_bytes<6, 3>(); }
    else if (ns == 6 && tr == 7) { kern = i8gemm_kernel<6, 7, 3, StoreEpi>; smem = smem_bytes<6, 3>(); }
    else if (ns == 7 && tr == 9) { kern = i8gemm_kernel<7, 9, 2, StoreEpi>; smem = smem_bytes<7, 2>(); }
    else if (ns == 5 && tr == 7) { kern = i8gemm_kernel<5, 7, 3, StoreEpi>; smem = smem_bytes<5, 3>(); }
    else if (ns == 5 && tr == 6) { kern = i8gemm_kernel<5, 6, 4, StoreEpi>; smem = smem_bytes<5, 4>(); }
    else { printf("unsupported scheme\n"); return 1; }
    cudaError_t e = mc ? cudaFuncSetAttribute(kern_mc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                  : kern_ta ? cudaFuncSetAttribute(kern_ta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("smem %zu bytes: %s\n", smem, cudaGetErrorString(e));
    int grid = std::min(prop.multiProcessorCount, cm.tile_start[1]);
    if (getenv("GRID")) grid = std::min(grid, atoi(getenv("GRID")));
    if (mc) {
        grid = prop.multiProcessorCount / mc * mc;
        if (getenv("GRID")) grid = std::min(grid, atoi(getenv("GRID")) / mc * mc);
        int ncl = 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(nthreads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = mc; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        cudaOccupancyMaxActiveClusters(&ncl, kern_mc, &cfg);
        printf("cluster size %d: grid %d, max active clusters %d\n", mc, grid, ncl);
        kern_mc<<<grid, nthreads, smem>>>(dcm, dPM, epi);
    } else if (kern_ta) kern_ta<<<grid, nthreads, smem>>>(cm, dPT, epi);
    else kern<<<grid, nthreads, smem>>>(dcm, dP, epi);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    if (check) {
        std::vector<double> C((size_t)M * N);
        cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += A[(size_t)i * K + k] * B[(size_t)j * K + k];
                maxerr = fmax(maxerr, fabs(s - C[(size_t)i * N + j]));
                maxref = fmax(maxref, fabs(s));
            }
        printf("max |C - ref| = %.3e  (max |ref| = %.3f)\n", maxerr, maxref);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) {
        if (mc) kern_mc<<<grid, nthreads, smem>>>(dcm, dPM, epi);
        else if (kern_ta) kern_ta<<<grid, nthreads, smem>>>(cm, dPT, epi);
        else kern<<<grid, nthreads, smem>>>(dcm, dP, epi);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    int npairs = 0;
    for (int t = 1; t <= ns; ++t)
        for (int u = 1; u <= ns; ++u) npairs += (t + u <= tr);
    printf("%.3f ms : %.2f 'FP64-equivalent' TFLOP/s, %.1f int8 TOP/s (%d slice products)\n", ms, 2.0 * M * N * K / ms / 1e9,
           2.0 * M * N * Kpad * npairs / ms / 1e9, npairs);
    return 0;
}
