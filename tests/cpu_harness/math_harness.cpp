// Test-only: compiles autoforce_b200/csrc/sgpr_math.cuh (the header the CUDA kernels use)
// with the host compiler so that tests/test_math_header.py can check the harmonics /
// radial functions against the oracle without a GPU.  Not part of the product.
#include "../../autoforce_b200/csrc/sgpr_math.cuh"

using namespace sgpr;

template <int LMAX>
static void run(int lmax, int n, const double* xyz, double* Y, double* dY) {
    HarmCoef hc;
    fill_harm_coef(hc);
    const int L2 = (lmax + 1) * (lmax + 1);
    for (int i = 0; i < n; ++i) {
        double* y = Y + (size_t)i * L2;
        double* d = dY + (size_t)i * L2 * 3;
        solid_harmonics<LMAX, true>(hc, lmax, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2],
                                    [&](int idx, double v, double gx, double gy, double gz) {
                                        y[idx] = v;
                                        d[3 * idx] = gx;
                                        d[3 * idx + 1] = gy;
                                        d[3 * idx + 2] = gz;
                                    });
    }
}

extern "C" void harness_harmonics(int lmax, int bucket, int n, const double* xyz, double* Y, double* dY) {
    if (bucket == 3) run<3>(lmax, n, xyz, Y, dY);
    else if (bucket == 6) run<6>(lmax, n, xyz, Y, dY);
    else run<8>(lmax, n, xyz, Y, dY);
}

extern "C" void harness_radial(int n, const double* d, double u, double rc, double* R, double* Rpd) {
    for (int i = 0; i < n; ++i) radial(d[i], u, rc, R[i], Rpd[i]);
}

extern "C" double harness_anl(int n, int l) { return anl(n, l); }
