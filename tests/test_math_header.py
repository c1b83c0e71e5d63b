"""CPU: the math header used by the CUDA kernels (autoforce_b200/csrc/sgpr_math.cuh),
compiled with g++ through a test-only harness, against the oracle's restatement of
the reference's spherical harmonics (descriptor/ylm.py:113-225) and radial function."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import sgpr_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("harness") / "libmath_harness.so")
    src = os.path.join(HERE, "cpu_harness", "math_harness.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    return ctypes.CDLL(out)


def _harm(lib, lmax, bucket, xyz):
    n = len(xyz)
    L2 = (lmax + 1) ** 2
    Y = np.zeros((n, L2))
    dY = np.zeros((n, L2, 3))
    P = ctypes.POINTER(ctypes.c_double)
    lib.harness_harmonics(lmax, bucket, n, xyz.ctypes.data_as(P), Y.ctypes.data_as(P), dY.ctypes.data_as(P))
    return Y, dY


@pytest.mark.parametrize("lmax,bucket", [(0, 3), (2, 3), (3, 3), (3, 6), (5, 6), (6, 6), (6, 8), (8, 8)])
def test_solid_harmonics_and_gradients(harness, lmax, bucket):
    rng = np.random.default_rng(lmax)
    xyz = np.ascontiguousarray(rng.normal(0, 1.5, (50, 3)))
    Y, dY = _harm(harness, lmax, bucket, xyz)
    tab = o.YlmTables(lmax)
    Yo, dYo = o.ylm(xyz, tab, np.zeros(len(xyz), dtype=bool), grad=True)
    for l in range(lmax + 1):
        for m in range(l + 1):
            w = 1.0 if m == 0 else np.sqrt(2.0)  # sqrt(2 - delta_m0) folded into the header's Y
            idx = l * l + (0 if m == 0 else 2 * m - 1)
            scale = np.abs(Yo[:, l, l - m]).max() + 1e-300
            assert np.abs(Y[:, idx] - w * Yo[:, l, l - m]).max() < 1e-13 * scale
            assert np.abs(dY[:, idx] - w * dYo[:, l, l - m]).max() < 1e-11 * max(1.0, np.abs(dYo[:, l, l - m]).max())
            if m > 0:
                assert np.abs(Y[:, idx + 1] - w * Yo[:, l - m, l]).max() < 1e-13 * scale
                assert np.abs(dY[:, idx + 1] - w * dYo[:, l - m, l]).max() < 1e-11 * max(1.0, np.abs(dYo[:, l - m, l]).max())


def test_harmonics_regular_on_z_axis(harness):
    # the polynomial form needs no special-casing on the z axis (the reference divides by sin(theta))
    xyz = np.array([[0.0, 0.0, 1.3], [0.0, 0.0, -0.7], [1e-9, -1e-9, 2.0]])
    Y, dY = _harm(harness, 6, 6, xyz)
    assert np.isfinite(Y).all() and np.isfinite(dY).all()
    eps = 1e-6
    for k in range(3):
        dx = np.zeros(3)
        dx[k] = eps
        Yp, _ = _harm(harness, 6, 6, np.ascontiguousarray(xyz + dx))
        Ym, _ = _harm(harness, 6, 6, np.ascontiguousarray(xyz - dx))
        assert np.abs((Yp - Ym) / (2 * eps) - dY[:, :, k]).max() < 1e-6


def test_radial_and_nnl(harness):
    rc, u = 6.0, 0.5
    d = np.linspace(0.3, 13.0, 200)
    R = np.zeros_like(d)
    Rpd = np.zeros_like(d)
    P = ctypes.POINTER(ctypes.c_double)
    harness.harness_radial(len(d), d.ctypes.data_as(P), ctypes.c_double(u), ctypes.c_double(rc), R.ctypes.data_as(P), Rpd.ctypes.data_as(P))
    m = o.OracleModel(lmax=1, nmax=1, xi=4, rc=rc, radii={}, ind_Z=np.zeros(0), ind_r=[], ind_b=[], mu=np.zeros(0))
    Ro, dRo = o._radial(m, np.full_like(d, u), d, grad=True)
    assert np.abs(R - Ro).max() < 1e-15
    assert np.abs(Rpd * d - dRo).max() < 1e-14
    harness.harness_anl.restype = ctypes.c_double
    nnl = o.nnl_table(6, 8)
    for n1 in range(9):
        for l in range(7):
            a = harness.harness_anl(n1, l)
            assert abs(np.sqrt(a * a) - nnl[n1, n1, l]) < 1e-15 * nnl[n1, n1, l] + 1e-300
