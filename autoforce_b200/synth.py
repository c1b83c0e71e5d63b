"""Deterministic synthetic structures and frozen models of the shapes BASELINE.json
names.  Product-side generator used by bench.py / smoke();
the inducing environments are cut out of a source structure with the GPU neighbour
list (sgpr_neighbors), so nothing here touches the oracle."""
from __future__ import annotations

import numpy as np

from .model import SgprModel

A0 = 3.61  # fcc Cu lattice constant used throughout the survey


def fcc(rep, Zs, sigma, seed, a0=A0):
    rng = np.random.default_rng(seed)
    rep = (rep,) * 3 if np.isscalar(rep) else tuple(rep)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    ii, jj, kk = np.meshgrid(np.arange(rep[0]), np.arange(rep[1]), np.arange(rep[2]), indexing="ij")
    cells = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)
    pos = (cells[:, None, :] + base[None]).reshape(-1, 3) * a0
    if sigma > 0:
        pos = pos + rng.normal(0, sigma, pos.shape)
    numbers = rng.choice(np.asarray(Zs), len(pos)).astype(np.int32)
    cell = np.diag([a0 * r for r in rep]).astype(float)
    return pos, cell, numbers


def cut_environments(pos, cell, pbc, numbers, rc, indices, first, J, S):
    envs = []
    for i in indices:
        sl = slice(first[i], first[i + 1])
        r = pos[J[sl]] - pos[i] + S[sl].astype(float) @ np.asarray(cell, dtype=float)
        envs.append((int(numbers[i]), r, numbers[J[sl]].copy()))
    return envs


def synth_model(Zs, M, seed, lmax=3, nmax=3, xi=4, rc=6.0, kind="sesoap", src_rep=5, mu_scale=0.1, with_choli=False,
                neighbors_fn=None):
    """Frozen model: M inducing LCEs drawn per-species-balanced from fcc(src_rep, Zs, 0.15, seed+100),
    mu ~ N(0,1)*mu_scale, mean weight -3.0 per species, vscale = 1, choli = 0.5 I (with_choli=True) or a
    dense lower-triangular matrix (with_choli="tril").
    ``neighbors_fn(pos, cell, pbc, rc) -> (first, j, S)`` overrides the GPU neighbour hook
    (the CPU arm of bench.py passes the oracle's, so that it needs no GPU)."""
    rng = np.random.default_rng(seed)
    pos, cell, numbers = fcc(src_rep, Zs, 0.15, seed + 100)
    base = dict(lmax=lmax, nmax=nmax, xi=float(xi), rc=float(rc), kind=kind, radii={1: 0.5} if kind == "sesoap" else {},
                default_radius=1.0 if kind == "sesoap" else rc / 6)
    if neighbors_fn is not None:
        first, J, S = neighbors_fn(pos, cell, True, rc)
    else:
        from .engine import SgprEngine

        probe = SgprEngine(SgprModel.from_envs([], **base), species=sorted(set(int(z) for z in Zs)))
        try:
            first, J, S = probe.neighbors(pos, numbers, cell, True)
        finally:
            probe.close()
    sel = []
    Zs = sorted(set(int(z) for z in Zs))
    for k, z in enumerate(Zs):
        idx = np.nonzero(numbers == z)[0]
        m = M // len(Zs) + (1 if k < M % len(Zs) else 0)
        sel += list(rng.choice(idx, m, replace=len(idx) < m))
    envs = cut_environments(pos, cell, True, numbers, rc, sel, first, J, S)
    Mtot = len(envs)
    mu = rng.normal(0, 1, Mtot) * mu_scale
    choli = None
    if with_choli == "tril":
        # dense lower-triangular like the reference's choli = L^-1 (regression/gppotential.py:588)
        choli = 0.5 * np.eye(Mtot) + np.tril(np.random.default_rng(seed + 7).normal(0, 0.02 / np.sqrt(Mtot), (Mtot, Mtot)), -1)
    elif with_choli:
        choli = 0.5 * np.eye(Mtot)
    return SgprModel.from_envs(envs, mu=mu, mean_w={z: -3.0 for z in Zs}, vscale={z: 1.0 for z in Zs}, choli=choli, **base)


WORKLOADS = {
    # name: (rep, species, M, lmax, nmax, rc)        SURVEY.md section 8d
    "c1": dict(rep=3, Zs=[29], M=50, lmax=3, nmax=3, rc=6.0),
    "c2": dict(rep=10, Zs=[29], M=500, lmax=3, nmax=3, rc=6.0),
    "c3": dict(rep=29, Zs=[3, 15, 16, 8], M=2000, lmax=3, nmax=3, rc=6.0),
    "c4": dict(rep=17, Zs=[29], M=1000, lmax=6, nmax=8, rc=7.0),
    "c5": dict(rep=63, Zs=[29], M=4000, lmax=3, nmax=3, rc=6.0),
}
