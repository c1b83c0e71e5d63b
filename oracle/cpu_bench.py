"""CPU timing of the oracle port on a bounded sample of a workload (bench.py's
``cpu_baseline`` leg and ``--impl reference`` arm).  TEST/BENCH INFRASTRUCTURE ONLY.

What is timed per step: for ``sample`` atoms of the structure (full neighbour
environments, full inducing set) -- neighbour search for those centres, descriptors,
kernel rows, energies and the analytic force/virial back-projection -- spread over
``procs`` worker processes (the reference's own decomposition is over atoms,
descriptor/atoms.py:228-259,321-341).  The inducing descriptors are evaluated once
outside the timed region (the reference caches them too, universal.py:100-107).
"""
from __future__ import annotations

import os
import time
from multiprocessing import get_context

import numpy as np

from . import sgpr_oracle as o

_G = {}


def _work(args):
    k, idx = args
    g = _G
    os.environ["OMP_NUM_THREADS"] = "1"
    out = o.predict(g["model"], g["pos"][k], g["cell"], True, g["numbers"], atoms=idx, Zh=g["Zh"], chunk=64)
    return out["e_local"].sum(), np.abs(out["forces"]).max()


def to_oracle_model(model):
    first = model.ind_first
    return o.OracleModel(
        lmax=model.lmax, nmax=model.nmax, xi=model.xi, rc=model.rc, radii=model.radii, default_radius=model.default_radius,
        ind_Z=model.ind_Z.astype(np.int64), ind_r=[model.ind_r[first[m]:first[m + 1]] for m in range(model.M)],
        ind_b=[model.ind_b[first[m]:first[m + 1]].astype(np.int64) for m in range(model.M)], mu=model.mu,
        mean_w=model.mean_w, choli=model.choli, vscale=model.vscale, a_not=model.a_not)


class CpuBench:
    def __init__(self, model, pos_variants, cell, numbers, sample, procs=None, seed=0):
        self.procs = procs or min(os.cpu_count() or 1, 64)
        self.sample = int(min(sample, len(numbers)))
        rng = np.random.default_rng(seed)
        self.idx = np.sort(rng.choice(len(numbers), self.sample, replace=False))
        om = to_oracle_model(model)
        species = om.species_table(extra=numbers)
        _G.update(model=om, pos=[np.asarray(p) for p in pos_variants], cell=np.asarray(cell), numbers=np.asarray(numbers, dtype=np.int64),
                  Zh=o.inducing_descriptors(om, species))
        self.chunks = [c for c in np.array_split(self.idx, self.procs) if len(c)]
        self.pool = get_context("fork").Pool(len(self.chunks)) if len(self.chunks) > 1 else None
        self.nvar = len(pos_variants)

    def step(self, it):
        jobs = [(it % self.nvar, c) for c in self.chunks]
        res = self.pool.map(_work, jobs) if self.pool else [_work(j) for j in jobs]
        return sum(r[0] for r in res)

    def run(self, steps, warmup=0):
        for it in range(warmup):
            self.step(it)
        t0 = time.perf_counter()
        for it in range(steps):
            self.step(it)
        dt = time.perf_counter() - t0
        return self.sample * steps / dt, dt / steps

    def close(self):
        if self.pool:
            self.pool.terminate()
            self.pool = None
