"""Config 1 of BASELINE.json as a regression harness: bulk Cu (108 atoms) on-the-fly NVT MD with an
analytic pair surrogate as the "ab initio" calculator.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

The reference's template is ``examples/active/md.py`` (ASE EMT behind a socket + ASE NPT); neither ASE's
EMT nor its MD integrators exist in this image, so -- as SURVEY.md section 8(c) prescribes -- the
surrogate is a small analytic pair potential (Morse, Girifalco-Weizer Cu parameters, smoothly shifted to
zero at 6 A) and the integrator a ~20-line Langevin/velocity-Verlet loop.  The same harness drives the
unmodified reference ``ActiveCalculator`` and the B200 plugin, so the two runs can be compared step by
step (energies, forces, stress, sampled LCE indices, model sizes).
"""
from __future__ import annotations

import numpy as np

FS = 0.09822694788464063      # ase.units.fs
KB = 8.617330337217213e-05    # ase.units.kB
MASS = {29: 63.546, 3: 6.94, 15: 30.974, 16: 32.06, 8: 15.999, 1: 1.008}


def make_surrogate():
    """ASE-protocol calculator class (built lazily so that `ase` -- real or shim -- is imported by the caller)."""
    from ase.calculators.calculator import Calculator, all_changes

    from .sgpr_oracle import complete_cell, neighbor_list

    class MorseSurrogate(Calculator):
        implemented_properties = ["energy", "forces", "stress", "free_energy"]

        def __init__(self, D=0.3429, alpha=1.3588, r0=2.866, rc=6.0):
            Calculator.__init__(self)
            self.D, self.alpha, self.r0, self.rc = D, alpha, r0, rc
            self.calls = 0

        def _phi(self, r):
            x = np.exp(-self.alpha * (r - self.r0))
            return self.D * (x * x - 2 * x), self.D * (-2 * self.alpha) * (x * x - x)

        def calculate(self, atoms=None, properties=["energy"], system_changes=all_changes):
            Calculator.calculate(self, atoms, properties, system_changes)
            a = atoms if atoms is not None else self.atoms
            pos = np.asarray(a.positions, dtype=float)
            cell = complete_cell(np.asarray(a.cell, dtype=float).reshape(3, 3))
            first, J, S = neighbor_list(pos, cell, np.asarray(a.pbc), self.rc)
            I = np.repeat(np.arange(len(pos)), np.diff(first))
            rij = pos[J] - pos[I] + S @ cell
            r = np.linalg.norm(rij, axis=1)
            v, dv = self._phi(r)
            vc, dvc = self._phi(np.array([self.rc]))
            v = v - vc - dvc * (r - self.rc)          # value and slope vanish at the cutoff
            dv = dv - dvc
            g = (dv / r)[:, None] * rij               # d phi / d r_ij
            F = np.zeros_like(pos)
            np.add.at(F, I, 0.5 * g)
            np.add.at(F, J, -0.5 * g)
            W = 0.5 * np.einsum("ka,kb->ab", rij, g)  # both directions are listed: half of each
            try:
                vol = a.get_volume()
            except ValueError:
                vol = 1.0
            self.results = {"energy": float(0.5 * v.sum()), "forces": F, "stress": (W / vol).reshape(-1)[[0, 4, 8, 5, 2, 1]]}
            self.results["free_energy"] = self.results["energy"]
            self.calls += 1

    return MorseSurrogate


def cu108(sigma=0.1, seed=7):
    a0 = 3.61
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a0
    grid = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], dtype=float) * a0
    pos = (grid[:, None, :] + base[None, :, :]).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    pos = pos + rng.normal(0, sigma, pos.shape)
    return pos, np.eye(3) * 3 * a0, np.full(len(pos), 29)


def run_md(calc, steps, dt_fs=3.0, T=600.0, friction=0.02, seed=11, sigma=0.1, numpy_seed=1234, kick_steps=(), kick=0.7):
    """Langevin MD (BAOAB-free, plain velocity Verlet + friction/noise kick) of the rattled 108-atom Cu cell with
    `calc` attached.  Returns one record per force evaluation.  ``kick_steps``: before these evaluations three atoms
    are displaced by ``kick`` A in seeded random directions (novel environments -> the learner samples again)."""
    import ase

    np.random.seed(numpy_seed)   # the reference draws from numpy's global RNG (sample_rand_lces, index_distribute)
    pos, cell, numbers = cu108(sigma, seed)
    atoms = ase.Atoms(positions=pos, cell=cell, numbers=numbers, pbc=True)
    atoms.calc = calc
    rng = np.random.default_rng(seed + 1)
    m = np.array([MASS[int(z)] for z in numbers])[:, None]
    v = rng.normal(0, 1, pos.shape) * np.sqrt(KB * T / m)
    dt = dt_fs * FS
    records = []

    def evaluate():
        e = float(atoms.get_potential_energy())
        f = np.array(atoms.get_forces(), dtype=float)
        s = np.array(atoms.get_stress(), dtype=float)
        model = calc.model
        records.append(dict(energy=e, forces=f.copy(), stress=s.copy(), positions=np.array(atoms.positions).copy(),
                            ndata=int(model.ndata), ninducing=len(model.X), lce_index=[int(x.index) for x in model.X],
                            covlog=str(getattr(calc, "covlog", ""))))
        return f

    f = evaluate()
    for k in range(1, steps):
        v = v + 0.5 * dt * f / m
        new = np.array(atoms.positions) + dt * v
        if k in kick_steps:
            who = rng.choice(len(new), 3, replace=False)
            d = rng.normal(0, 1, (3, 3))
            new[who] += kick * d / np.linalg.norm(d, axis=1)[:, None]
        atoms.set_positions(new)
        f = evaluate()
        v = v + 0.5 * dt * f / m
        c = np.exp(-friction * dt_fs)
        v = c * v + np.sqrt((1 - c * c) * KB * T / m) * rng.normal(0, 1, pos.shape)
    return records
