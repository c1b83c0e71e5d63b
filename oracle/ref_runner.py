"""Run the UNMODIFIED reference (`/root/reference/theforce`) in this container.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py): used by
``tests/golden/make_golden.py`` to generate golden vectors, by the tests that run the
reference side by side with the product, and by ``bench.py --impl reference``.  The
package is imported from /root/reference (build container) or from the unmodified copy
staged under ``oracle/_ref`` by ``oracle/stage_ref.py`` (GPU box).  Follows SURVEY.md
Appendix B.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("AUTOFORCE_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "theforce")):
    # outside the build container (GPU box): the copy staged by oracle/stage_ref.py (git-ignored oracle/_ref/)
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")
_SHIMS = os.path.join(_HERE, "shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "theforce"))


def reference_kind():
    """"source" = /root/reference itself, "staged" = the unmodified copy under oracle/_ref."""
    return "staged" if REFERENCE_ROOT.endswith("_ref") else "source"


def import_reference():
    """Put the shims + the reference on sys.path and import theforce.
    NOTE: importing theforce sets torch's default dtype to float64
    (theforce/__init__.py:13)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo, REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if not hasattr(np, "int"):
        np.int = int  # descriptor/atoms.py:157-158 still uses np.int
    import theforce  # noqa: F401

    return theforce


def make_kernel(kind, lmax, nmax, xi, rc, radii=None, atomic_unit=None, a_not=(), normalize=True, a=None):
    import_reference()
    if kind == "sesoap":
        from theforce.descriptor.sesoap import DefaultRadii, SpecialRadii
        from theforce.similarity.sesoap import SeSoapKernel
        from theforce.util.util import EqAll

        if radii is None:
            rad = DefaultRadii()
        else:
            rad = SpecialRadii({int(k): float(v) for k, v in radii.items() if k != "others"}, float(radii.get("others", 1.0)))
        if a is None:
            a = EqAll(list(a_not)) if len(a_not) else None   # else: a fixed central species (similarity/universal.py:101)
        return SeSoapKernel(lmax, nmax, xi, float(rc), a=a, radii=rad, normalize=normalize)
    elif kind == "subsesoap":
        # default_kernel(species=...) of calculator/active.py:28-38
        from theforce.descriptor.sesoap import DefaultRadii
        from theforce.similarity.sesoap import SubSeSoapKernel

        species = list(radii["species"])
        return [SubSeSoapKernel(lmax, nmax, xi, float(rc), z, species, radii=DefaultRadii()) for z in species]
    elif kind == "heterosoap":
        # one HeterogeneousSoapKernel per central species over NormalizedSoap(HeteroSoap) (similarity/heterosoap.py:10-29)
        from theforce.descriptor.cutoff import PolyCut
        from theforce.regression.kernel import DotProd
        from theforce.similarity.heterosoap import HeterogeneousSoapKernel

        species = list(radii["species"])
        return [HeterogeneousSoapKernel(DotProd() ** xi, z, species, lmax, nmax, PolyCut(float(rc)), atomic_unit=atomic_unit)
                for z in species]
    elif kind == "multi":
        # a kernel list with DIFFERENT hyper-parameters: EnergyForceKernel sums them (regression/gppotential.py:81-84)
        return [make_kernel(**k) for k in radii["kernels"]]
    elif kind == "universal":
        from theforce.similarity.universal import UniversalSoapKernel

        return UniversalSoapKernel(lmax, nmax, xi, float(rc), atomic_unit=atomic_unit, a_not=list(a_not), normalize=normalize, a=a)
    raise ValueError(kind)


def ase_atoms(pos, cell, pbc, numbers):
    import_reference()
    from ase.atoms import Atoms

    return Atoms(positions=np.array(pos), cell=np.array(cell), pbc=pbc, numbers=np.array(numbers))


def synth_model(kernel, inducing_envs, mu, mean_w, choli, vscale):
    """Frozen PosteriorPotential without training.
    inducing_envs: list of (Z, r[nn,3], b[nn])."""
    import_reference()
    import torch
    from theforce.descriptor.atoms import Local, LocalsData
    from theforce.regression.gppotential import AutoMean, PosteriorPotential

    model = PosteriorPotential(kernel)
    locs = []
    for Z, r, b in inducing_envs:
        nn = len(b)
        i = np.zeros(nn, dtype=np.int64)
        j = np.arange(1, nn + 1, dtype=np.int64)
        loc = Local(i, j, int(Z), np.asarray(b, dtype=np.int64), torch.as_tensor(np.asarray(r, dtype=float).reshape(nn, 3)))
        loc.stage(model.descriptors, dont_save_grads=True)
        locs.append(loc)
    model.X = LocalsData(locs)
    model.mu = torch.as_tensor(np.asarray(mu, dtype=float))
    model.choli = torch.as_tensor(np.asarray(choli, dtype=float))
    model.Mi = model.choli.t() @ model.choli
    model.ridge = torch.zeros([])
    model._vscale = {int(z): torch.tensor(float(v)) for z, v in vscale.items()}
    mean = AutoMean()
    mean.weights = {int(z): torch.tensor(float(w)) for z, w in mean_w.items()}
    mean._weights = {int(z): 0.0 for z in mean_w}
    model.mean = mean
    return model


def ref_predict(model, pos, cell, pbc, numbers, want_descriptors=()):
    """One ``ActiveCalculator.calculate`` in prediction mode; returns the
    reference's results plus the kernel matrix, covloss and neighbour list."""
    import_reference()
    from theforce.calculator.active import ActiveCalculator

    calc = ActiveCalculator(covariance=model, calculator=None, pckl=None, tape=None, logfile=None)
    atoms = ase_atoms(pos, cell, pbc, numbers)
    atoms.calc = calc
    cwd = os.getcwd()
    os.chdir("/tmp")  # active_uncertain.traj etc. must not land in the repo
    try:
        e = atoms.get_potential_energy()
        f = atoms.get_forces()
        s = atoms.get_stress()
    finally:
        os.chdir(cwd)
    out = dict(
        energy=np.array(e, dtype=float),
        forces=np.array(f, dtype=float),
        stress=np.array(s, dtype=float),
        K=calc.cov.detach().numpy().copy(),
        covloss=calc.get_covloss().detach().numpy().copy(),
    )
    ta = calc.atoms
    first = [0]
    J, S = [], []
    for a in range(len(numbers)):
        j, off = ta.nl.get_neighbors(a)
        J.append(j)
        S.append(off)
        first.append(first[-1] + len(j))
    out["nl_first"] = np.array(first, dtype=np.int64)
    out["nl_j"] = np.concatenate(J).astype(np.int64) if J else np.zeros(0, np.int64)
    out["nl_S"] = np.concatenate(S).astype(np.int64).reshape(-1, 3) if S else np.zeros((0, 3), np.int64)
    # dense copies of a few cached descriptors  (loc.kern_0_value, sparse [120,120,D])
    desc = {}
    for a in want_descriptors:
        v = ta.loc[a].__dict__.get("kern_0_value")
        desc[int(a)] = None if (v is None or not v.is_sparse) else v.detach().to_dense().numpy().copy()
    out["descriptors"] = desc
    return out
