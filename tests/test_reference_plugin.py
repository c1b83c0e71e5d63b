"""B200ActiveCalculator inside the UNMODIFIED reference.

The reference package is imported from /root/reference (build container) or from the copy staged under
oracle/_ref by oracle/stage_ref.py (GPU box) together with the minimal ase / mpi4py stand-ins of oracle/shims.
Every test runs the reference's own ``ActiveCalculator`` and the plugin side by side on identical inputs:

* on-the-fly training (BASELINE.json config 1: Cu-108, SeSoapKernel(3,3,4,6.0), analytic pair surrogate as the ab
  initio calculator, Langevin MD with two "kicks" so that the learner samples again at later steps): same sampled LCE
  indices, same model sizes, |dE|/N < 1e-6 eV, |dF| < 1e-5 eV/A, |d stress| < 1e-6 eV/A^3 at every step;
* the model folder the plugin writes (``to_folder``) is a pure reference pickle: it loads with
  ``PosteriorPotentialFromFolder`` and flattens to the same ``SgprModel`` as the in-memory model;
* prediction mode (``calculator=None``) from that folder: results, covloss and log line equal the reference's.

`-m gpu`: the real libsgpr_b200 engine.  Without a GPU the same tests run with tests/oracle_engine.py (numpy oracle)
in place of the engine, which checks the host logic of the plugin.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_runner  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_runner.reference_available(), reason="reference package not available (oracle/_ref not staged)")

STEPS, KICKS = 6, (2, 4)
SETTINGS = dict(ediff=0.01, fdiff=3e-4)


def _kernel():
    ref_runner.import_reference()
    from theforce.descriptor.sesoap import DefaultRadii
    from theforce.similarity.sesoap import SeSoapKernel

    return SeSoapKernel(3, 3, 4, 6.0, radii=DefaultRadii())


def _run(tmp, which, engine_cls=None):
    """One on-the-fly run in its own directory.  which: 'reference' | 'plugin'."""
    ref_runner.import_reference()
    from oracle.onthefly import make_surrogate, run_md

    os.makedirs(tmp, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        kw = dict(covariance=_kernel(), calculator=make_surrogate()(), pckl="model.pckl", tape="model.sgpr", logfile="active.log",
                  **SETTINGS)
        if which == "reference":
            from theforce.calculator.active import ActiveCalculator

            calc = ActiveCalculator(**kw)
        else:
            import autoforce_b200.reference_plugin as rp

            if engine_cls is not None:
                rp.SgprEngine = engine_cls
            calc = rp.B200ActiveCalculator(**kw)
        rec = run_md(calc, STEPS, kick_steps=KICKS)
        return calc, rec
    finally:
        os.chdir(cwd)


def _compare_runs(ref, plug):
    assert len(ref) == len(plug) == STEPS
    N = len(ref[0]["forces"])
    sampled_later = False
    for k, (a, b) in enumerate(zip(ref, plug)):
        assert a["lce_index"] == b["lce_index"], f"step {k}: sampled LCEs differ"
        assert a["ndata"] == b["ndata"] and a["ninducing"] == b["ninducing"]
        assert np.abs(a["positions"] - b["positions"]).max() < 1e-7
        assert abs(a["energy"] - b["energy"]) / N < 1e-6, f"step {k}"
        assert np.abs(a["forces"] - b["forces"]).max() < 1e-5, f"step {k}"
        assert np.abs(a["stress"] - b["stress"]).max() < 1e-6, f"step {k}"
        assert abs(float(a["covlog"]) - float(b["covlog"])) < 1e-6
        if k > 0 and a["ninducing"] > ref[k - 1]["ninducing"]:
            sampled_later = True
    assert sampled_later, "the harness must trigger sampling after step 0"
    assert ref[-1]["ndata"] > 1, "the harness must add ab initio data after step 0"


def _engine_for_cpu():
    from oracle_engine import OracleEngine

    return OracleEngine


@pytest.fixture(scope="module")
def runs_cpu(tmp_path_factory):
    base = str(tmp_path_factory.mktemp("otf_cpu"))
    calc_r, ref = _run(os.path.join(base, "ref"), "reference")
    calc_p, plug = _run(os.path.join(base, "plug"), "plugin", _engine_for_cpu())
    return base, calc_r, ref, calc_p, plug


def test_onthefly_host_logic_matches_reference(runs_cpu):
    _, _, ref, calc_p, plug = runs_cpu
    _compare_runs(ref, plug)
    # the hot path never built the per-atom Local objects except where the reference's sampler asked for them
    assert type(calc_p.atoms).__name__ == "B200TorchAtoms"
    assert "forward" not in calc_p.model.gp.kern.__dict__, "the kern dispatch must be uninstalled after calculate()"


def _check_folder_and_prediction(base, calc_r, calc_p, engine_cls):
    """model folder written by the plugin == pure reference pickle; prediction mode from it == reference."""
    ref_runner.import_reference()
    import autoforce_b200.reference_plugin as rp
    from autoforce_b200 import SgprModel
    from oracle.onthefly import cu108
    from theforce.calculator.active import ActiveCalculator
    from theforce.regression.gppotential import PosteriorPotentialFromFolder

    folder = os.path.join(base, "plug", "model.pckl")
    assert all(os.path.isfile(os.path.join(folder, f)) for f in ("cutoff", "gp", "model", "info", "stats"))
    # no trace of the plugin in the pickle: it loads with the reference alone and holds reference classes only
    loaded = PosteriorPotentialFromFolder(folder, load_data=False, update_data=False)
    assert type(loaded).__module__.startswith("theforce.") and "forward" not in loaded.gp.kern.__dict__
    raw = open(os.path.join(folder, "model"), "rb").read()
    assert b"autoforce_b200" not in raw and b"reference_plugin" not in raw
    flat_disk = SgprModel.from_posterior_potential(loaded)
    flat_mem = SgprModel.from_posterior_potential(calc_p.model)
    for f in ("lmax", "nmax", "xi", "rc", "kind", "normalize", "radii", "default_radius", "a_not", "mean_w", "vscale"):
        assert getattr(flat_disk, f) == getattr(flat_mem, f), f
    for f in ("ind_Z", "ind_first", "ind_r", "ind_b", "mu", "choli"):
        assert np.array_equal(getattr(flat_disk, f), getattr(flat_mem, f)), f
    # the tape holds the same inducing LCEs (io/sgprio.py) in the same order, to the 8 decimals it prints
    from autoforce_b200.sgprio import read_lces

    tape = read_lces(os.path.join(base, "plug", "model.sgpr"))
    assert [t[0] for t in tape] == [int(z) for z in flat_mem.ind_Z]
    assert np.abs(np.concatenate([t[1] for t in tape]) - flat_mem.ind_r).max() < 1e-8
    # the reference's own folder (pure-reference run) gives the same flat model up to the round-off of its K
    flat_ref = SgprModel.from_posterior_potential(calc_r.model)
    assert np.array_equal(flat_ref.ind_Z, flat_mem.ind_Z) and np.array_equal(flat_ref.ind_first, flat_mem.ind_first)
    for m in range(flat_ref.M):   # neighbour order inside an LCE is implementation-defined: compare as sets
        sl = slice(flat_ref.ind_first[m], flat_ref.ind_first[m + 1])
        ra, rb = flat_ref.ind_r[sl], flat_mem.ind_r[sl]
        ra, rb = ra[np.lexsort(np.round(ra, 6).T)], rb[np.lexsort(np.round(rb, 6).T)]
        assert np.abs(ra - rb).max() < 1e-9
    assert np.abs(flat_ref.mu - flat_mem.mu).max() < 1e-6 * max(1.0, np.abs(flat_ref.mu).max())

    # ---- prediction mode from the folder: plugin vs reference
    cwd = os.getcwd()
    os.chdir(os.path.join(base, "plug"))
    try:
        if engine_cls is not None:
            rp.SgprEngine = engine_cls
        cp = rp.B200ActiveCalculator(covariance=folder, calculator=None, pckl=None, tape=None, logfile="pred_plug.log")
        cr = ActiveCalculator(covariance=folder, calculator=None, pckl=None, tape=None, logfile="pred_ref.log")
        import ase

        for seed in (3, 4):
            pos, cell, numbers = cu108(0.12, seed)
            out = []
            for c in (cr, cp):
                atoms = ase.Atoms(positions=pos, cell=cell, numbers=numbers, pbc=True)
                atoms.calc = c
                out.append((float(atoms.get_potential_energy()), np.array(atoms.get_forces()), np.array(atoms.get_stress()),
                            c.get_covloss().detach().numpy().copy(), c.covlog, dict(c.results)))
            (e0, f0, s0, b0, l0, r0), (e1, f1, s1, b1, l1, r1) = out
            assert abs(e0 - e1) / len(numbers) < 1e-6
            assert np.abs(f0 - f1).max() < 1e-5 and np.abs(s0 - s1).max() < 1e-6
            assert np.abs(b0 - b1).max() < 1e-6 and abs(float(l0) - float(l1)) < 1e-6
            assert r1["energy"].shape == () and r1["energy"].dtype == np.float64 and r1["forces"].shape == (len(numbers), 3)
            assert r1["stress"].shape == (6,) and "free_energy" in r1
        assert cp.step == cr.step == 2
        # same log line format: "<date> <time> <step> <energy> <temperature> <covloss> "
        lp = open("pred_plug.log").read().strip().splitlines()[-1].split()
        lr = open("pred_ref.log").read().strip().splitlines()[-1].split()
        assert len(lp) == len(lr) and lp[2] == lr[2] and abs(float(lp[3]) - float(lr[3])) < 1e-6 * len(numbers)
        assert cp._cov is None, "prediction mode must not materialise the N x M kernel matrix"
        cov = cp.cov   # ... but hands it out on request (calculator/active.py:464), differentiable
        assert tuple(cov.shape) == (len(numbers), len(cp.model.X)) and cov.requires_grad
        assert np.abs(cov.detach().numpy() - cr.cov.detach().numpy()).max() < 1e-9
        cp.close()
    finally:
        os.chdir(cwd)


def test_folder_roundtrip_and_prediction_host_logic(runs_cpu):
    base, calc_r, _, calc_p, _ = runs_cpu
    _check_folder_and_prediction(base, calc_r, calc_p, _engine_for_cpu())


# ------------------------------------------------------------------------------------ the same on the B200
@pytest.fixture(scope="module")
def runs_gpu(tmp_path_factory):
    ref_runner.import_reference()   # puts the reference + the ase / mpi4py stand-ins on sys.path
    import autoforce_b200.reference_plugin as rp
    from autoforce_b200.engine import SgprEngine

    rp.SgprEngine = SgprEngine
    base = str(tmp_path_factory.mktemp("otf_gpu"))
    calc_r, ref = _run(os.path.join(base, "ref"), "reference")
    calc_p, plug = _run(os.path.join(base, "plug"), "plugin", SgprEngine)
    return base, calc_r, ref, calc_p, plug


@pytest.mark.gpu
def test_onthefly_config1_matches_reference_gpu(runs_gpu):
    _, _, ref, calc_p, plug = runs_gpu
    _compare_runs(ref, plug)
    eng = calc_p._engines[0]
    assert type(eng).__module__ == "autoforce_b200.engine" and eng.stats()["kernel_launches"] > 0


@pytest.mark.gpu
def test_folder_roundtrip_and_prediction_gpu(runs_gpu):
    ref_runner.import_reference()
    from autoforce_b200.engine import SgprEngine

    base, calc_r, _, calc_p, _ = runs_gpu
    _check_folder_and_prediction(base, calc_r, calc_p, SgprEngine)
