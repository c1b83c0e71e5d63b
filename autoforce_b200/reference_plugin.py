"""B200ActiveCalculator -- the reference's own ``ActiveCalculator`` with its hot path on the GPU.

This module is the binding a reference maintainer adds: it imports the host
package ``theforce`` (AutoForce) and subclasses ``theforce.calculator.active.ActiveCalculator``
(calculator/active.py:104) so that every seam of SURVEY.md section 8(b) is served by libsgpr_b200:

  seam 1  ``TorchAtoms.update`` (descriptor/atoms.py:384-413): ``B200TorchAtoms`` keeps ``xyz`` / ``lll`` /
          ``indices`` / ``local(k)`` but takes its neighbour list from ``sgpr_neighbors`` on first use and
          builds the per-atom ``Local`` objects only if reference code actually touches ``atoms.loc``.
  seam 2/3 ``model.gp.kern(atoms, X)`` (similarity/similarity.py:17-31, regression/gppotential.py:47-84):
          while ``calculate`` runs, ``EnergyForceKernel.forward`` of the loaded model is routed to
          ``sgpr_kernel_forward`` / ``sgpr_kernel_backward`` -- a ``[N, M]`` tensor (or ``[N, 1]`` column for one new
          LCE) that back-propagates into ``atoms.xyz`` and ``atoms.lll`` exactly like the reference's autograd
          graph (calculator/active.py:464,587-599,867-868; gppotential.py:905-911).
  seam 4  ``calculate / update_results / grads / get_covloss`` (calculator/active.py:425-611,781-804): in
          prediction mode (``calculator=None``) one fused ``sgpr_predict_host`` call produces energy, forces,
          virial and the covloss beta; nothing of size N x M exists on the host.
  seam 5  the model itself stays the reference's ``PosteriorPotential`` (pickled folder, ``to_folder`` /
          ``PosteriorPotentialFromFolder``, gppotential.py:1073-1119,1342-1368): the device copy follows it through
          ``sgpr_append_inducing`` / ``sgpr_set_weights`` after ``add_inducing`` / ``make_munu``.

What stays on the reference path (north_star): the on-the-fly sampling logic (``update``, ``update_inducing``,
``update_data``, ``update_lce``), the ab initio calls, the M x M algebra and hyper-parameter optimisation, and the
training-time kernels ``Kf`` / ``Kv`` of sampled data (``SgprEngine.kernel_jacobian`` offers them on the GPU, but the
reference's hand-written ``leftgrad`` has an index-assignment defect in cells narrower than 2 rc -- DESIGN.md section 7 --
and a drop-in must not change what the trainer sees).

There is no CPU fallback for the hot path: without the CUDA library / a device the constructor of the engine raises.
"""
from __future__ import annotations

import time
import weakref

import ase
import numpy as np
import torch
from ase.calculators.calculator import all_changes
from theforce.calculator.active import ActiveCalculator
from theforce.descriptor.atoms import Local, TorchAtoms
from theforce.util.util import iterable

from .engine import SgprEngine
from .model import SgprModel

inf = float("inf")


class _Range:
    """NVTX range on the reference's timing nodes (calculator/active.py:427-535) + wall clock."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False


class _GpuNeighborList:
    """``ase.neighborlist.NeighborList`` as the reference uses it (descriptor/atoms.py:348-355,366,402: rc/2 radii,
    skin 0, both ways, no self interaction): ``update(atoms)`` only records the structure, the list itself comes
    from ``sgpr_neighbors`` when ``get_neighbors`` is first called."""

    def __init__(self, provider):
        self._provider = provider
        self._atoms = None
        self._csr = None

    def update(self, atoms):
        self._atoms = (np.array(atoms.positions, dtype=np.float64), np.array(atoms.numbers),
                       np.array(atoms.cell, dtype=np.float64).reshape(3, 3), np.array(atoms.pbc))
        self._csr = None
        return True

    def get_neighbors(self, a):
        if self._csr is None:
            if self._atoms is None:
                raise RuntimeError("neighbour list used before update()")
            first, j, S = self._provider(*self._atoms)
            self._csr = (np.asarray(first), np.asarray(j, dtype=np.int64), np.asarray(S, dtype=np.int64).reshape(-1, 3))
        first, j, S = self._csr
        sl = slice(int(first[a]), int(first[a + 1]))
        return j[sl].copy(), S[sl].copy()


class B200TorchAtoms(TorchAtoms):
    """``TorchAtoms`` (descriptor/atoms.py:262) whose expensive parts are lazy: the neighbour list is the GPU's and
    the ``Local`` objects (one reference descriptor evaluation each) exist only once somebody reads ``atoms.loc``
    -- the prediction path never does.  ``copy()`` is inherited: it returns a plain, eagerly staged ``TorchAtoms``
    (snapshots become training data and are pickled with the model, they stay pure reference objects)."""

    _nl_provider = None
    _lazy_args = None
    _loc = None

    @property
    def loc(self):
        if self._loc is None and self._lazy_args is not None:
            stage, dont_save_grads = self._lazy_args
            self._loc = [self.local(a, stage=stage, dont_save_grads=dont_save_grads) for a in self.indices]
        return self._loc

    @loc.setter
    def loc(self, value):
        self._loc = value

    def build_nl(self, rc):
        super().build_nl(rc)
        if self._nl_provider is not None:
            self.nl = _GpuNeighborList(self._nl_provider)

    def update(self, cutoff=None, descriptors=None, forced=False, build_locals=True, stage=True, posgrad=False,
               cellgrad=False, dont_save_grads=False):
        self._lazy_args = (stage, dont_save_grads) if build_locals else None
        super().update(cutoff=cutoff, descriptors=descriptors, forced=forced, build_locals=False, stage=stage,
                       posgrad=posgrad, cellgrad=cellgrad, dont_save_grads=dont_save_grads)

    def lazy_copy(self):
        """What ``Calculator.calculate`` stores in ``calc.atoms`` (``atoms.copy()``), without leaving this class."""
        new = B200TorchAtoms(positions=np.array(self.positions, dtype=np.float64), cell=np.array(self.cell, dtype=np.float64),
                             numbers=np.array(self.numbers), pbc=np.array(self.pbc), ranks=self.ranks)
        new._nl_provider = self._nl_provider
        vel = self.get_velocities()
        if vel is not None:
            new.set_velocities(np.array(vel))
        return new


def _safe_weights(model, M):
    """mu / choli of the reference model when they exist and match the inducing set (``make_munu`` returns early
    while there is no data, gppotential.py:548-550), else zeros / None."""
    mu = getattr(model, "mu", None)
    mu = mu.detach() if (mu is not None and mu.numel() == M) else torch.zeros(M)
    choli = getattr(model, "choli", None)
    choli = choli.detach() if (choli is not None and tuple(choli.shape) == (M, M)) else None
    return mu, choli


def _flat_models(model):
    """The reference model as flat SgprModel(s) (one per similarity kernel when they cannot be merged)."""
    import types

    M = len(model.X)
    mu, choli = _safe_weights(model, M)
    view = types.SimpleNamespace(gp=model.gp, X=list(model.X), mu=mu, choli=choli, mean=model.mean,
                                 _vscale=getattr(model, "_vscale", {}))
    return SgprModel.list_from_posterior_potential(view)


def _env_of(loc):
    return (int(loc.number), loc._r.detach().cpu().numpy().reshape(-1, 3), loc._b.detach().cpu().numpy().reshape(-1))


class B200ActiveCalculator(ActiveCalculator):
    """Drop-in for ``theforce.calculator.active.ActiveCalculator``: same constructor (+ ``device``), same results,
    log lines, model folder and tape; ``calculate`` runs on libsgpr_b200."""

    def __init__(self, *args, device=None, **kwargs):
        self._device = device
        self._engines = None          # one SgprEngine per flat model (kernel sums with different hyper-parameters: > 1)
        self._engine_locs = []        # the Local objects (strong references) the engines currently hold, in order
        self._engine_sig = None
        self._column_engines = []     # engines of single new LCEs, alive while this step's autograd graph is
        self._cov = None
        self._beta = None
        self._kern_patched = False
        self.to_ase = True
        super().__init__(*args, **kwargs)
        if self.process_group is not None:
            raise NotImplementedError("B200ActiveCalculator shards over GPUs with torch.distributed (NCCL), not over an "
                                      "MPI process group: launch one process per GPU and leave process_group=None")

    # ------------------------------------------------------------------ engines follow the reference model
    def _dev(self):
        if self._device is not None:
            return int(self._device)
        return torch.cuda.current_device() if torch.cuda.is_available() else 0

    def _signature(self):
        return tuple(k.state for k in self.model.gp.kern.kernels)

    def close(self):
        for e in (self._engines or []) + self._column_engines:
            e.close()
        self._engines, self._engine_locs, self._column_engines = None, [], []

    def _sync_engines(self, numbers, weights):
        """Make the device copy equal to ``self.model``: rebuild when the kernels / species table changed or inducing
        LCEs were removed, ``sgpr_append_inducing`` when LCEs were appended, ``sgpr_set_weights`` when only mu / choli /
        mean / vscale moved (``make_munu``)."""
        X = list(self.model.X)
        need = set(int(z) for z in np.unique(numbers))
        sig = self._signature()
        old = self._engine_locs
        prefix = len(old) <= len(X) and all(a is b for a, b in zip(old, X))
        rebuild = (self._engines is None or sig != self._engine_sig or not prefix
                   or not need.issubset(self._engines[0].species))
        if not rebuild and len(X) > len(old):
            new = X[len(old):]
            known = set(self._engines[0].species)
            if all(int(l.number) in known and set(int(z) for z in l._b.tolist()) <= known for l in new):
                mu, choli = _safe_weights(self.model, len(X))
                for e in self._engines:
                    e.append_inducing([_env_of(l) for l in new], mu.numpy(), None if choli is None else choli.numpy())
            else:
                rebuild = True
        if rebuild:
            for e in self._engines or []:
                e.close()
            models = _flat_models(self.model)
            species = sorted(need | set(z for m in models for z in m.species()))
            self._engines = [SgprEngine(m, species=species, device=self._dev()) for m in models]
            self._engine_sig = sig
        self._engine_locs = X
        if weights and len(X) > 0:
            mu, choli = _safe_weights(self.model, len(X))
            mean = self.model.mean
            mean_w = {int(z): float(w) + float(getattr(mean, "_weights", {}).get(z, 0.0))
                      for z, w in getattr(mean, "weights", {}).items()}
            vscale = {int(z): float(v) for z, v in getattr(self.model, "_vscale", {}).items()}
            lead = max(range(len(self._engines)), key=lambda i: self._engines[i].model.rc)
            for i, e in enumerate(self._engines):
                m = e.model
                mw = mean_w if i == lead else {}
                same = (np.array_equal(m.mu, mu.numpy()) and m.mean_w == mw and m.vscale == vscale
                        and ((m.choli is None) == (choli is None)) and (choli is None or np.array_equal(m.choli, choli.numpy())))
                if not same:
                    e.set_weights(mu=mu.numpy(), mean_w=mw, choli=None if choli is None else choli.numpy(), vscale=vscale)

    def _neighbors(self, pos, numbers, cell, pbc):
        return self._engines[0].neighbors(pos, numbers, cell, pbc)

    # ------------------------------------------------------------------ seam 2/3: model.gp.kern(atoms, X) on the GPU
    def _gpu_cov(self, second):
        """``kern(self.atoms, second)`` -> differentiable ``[N, len(second)]`` (similarity.py:17-31 with operation
        "func", summed over the kernels, gppotential.py:81-84)."""
        a = self.atoms
        numbers, pbc = np.asarray(a.numbers), np.asarray(a.pbc)
        if second is self.model.X:
            engines = self._engines
            if len(self._engine_locs) != len(self.model.X):
                self._sync_engines(numbers, weights=False)
                engines = self._engines
        else:
            locs = list(iterable(second))
            if not all(isinstance(l, Local) for l in locs):
                raise TypeError("second must be Local / LocalsData")
            view_models = _flat_models(_ModelView(self.model, locs))
            species = self._engines[0].species
            engines = [SgprEngine(m, species=species, device=self._dev()) for m in view_models]
            self._column_engines += engines
        if len(engines[0].model.ind_Z) == 0:
            return torch.zeros(len(numbers), 0)
        K = None
        for e in engines:
            k = e.cov(a.xyz, a.lll, numbers, pbc)
            K = k if K is None else K + k
        return K

    def _install_kern(self):
        kern = self.model.gp.kern
        if "forward" in kern.__dict__:
            return False
        me = weakref.ref(self)
        reference_forward = kern.forward

        def forward(first, second=None, cov="energy_energy", inducing=None):
            calc = me()
            if (calc is not None and cov == "energy_energy" and inducing is None and second is not None
                    and first is calc.atoms and isinstance(first, B200TorchAtoms)):
                return calc._gpu_cov(second)
            return reference_forward(first, second, cov=cov, inducing=inducing)

        kern.__dict__["forward"] = forward
        return True

    def _uninstall_kern(self):
        return self.model.gp.kern.__dict__.pop("forward", None) is not None

    def save_model(self):
        # the pickled folder must hold pure reference objects (gppotential.py:1060-1119)
        was = self._uninstall_kern()
        try:
            super().save_model()
        finally:
            if was:
                self._install_kern()

    # ------------------------------------------------------------------ self.cov: materialised on demand
    @property
    def cov(self):
        if self._cov is None and self.atoms is not None and self._engines is not None:
            if getattr(self.atoms, "xyz", None) is None or getattr(self.atoms, "cutoff", None) is None:
                self.atoms.update(posgrad=True, cellgrad=True, forced=True, dont_save_grads=True,
                                  cutoff=self.model.cutoff, descriptors=self.model.gp.kern.kernels)
            self._cov = self._gpu_cov(self.model.X)
        return self._cov

    @cov.setter
    def cov(self, value):
        self._cov = value

    def get_covloss(self):
        """calculator/active.py:781-804; in prediction mode the fused call already produced beta on the device."""
        if self._beta is not None:
            return torch.as_tensor(self._beta)
        return super().get_covloss()

    # ------------------------------------------------------------------ the call
    def _wrap(self, _atoms):
        if isinstance(_atoms, TorchAtoms):
            if not isinstance(_atoms, B200TorchAtoms):
                _atoms.__class__ = B200TorchAtoms
            self.to_ase = False
            return _atoms
        self.to_ase = True
        return B200TorchAtoms(ase_atoms=_atoms, ranks=self.distrib)

    def _fusable(self):
        if self.active or self.meta is not None or len(self.model.X) == 0:
            return False
        kerns = self.model.gp.kern.kernels
        normalized = all(bool(getattr(getattr(k, "descriptor", None), "normalize", True)) for k in kerns)
        return normalized and getattr(self.model, "choli", None) is not None and self.normalized is not False

    def calculate(self, _atoms=None, properties=["energy"], system_changes=all_changes):
        timings = [time.time()]                                   # node 0 (active.py:427)
        if self.size[1] == 0 and not self.active:
            raise RuntimeError("you forgot to assign a DFT calculator!")
        if _atoms is None:
            _atoms = self.atoms
        wrapper = self._wrap(_atoms)
        wrapper._nl_provider = self._neighbors
        self.atoms = wrapper.lazy_copy()
        for e in self._column_engines:
            e.close()
        self._column_engines = []
        self._cov, self._beta = None, None
        fused = self._fusable()
        self._sync_engines(self.atoms.numbers, weights=fused)
        if fused and len(self._engines) == 1:
            self._calculate_fused(timings)
        else:
            self._calculate_differentiable(wrapper, timings)

    def _calculate_fused(self, timings):
        a = self.atoms
        eng = self._engines[0]
        with _Range("sgpr:nl+desc+kernel+results+covloss (fused sgpr_predict_host)"):
            E, F, W, _, beta = eng.predict(a.positions, a.numbers, np.asarray(a.cell, dtype=np.float64).reshape(3, 3), a.pbc,
                                           want_beta=True)
        timings += [time.time()] * 3                              # nodes 1-3 are one device call here
        try:
            volume = a.get_volume()
        except ValueError:
            volume = -2                                           # active.py:606-609
        self.results["energy"] = np.array(E, dtype=np.float64)
        self.results["forces"] = F
        self.results["stress"] = (W / volume).flat[[0, 4, 8, 5, 2, 1]]
        self.maximum_force = abs(F).max() if F.size else 0.0
        self.deltas = None
        self._beta = beta
        if self.normalized is None:
            self.normalized = True
            self.log(f"kernel normalization status {self.normalized}")
        covloss_max = float(torch.as_tensor(beta).max()) if len(beta) else 0.0
        self.covlog = f"{covloss_max}"
        if covloss_max > self.ediff:                              # active.py:495-499
            tmp = a.as_ase()
            tmp.calc = None
            if self.rank == 0:
                ase.io.Trajectory("active_uncertain.traj", "a").write(tmp)
        timings.append(time.time())                               # node 4
        self.post_calculate(timings)

    def _calculate_differentiable(self, wrapper, timings):
        uargs = {"cutoff": self.model.cutoff, "descriptors": self.model.gp.kern.kernels}
        dat1 = self.size[0]
        with _Range("sgpr:nl+desc"):
            self.atoms.update(posgrad=True, cellgrad=True, forced=True, dont_save_grads=True, **uargs)
        timings.append(time.time())                               # node 1
        self.maximum_force = inf
        self._install_kern()
        try:
            if self.step == 0 and self.active and self.model.ndata == 0:
                self.initiate_model()
                self._update_args = dict(data=False)
            with _Range("sgpr:kernel"):
                self.cov = self.model.gp.kern(self.atoms, self.model.X)
            timings.append(time.time())                           # node 2
            with _Range("sgpr:results"):
                self.update_results(self.active or (self.meta is not None))
            timings.append(time.time())                           # node 3
            with _Range("sgpr:active"):
                self._active_node(wrapper, dat1)
        finally:
            self._uninstall_kern()
        timings.append(time.time())                               # node 4
        self.post_calculate(timings)

    def _active_node(self, wrapper, dat1):
        """The sampling step of calculator/active.py:473-500, on the reference's own methods."""
        self.deltas = None
        self.covlog = ""
        if not self.active or self.veto():
            covloss_max = float(self.get_covloss().max())
            self.covlog = f"{covloss_max}"
            if covloss_max > self.ediff:
                tmp = self.atoms.as_ase()
                tmp.calc = None
                if self.rank == 0:
                    ase.io.Trajectory("active_uncertain.traj", "a").write(tmp)
            return
        first_bead = self.nbeads == 1 or (self.step + 1) % self.nbeads == 1
        if first_bead:                                            # PIMD: only the first bead is sampled
            before = self.results.copy()
            m, n = self.update(**self._update_args)
            if n > 0 or m > 0:
                self.update_results(self.meta is not None)
                if self.step > 0:
                    self.deltas = {q: self.results[q] - before[q] for q in ("energy", "forces", "stress")}
        if self.size[0] == dat1:
            self.distrib.unload(wrapper)

    # rank / world of the reference (theforce.distributed): single process per GPU here
    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _ModelView:
    """A reference model restricted to the given inducing LCEs (for ``kern(atoms, loc)`` columns)."""

    def __init__(self, model, locs):
        self.gp = model.gp
        self.X = locs
        self.mean = type("NoMean", (), {"weights": {}, "_weights": {}})()
        self._vscale = {}
