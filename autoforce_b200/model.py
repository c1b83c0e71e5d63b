"""Flat, ASE-free description of a frozen SGPR model.

Holds exactly what the prediction path reads from the reference's PosteriorPotential
(theforce/regression/gppotential.py:453-478,548-649):
kernel hyper-parameters, the inducing LCEs (number, _r, _b), mu, choli, the constant
mean and the variance scales.  ``from_posterior_potential`` extracts it by duck-typing
from a loaded reference model (no reference import here); ``save``/``load`` use a flat
``.npz`` so deployment needs neither pickle nor ASE.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

MAX_SPECIES = 8


@dataclass
class SgprModel:
    lmax: int
    nmax: int
    xi: float
    rc: float
    kind: str = "sesoap"              # "sesoap" | "universal" (informational)
    normalize: bool = True
    radii: dict = field(default_factory=dict)   # Z -> length unit; others -> default_radius
    default_radius: float = 1.0
    a_not: tuple = ()                 # species excluded as centres (EqAll exceptions)
    a_only: tuple = ()                # if non-empty: the only central species (SubSeSoapKernel `a`s)
    b_only: tuple = ()                # if non-empty: the only neighbour species that enter the descriptor (`b`)
    ind_Z: np.ndarray = None          # [M]
    ind_first: np.ndarray = None      # [M+1]
    ind_r: np.ndarray = None          # [nnz,3]
    ind_b: np.ndarray = None          # [nnz]
    mu: np.ndarray = None             # [M]
    mean_w: dict = field(default_factory=dict)   # Z -> weights[Z] + _weights[Z]
    choli: np.ndarray = None          # [M,M] or None
    lone_weight: float = 1.0          # kernels in the list (each adds the lone-lone term, similarity.py:41-43,94-103)
    vscale: dict = field(default_factory=dict)   # Z -> _vscale[Z]

    # ------------------------------------------------------------------ basics
    def __post_init__(self):
        self.ind_Z = np.ascontiguousarray(np.asarray(self.ind_Z if self.ind_Z is not None else [], dtype=np.int32))
        M = len(self.ind_Z)
        self.ind_first = np.ascontiguousarray(np.asarray(self.ind_first if self.ind_first is not None else np.zeros(M + 1), dtype=np.int64))
        self.ind_r = np.ascontiguousarray(np.asarray(self.ind_r if self.ind_r is not None else np.zeros((0, 3)), dtype=np.float64).reshape(-1, 3))
        self.ind_b = np.ascontiguousarray(np.asarray(self.ind_b if self.ind_b is not None else [], dtype=np.int32))
        self.mu = np.ascontiguousarray(np.asarray(self.mu if self.mu is not None else np.zeros(M), dtype=np.float64))
        if self.choli is not None:
            self.choli = np.ascontiguousarray(np.asarray(self.choli, dtype=np.float64))
            assert self.choli.shape == (M, M)
        assert len(self.ind_first) == M + 1 and len(self.mu) == M
        assert len(self.ind_b) == len(self.ind_r) == int(self.ind_first[-1])
        self.radii = {int(k): float(v) for k, v in self.radii.items()}
        self.mean_w = {int(k): float(v) for k, v in self.mean_w.items()}
        self.vscale = {int(k): float(v) for k, v in self.vscale.items()}
        self.a_not = tuple(int(z) for z in self.a_not)
        self.a_only = tuple(int(z) for z in self.a_only)
        self.b_only = tuple(int(z) for z in self.b_only)

    @property
    def M(self):
        return len(self.ind_Z)

    def unit_of(self, z):
        return self.radii.get(int(z), self.default_radius)

    def is_centre(self, z):
        return int(z) not in self.a_not and (not self.a_only or int(z) in self.a_only)

    def is_neighbour(self, z):
        return not self.b_only or int(z) in self.b_only

    def species(self, extra=()):
        """Sorted atomic numbers of the model (+ ``extra``): the dense species table."""
        s = set(int(z) for z in self.ind_Z) | set(int(z) for z in self.ind_b) | set(int(z) for z in np.asarray(extra).reshape(-1))
        return sorted(s)

    @classmethod
    def from_envs(cls, envs, **kw):
        """envs: list of (Z, r[nn,3], b[nn])."""
        first = np.cumsum([0] + [len(e[2]) for e in envs]).astype(np.int64)
        r = np.concatenate([np.asarray(e[1], dtype=float).reshape(-1, 3) for e in envs]) if envs else np.zeros((0, 3))
        b = np.concatenate([np.asarray(e[2], dtype=np.int32).reshape(-1) for e in envs]) if envs else np.zeros(0, np.int32)
        return cls(ind_Z=np.array([e[0] for e in envs], dtype=np.int32), ind_first=first, ind_r=r, ind_b=b, **kw)

    @classmethod
    def from_tape(cls, path, mu=None, **kw):
        """Inducing LCEs from an AutoForce ``.sgpr`` tape (theforce/io/sgprio.py); hyper-parameters and the
        weights are passed by the caller (the tape stores neither)."""
        from .sgprio import read_lces

        return cls.from_envs(read_lces(path), mu=mu, **kw)

    # ------------------------------------------------------------------ reference model
    @classmethod
    def from_posterior_potential(cls, model):
        """Extract from a reference ``PosteriorPotential`` (regression/gppotential.py:453).
        Supports one SeSoapKernel / UniversalSoapKernel in ``model.gp.kern.kernels``."""
        kerns = list(model.gp.kern.kernels)
        a_only, b_only = (), ()
        names = {type(k).__name__ for k in kerns}
        if len(kerns) > 1 or names & {"SubSeSoapKernel", "HeterogeneousSoapKernel"}:
            # default_kernel(species=...) (calculator/active.py:28-38): one SubSeSoapKernel per central species,
            # all with the same hyper-parameters and neighbour list b -> one dense model, centres = the a's.
            # Same for lists of its parent class HeterogeneousSoapKernel (similarity/heterosoap.py:10-29).
            if len(names) != 1 or not names <= {"SubSeSoapKernel", "HeterogeneousSoapKernel"}:
                raise NotImplementedError("kernel lists are supported for SubSeSoapKernel / HeterogeneousSoapKernel only")
            k0 = kerns[0]
            if any(_list_signature(k) != _list_signature(k0) for k in kerns):
                raise NotImplementedError("kernels with different hyper-parameters need one engine each")
            a_only = tuple(int(k.a) for k in kerns)
            b_only = tuple(int(z) for z in k0.b)
        lone_weight = float(len(kerns))
        k = kerns[0]
        cname = type(k).__name__
        desc = k.descriptor
        inner = getattr(desc, "soap", desc)   # NormalizedSoap wraps the descriptor proper
        lmax, nmax = int(inner.ylm.lmax), int(inner.nmax)
        species = set()
        envs = []
        for loc in model.X:
            b = np.asarray(loc._b.detach().cpu().numpy() if hasattr(loc._b, "detach") else loc._b).reshape(-1)
            r = np.asarray(loc._r.detach().cpu().numpy() if hasattr(loc._r, "detach") else loc._r).reshape(-1, 3)
            envs.append((int(loc.number), r, b))
            species.add(int(loc.number))
            species.update(int(z) for z in b)
        if cname in ("SeSoapKernel", "SubSeSoapKernel"):
            kind = "sesoap"
            radii = {z: float(desc.radii.get(z)) for z in species}
            default = float(desc.radii.get(10 ** 6)) if _safe_default(desc.radii) else 1.0
        elif cname == "UniversalSoapKernel":
            kind = "universal"
            radii, default = {}, float(desc.unit)
        elif cname == "HeterogeneousSoapKernel":
            # descriptor = [NormalizedSoap(] HeteroSoap [)]: one length unit for all species (descriptor/soap.py:13-26)
            normalized = type(desc).__name__ == "NormalizedSoap"
            soap = desc.soap if normalized else desc
            if type(soap).__name__ != "HeteroSoap":
                raise NotImplementedError(f"HeterogeneousSoapKernel over {type(soap).__name__} is not supported")
            kind = "universal"
            radii, default = {}, float(soap.unit)
            lmax, nmax = int(soap.ylm.lmax), int(soap.nmax)
            desc = type("D", (), {"normalize": normalized})()
        else:
            raise NotImplementedError(f"kernel class {cname} is not supported")
        a = getattr(k, "_a", None)
        a_not = tuple(getattr(a, "exceptions", ()) or ())
        if cname not in ("SubSeSoapKernel", "HeterogeneousSoapKernel") and a is not None and not hasattr(a, "exceptions"):
            a_only = (int(a),)   # a kernel restricted to one central species: loc.number == self.a (universal.py:101)
        if cname == "SubSeSoapKernel":
            exponent = _subse_exponent(k)
        elif cname == "HeterogeneousSoapKernel":
            exponent = _dotprod_exponent(k.kern)
        else:
            exponent = k.exponent
        mean = model.mean
        mean_w = {}
        for z, w in getattr(mean, "weights", {}).items():
            mean_w[int(z)] = float(w) + float(getattr(mean, "_weights", {}).get(z, 0.0))
        vscale = {int(z): float(v) for z, v in getattr(model, "_vscale", {}).items()}
        choli = getattr(model, "choli", None)
        return cls.from_envs(
            envs, lmax=lmax, nmax=nmax, xi=float(exponent), rc=float(k.cutoff), kind=kind,
            normalize=bool(desc.normalize), radii=radii, default_radius=default, a_not=a_not, a_only=a_only, b_only=b_only,
            mu=np.asarray(model.mu.detach().cpu().numpy(), dtype=float), mean_w=mean_w,
            choli=None if choli is None else np.asarray(choli.detach().cpu().numpy(), dtype=float), vscale=vscale,
            lone_weight=lone_weight,
        )

    @classmethod
    def list_from_posterior_potential(cls, model):
        """A reference model whose ``gp.kern.kernels`` cannot be merged into one dense model (similarity kernels with
        different lmax / nmax / exponent / cutoff, summed by EnergyForceKernel, regression/gppotential.py:81-84) as one
        SgprModel per kernel: energies, forces and virials of the handles add up.  The kernels of the reference share
        the neighbour list of the largest cutoff, so "neighbour-less" refers to that cutoff: the lone-atoms term (once
        per kernel, similarity.py:41-43,94-103) is carried by the model with the largest cutoff alone, and so is the mean.
        Falls back to ``[from_posterior_potential(model)]`` when one model suffices."""
        import types

        try:
            return [cls.from_posterior_potential(model)]
        except NotImplementedError:
            pass
        kerns = list(model.gp.kern.kernels)
        out = []
        for k in kerns:
            one = types.SimpleNamespace(gp=types.SimpleNamespace(kern=types.SimpleNamespace(kernels=[k])), X=model.X, mu=model.mu,
                                        mean=model.mean, choli=getattr(model, "choli", None), _vscale=getattr(model, "_vscale", {}))
            out.append(cls.from_posterior_potential(one))
        lead = max(range(len(out)), key=lambda i: out[i].rc)
        for i, m in enumerate(out):
            m.lone_weight = float(len(kerns)) if i == lead else -1.0
            if i != lead:
                m.mean_w = {}
        return out

    # ------------------------------------------------------------------ flat file format
    def save(self, path):
        meta = dict(format="autoforce_b200.sgpr_model", version=1, lmax=self.lmax, nmax=self.nmax, xi=self.xi, rc=self.rc,
                    kind=self.kind, normalize=self.normalize, radii={str(k): v for k, v in self.radii.items()},
                    default_radius=self.default_radius, a_not=list(self.a_not), a_only=list(self.a_only), b_only=list(self.b_only), lone_weight=self.lone_weight,
                    mean_w={str(k): v for k, v in self.mean_w.items()}, vscale={str(k): v for k, v in self.vscale.items()})
        arrays = dict(ind_Z=self.ind_Z, ind_first=self.ind_first, ind_r=self.ind_r, ind_b=self.ind_b, mu=self.mu)
        if self.choli is not None:
            arrays["choli"] = self.choli
        np.savez_compressed(path, meta=json.dumps(meta), **arrays)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        meta = json.loads(str(z["meta"]))
        if meta.get("format") != "autoforce_b200.sgpr_model":
            raise ValueError(f"{path}: not an autoforce_b200 model file")
        return cls(lmax=meta["lmax"], nmax=meta["nmax"], xi=meta["xi"], rc=meta["rc"], kind=meta["kind"],
                   normalize=meta["normalize"], radii=meta["radii"], default_radius=meta["default_radius"],
                   a_not=tuple(meta["a_not"]), a_only=tuple(meta.get("a_only", ())), b_only=tuple(meta.get("b_only", ())),
                   ind_Z=z["ind_Z"], ind_first=z["ind_first"], ind_r=z["ind_r"],
                   ind_b=z["ind_b"], mu=z["mu"], mean_w=meta["mean_w"], choli=z["choli"] if "choli" in z.files else None,
                   vscale=meta["vscale"], lone_weight=float(meta.get("lone_weight", 1.0)))


def _list_signature(k):
    """Everything of a SubSeSoapKernel / HeterogeneousSoapKernel but its central species."""
    desc = k.descriptor
    inner = getattr(desc, "soap", desc)
    if type(k).__name__ == "SubSeSoapKernel":
        return (int(inner.ylm.lmax), int(inner.nmax), _subse_exponent(k), float(k.cutoff), tuple(k.b), repr(desc.radii),
                bool(desc.normalize))
    return (int(inner.ylm.lmax), int(inner.nmax), _dotprod_exponent(k.kern), float(k.cutoff), tuple(k.b), float(inner.unit),
            type(desc).__name__)


def _dotprod_exponent(kern):
    """Base kernel of a HeterogeneousSoapKernel: only ``DotProd() ** eta`` (regression/kernel.py:187-199) maps onto
    (q_hat . z_hat)^xi."""
    if type(kern).__name__ == "Pow" and type(kern.kern).__name__ == "DotProd":
        return float(kern.eta)
    if type(kern).__name__ == "DotProd":
        return 1.0
    raise NotImplementedError(f"base kernel {getattr(kern, 'state', kern)} is not supported (DotProd() ** eta only)")


def _subse_exponent(k):
    """SubSeSoapKernel keeps its exponent inside ``kern = DotProd() ** exponent`` (similarity/sesoap.py:29);
    the eval-able argument string is '{lmax}, {nmax}, {exponent}, {cutoff}, {a}, {b}'."""
    return float(k._args.split(",")[2])


def _safe_default(radii):
    try:
        radii.get(10 ** 6)
        return True
    except Exception:
        return False
