"""Host-side mirror of the reference's similarity-kernel interface for the hot path.

Same constructor arguments, attribute names and eval-able ``state`` strings as
theforce.similarity.sesoap.SeSoapKernel (similarity/sesoap.py:10-24) and
theforce.similarity.universal.UniversalSoapKernel (similarity/universal.py:52-107);
``kern(atoms, inducing)`` returns the kernel matrix like
SimilarityKernel.forward(first, second, operation="func") (similarity/similarity.py:17-31)
-- evaluated on the GPU through libsgpr_b200 (no CPU path).
"""
from __future__ import annotations

import numpy as np

from .model import SgprModel


class DefaultRadii:
    """descriptor/sesoap.py:84-99: H -> 0.5, everything else -> 1.0."""

    def __init__(self, default=1.0, special=None):
        self.default = float(default)
        self.special = {1: 0.5} if special is None else {int(k): float(v) for k, v in special.items()}

    def get(self, number):
        return self.special.get(int(number), self.default)

    def __repr__(self):
        return f"DefaultRadii({self.default}, {self.special})"


class UniformRadii(DefaultRadii):
    def __init__(self, value=1.0):
        super().__init__(value, {})

    def __repr__(self):
        return f"UniformRadii({self.default})"


class EqAll:
    """util/util.py:113-124 / similarity/universal.py:44-49: equals everything but the exceptions."""

    def __init__(self, exceptions=()):
        self.exceptions = list(exceptions)

    def __eq__(self, val):
        return val not in self.exceptions


class _SoapKernelBase:
    kind = None

    def _base_model(self, envs=(), mu=None):
        return SgprModel.from_envs(
            list(envs), lmax=self.lmax, nmax=self.nmax, xi=float(self.exponent), rc=float(self.cutoff), kind=self.kind,
            normalize=self.normalize, radii=self._radii_dict(), default_radius=self._default_radius(),
            a_not=tuple(getattr(self._a, "exceptions", ())), a_only=tuple(getattr(self, "a_only", ())),
            b_only=tuple(getattr(self, "b", ())), mu=mu)

    @property
    def a(self):
        return self._a

    @property
    def state(self):
        return self.__class__.__name__ + "({})".format(self.state_args)

    def __repr__(self):
        return self.state

    def __call__(self, first, second, operation="func"):
        """first: structure with .positions/.cell/.pbc/.numbers; second: list of
        (Z, r[nn,3], b[nn]) local chemical environments.  ``operation`` as in similarity/similarity.py:17-31:
        "func" -> K [N, M];  "leftgrad" -> d(sum_i K[i,m])/d xyz [3N, M];  "virial" -> [6, M]
        (similarity/universal.py:109-183; true derivatives, see include/sgpr_b200.h sgpr_kernel_jacobian)."""
        if operation not in ("func", "leftgrad", "virial"):
            raise NotImplementedError(f"operation={operation!r}: only func / leftgrad / virial are implemented")
        from .engine import SgprEngine

        model = self._base_model(second)
        eng = SgprEngine(model, species=np.unique(first.numbers))
        try:
            args = (first.positions, first.numbers, np.asarray(first.cell), first.pbc)
            if operation == "func":
                return eng.kernel_matrix(*args).cpu().numpy()
            _, Kf, Kv = eng.kernel_jacobian(*args)
            return -Kf.cpu().numpy() if operation == "leftgrad" else Kv.cpu().numpy()
        finally:
            eng.close()


class SeSoapKernel(_SoapKernelBase):
    kind = "sesoap"

    def __init__(self, lmax, nmax, exponent, cutoff, a=None, radii=1.0, normalize=True):
        self.lmax, self.nmax, self.exponent, self.cutoff = int(lmax), int(nmax), exponent, float(cutoff)
        self.radii = UniformRadii(radii) if isinstance(radii, (int, float)) else radii
        self.normalize = bool(normalize)
        self.dim = (nmax + 1) ** 2 * (lmax + 1)
        self._a = EqAll() if a is None else a
        if not hasattr(self._a, "exceptions"):
            raise NotImplementedError("fixed central species (a=Z) belongs to SubSeSoapKernel, not on this path")
        self._args = f"{lmax}, {nmax}, {exponent}, {cutoff}, a={a}"
        self.name = "kern_0"
        self.params = []

    def _radii_dict(self):
        if isinstance(self.radii, dict):
            return dict(self.radii)
        return dict(getattr(self.radii, "special", {}))

    def _default_radius(self):
        return float(getattr(self.radii, "default", 1.0))

    @property
    def state_args(self):
        return f"{self._args}, radii={self.radii}, normalize={self.normalize}"


class SubSeSoapKernel(SeSoapKernel):
    """similarity/sesoap.py:27-43: fixed central species ``a`` and neighbour species list ``b`` (dense
    [S^2 * dim] descriptor); ``default_kernel(species=...)`` builds one per central species."""

    def __init__(self, lmax, nmax, exponent, cutoff, a, b, radii=1.0, normalize=True):
        super().__init__(lmax, nmax, exponent, cutoff, a=None, radii=radii, normalize=normalize)
        self.a_only = (int(a),)
        self.b = sorted(int(z) for z in (b if hasattr(b, "__iter__") else [b]))
        self._a = int(a)
        self.dim = len(self.b) ** 2 * (nmax + 1) ** 2 * (lmax + 1)
        self._args = f"{lmax}, {nmax}, {exponent}, {cutoff}, {a}, {b}"

    @property
    def a(self):
        return self._a


class HeterogeneousSoapKernel(_SoapKernelBase):
    """similarity/heterosoap.py:10-29 with the base kernel ``DotProd() ** exponent``: central species ``a``,
    neighbour species list ``b``, descriptor [NormalizedSoap(] HeteroSoap [)] with ONE length unit for all species
    (``atomic_unit`` or cutoff / 3, descriptor/soap.py:20-23)."""

    kind = "universal"

    def __init__(self, exponent, a, b, lmax, nmax, cutoff, atomic_unit=None, normalize=True):
        self.lmax, self.nmax, self.exponent, self.cutoff = int(lmax), int(nmax), exponent, float(cutoff)
        self.unit = float(atomic_unit) if atomic_unit else self.cutoff / 3
        self.normalize = bool(normalize)
        self.a_only = (int(a),)
        self._a = int(a)
        self.b = sorted(int(z) for z in (b if hasattr(b, "__iter__") else [b]))
        self.dim = len(self.b) ** 2 * (self.nmax + 1) ** 2 * (self.lmax + 1)
        self._args = "Pow(DotProd(), {}), {}, {}, {}, {}, PolyCut({}, n=2), atomic_unit={}, normalize={}".format(
            exponent, a, b, lmax, nmax, self.cutoff, atomic_unit, normalize)
        self.name = "kern_0"
        self.params = []

    def _radii_dict(self):
        return {}

    def _default_radius(self):
        return self.unit

    @property
    def state_args(self):
        return self._args


class UniversalSoapKernel(_SoapKernelBase):
    kind = "universal"

    def __init__(self, lmax, nmax, exponent, cutoff, atomic_unit=None, chemical=None, normalize=True, a=None, a_not=[]):
        if chemical is not None:
            raise NotImplementedError("only the Dirac-delta chemical kernel is supported")
        self.lmax, self.nmax, self.exponent, self.cutoff = int(lmax), int(nmax), exponent, float(cutoff)
        self.unit = float(atomic_unit) if atomic_unit else self.cutoff / 6  # descriptor/soap.py:724-728
        self.normalize = bool(normalize)
        self.dim = (nmax + 1) ** 2 * (lmax + 1)
        self._a = EqAll(a_not) if a is None else a
        self._args = "{}, {}, {}, PolyCut({}, n=2), atomic_unit={}, chemical=DiracDeltaChemical(), normalize={}, a={}, a_not={}".format(
            lmax, nmax, exponent, self.cutoff, atomic_unit, normalize, a, a_not)
        self.name = "kern_0"
        self.params = []

    def _radii_dict(self):
        return {}

    def _default_radius(self):
        return self.unit

    @property
    def state_args(self):
        return self._args
