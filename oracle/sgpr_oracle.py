"""CPU oracle (numpy, float64) for AutoForce's SGPR prediction hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  This file restates, in
plain numpy, the algorithm of the reference (`/root/reference/theforce`, v2021.09)
for the path  neighbour list -> SOAP-type descriptor -> (p.z)^xi kernel ->
E / F / stress / covloss.  Every function cites the reference lines it follows.
It is pinned against (a) the reference's only known-answer vector
(descriptor/soap.py:488-532) and (b) outputs of the reference itself, generated
in the build container by ``tests/golden/make_golden.py`` and committed under
``tests/golden/``  (see tests/test_oracle_golden.py).

Neighbour-list semantics are those of ASE's ``NewPrimitiveNeighborList`` (third
party, no version pinned by the reference, not installed in this image): the
restatement below follows ASE's documented behaviour and the reference's call
site descriptor/atoms.py:348-368 -> *NL parity is unpinned* (DESIGN.md).

The derivative path here is deliberately written the way the reference's own
analytic derivatives are (spherical-coordinate partials of Y_lm, ylm.py:191-222),
which is a different formulation from the CUDA kernels (Cartesian polynomial
recursion), so that agreement between the two is a real cross-check.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from math import factorial, pi

import numpy as np

TINY_ANGLE = 1e-2  # descriptor/ylm.py:10
EPS = float(np.finfo(np.float64).eps)  # torch.finfo().eps, descriptor/sesoap.py:233


# --------------------------------------------------------------------------
# model container
# --------------------------------------------------------------------------
@dataclass
class OracleModel:
    """Flat description of a frozen SGPR model.

    radii : dict Z -> length unit of neighbour species Z.  SeSoap: ``radii(Z)``
            (descriptor/sesoap.py:84-99,162); UniversalSoap: the same ``unit``
            for every Z (descriptor/soap.py:724-728).
    """

    lmax: int
    nmax: int
    xi: float
    rc: float
    radii: dict
    ind_Z: np.ndarray  # [M] central species of inducing LCEs
    ind_r: list  # M arrays [nn_m,3]  (Local._r)
    ind_b: list  # M arrays [nn_m]    (Local._b)
    mu: np.ndarray  # [M]
    mean_w: dict = field(default_factory=dict)  # Z -> weights[Z] + _weights[Z]
    choli: np.ndarray | None = None  # [M,M]
    vscale: dict = field(default_factory=dict)  # Z -> vscale
    normalize: bool = True
    a_not: tuple = ()  # species excluded as centres (similarity/universal.py:44-49,85)
    a_only: tuple = ()  # if non-empty: the central species of the SubSeSoapKernels (similarity/sesoap.py:27-43)
    b_only: tuple = ()  # if non-empty: neighbour species that enter the descriptor (SubSeSoap `numbers`)
    default_radius: float = 1.0
    lone_weight: float = 1.0  # kernels in the list: each adds the lone-lone term (similarity/similarity.py:41-43,94-103)

    @property
    def M(self):
        return len(self.ind_Z)

    def unit_of(self, z):
        return float(self.radii.get(int(z), self.default_radius))

    def excluded_centres(self, Z):
        Z = np.asarray(Z, dtype=np.int64)
        ex = np.isin(Z, np.asarray(self.a_not, dtype=np.int64))
        if len(self.a_only):
            ex |= ~np.isin(Z, np.asarray(self.a_only, dtype=np.int64))
        return ex

    def species_table(self, extra=()):
        s = set(int(z) for z in self.ind_Z)
        for b in self.ind_b:
            s.update(int(z) for z in np.asarray(b).reshape(-1))
        s.update(int(z) for z in extra)
        return np.array(sorted(s), dtype=np.int64)


# --------------------------------------------------------------------------
# a1: neighbour list  (descriptor/atoms.py:348-363,402 -> ASE NeighborList)
# --------------------------------------------------------------------------
def complete_cell(cell):
    """ASE ``complete_cell``: replace zero cell vectors by unit vectors orthogonal
    to the others (used by ``Atoms.get_cell(complete=True)`` which ASE's
    NeighborList.update passes on)."""
    cell = np.array(cell, dtype=float)
    missing = np.nonzero(~cell.any(axis=1))[0]
    if len(missing) == 3:
        cell.flat[::4] = 1.0
    if len(missing) == 2:
        i = 3 - missing.sum()  # the one present
        assert abs(cell[i]).sum() > 0
        # two orthonormal vectors orthogonal to cell[i]
        v = cell[i] / np.linalg.norm(cell[i])
        t = np.eye(3)[np.argmin(abs(v))]
        e1 = np.cross(v, t)
        e1 /= np.linalg.norm(e1)
        e2 = np.cross(v, e1)
        cell[missing[0]] = e1
        cell[missing[1]] = e2
    if len(missing) == 1:
        i = missing[0]
        cell[i] = np.cross(cell[i - 2], cell[i - 1])
        cell[i] /= np.linalg.norm(cell[i])
    return cell


def face_distances(cell):
    inv = np.linalg.inv(cell)  # columns are reciprocal vectors
    return 1.0 / np.linalg.norm(inv, axis=0)


def neighbor_list(pos, cell, pbc, rc, centers=None):
    """All (i, j, S) with |pos[j] - pos[i] + S@cell| < rc  (strict), both ways,
    (i==j, S==0) excluded, images of the same atom (and of i itself) kept.

    ASE ``NeighborList(N*[rc/2], skin=0, self_interaction=False, bothways=True,
    primitive=NewPrimitiveNeighborList)`` as constructed at descriptor/atoms.py:
    348-355; offsets are relative to the *given* positions, distance is
    ``sqrt(sum((pos[j]-pos[i]+S@cell)**2))`` and the test is ``< c_i + c_j``.

    Returns CSR (first[N+1] int64, j[nnz] int64, S[nnz,3] int64) with each row
    sorted lexicographically by (j, S0, S1, S2) (ASE's order is implementation
    defined; comparisons are done on sorted rows).  ``centers`` restricts the rows
    that are filled to a subset of atoms (the other rows stay empty).
    """
    from scipy.spatial import cKDTree

    pos = np.asarray(pos, dtype=float).reshape(-1, 3)
    n = len(pos)
    pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,))
    cell = complete_cell(cell)
    if n == 0:
        return np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros((0, 3), np.int64)
    h = face_distances(cell)
    frac = pos @ np.linalg.inv(cell)
    lo, hi = frac.min(axis=0), frac.max(axis=0)
    rng = []
    for c in range(3):
        if pbc[c]:
            k = int(np.ceil(rc / h[c] + (hi[c] - lo[c]))) + 1
            rng.append(range(-k, k + 1))
        else:
            rng.append(range(0, 1))
    cidx = np.arange(n) if centers is None else np.asarray(centers, dtype=np.int64)
    tree = cKDTree(pos[cidx])
    I, J, S = [], [], []
    margin = rc / h
    for s in itertools.product(*rng):
        s = np.array(s, dtype=np.int64)
        # quick reject of image cells that cannot reach the atoms
        f2 = frac + s
        keep = np.all((f2 >= lo - margin - 1e-9) & (f2 <= hi + margin + 1e-9), axis=1)
        if not keep.any():
            continue
        idx = np.nonzero(keep)[0]
        img = pos[idx] + s.astype(float) @ cell
        # candidates (slightly generous), then the exact reference test
        lists = cKDTree(img).query_ball_tree(tree, rc * (1 + 1e-9) + 1e-9)
        jj = np.concatenate([np.full(len(l), idx[k], np.int64) for k, l in enumerate(lists)] or [np.zeros(0, np.int64)])
        ii = np.concatenate([np.asarray(l, np.int64) for l in lists] or [np.zeros(0, np.int64)])
        if len(ii) == 0:
            continue
        ii = cidx[ii]
        dv = pos[jj] - pos[ii] + (s.astype(float)[:, None] * cell).sum(axis=0)
        d = np.sqrt((dv * dv).sum(axis=1))
        m = d < rc
        if not s.any():
            m &= ii != jj
        I.append(ii[m])
        J.append(jj[m])
        S.append(np.repeat(s[None], int(m.sum()), axis=0))
    if not I:
        return np.zeros(n + 1, np.int64), np.zeros(0, np.int64), np.zeros((0, 3), np.int64)
    I = np.concatenate(I)
    J = np.concatenate(J)
    S = np.concatenate(S)
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], J, I))
    I, J, S = I[order], J[order], S[order]
    first = np.searchsorted(I, np.arange(n + 1)).astype(np.int64)
    return first, J, S


def neighbor_list_bruteforce(pos, cell, pbc, rc):
    """O(N^2 * images) version of :func:`neighbor_list` (no KD-tree) used to
    cross-check it in the tests."""
    pos = np.asarray(pos, dtype=float).reshape(-1, 3)
    n = len(pos)
    pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,))
    cell = complete_cell(cell)
    h = face_distances(cell)
    frac = pos @ np.linalg.inv(cell)
    span = frac.max(axis=0) - frac.min(axis=0) if n else np.zeros(3)
    rng = [
        range(-int(np.ceil(rc / h[c] + span[c])) - 1, int(np.ceil(rc / h[c] + span[c])) + 2) if pbc[c] else range(0, 1)
        for c in range(3)
    ]
    I, J, S = [], [], []
    for s in itertools.product(*rng):
        s = np.array(s, dtype=np.int64)
        dv = pos[None, :, :] - pos[:, None, :] + (s.astype(float)[:, None] * cell).sum(axis=0)
        d = np.sqrt((dv * dv).sum(-1))
        m = d < rc
        if not s.any():
            m &= ~np.eye(n, dtype=bool)
        i, j = np.nonzero(m)
        I.append(i)
        J.append(j)
        S.append(np.repeat(s[None], len(i), 0))
    I = np.concatenate(I)
    J = np.concatenate(J)
    S = np.concatenate(S)
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], J, I))
    I, J, S = I[order], J[order], S[order]
    first = np.searchsorted(I, np.arange(n + 1)).astype(np.int64)
    return first, J, S


# --------------------------------------------------------------------------
# a2: displacements  (descriptor/atoms.py:365-368)
# --------------------------------------------------------------------------
def displacements(pos, cell, i, j, S):
    """``r = xyz[n] - xyz[a] + (off[..., None] * lll).sum(dim=1)`` in this op order."""
    cells = (S.astype(float)[..., None] * np.asarray(cell, dtype=float)).sum(axis=1)
    return pos[j] - pos[i] + cells


def pad_environments(envs_r, envs_b):
    """list of [nn,3] / [nn] -> padded [B,nnmax,3], [B,nnmax], mask[B,nnmax]."""
    B = len(envs_r)
    nnmax = max([len(b) for b in envs_b] + [1])
    R = np.zeros((B, nnmax, 3))
    R[..., :] = 1.0  # harmless dummy for padded slots
    Z = np.full((B, nnmax), -1, dtype=np.int64)
    mask = np.zeros((B, nnmax), dtype=bool)
    for k, (r, b) in enumerate(zip(envs_r, envs_b)):
        nn = len(b)
        if nn:
            R[k, :nn] = np.asarray(r, dtype=float).reshape(nn, 3)
            Z[k, :nn] = np.asarray(b).reshape(nn)
            mask[k, :nn] = True
    return R, Z, mask


# --------------------------------------------------------------------------
# a4: solid harmonics r^l Y_lm   (descriptor/ylm.py:44-225)
# --------------------------------------------------------------------------
class YlmTables:
    def __init__(self, lmax):
        self.lmax = lmax
        L = lmax + 1
        self.Yoo = np.sqrt(1.0 / (4 * pi))  # ylm.py:57
        # ylm.py:58-76  (a_lm, b_lm for m <= l-2)
        self.al = [None, None] + [
            np.array([np.sqrt((4.0 * l * l - 1.0) / (l * l - m * m)) for m in range(l - 1)]) for l in range(2, L)
        ]
        self.bl = [None, None] + [
            np.array([-np.sqrt(((l - 1.0) ** 2 - m * m) / (4 * (l - 1.0) ** 2 - 1)) for m in range(l - 1)])
            for l in range(2, L)
        ]
        self.cl = [np.sqrt(2.0 * l + 1.0) for l in range(L)]  # ylm.py:77
        self.dl = [None] + [-np.sqrt(1.0 + 1.0 / (2.0 * l)) for l in range(1, L)]  # ylm.py:78-80
        i = np.arange(L)
        self.l = np.maximum(i[:, None], i[None, :]).astype(float)  # ylm.py:87-92
        self.m = np.abs(i[:, None] - i[None, :]).astype(float)  # ylm.py:93-96
        one = np.ones((L, L))
        self.sign = -np.tril(one, -1) + np.triu(one)  # ylm.py:99-100
        with np.errstate(invalid="ignore", divide="ignore"):
            coef = (self.l - self.m) * (self.l + self.m) * (2 * self.l + 1) / (2 * self.l - 1)
        self.coef = np.sqrt(coef[1:, 1:])  # ylm.py:103-106
        # contraction masks (descriptor/sesoap.py:116-118)
        self.Yr = 2 * np.tril(one) - np.eye(L)
        self.Yi = 2 * np.triu(one, 1)


def shear_flag(xyz, mask=None):
    """split_and_rotate_tiny_if_too_close_to_zaxis, descriptor/ylm.py:10-23:
    environment-wide flag (``.any()`` over the neighbours of ONE environment)."""
    tol = TINY_ANGLE * np.abs(xyz[..., 2])
    hit = (np.abs(xyz[..., 0]) < tol) & (np.abs(xyz[..., 1]) < tol)
    if mask is not None:
        hit = hit & mask
    return hit.any(axis=-1)


def ylm(xyz, tables, flag, grad=False):
    """Y[..., i, j] with Re(l,m) at [l, l-m] and Im(l,m) at [l-m, l]
    (descriptor/ylm.py:113-190) for points xyz[..., 3]; ``flag[...]`` (bool,
    broadcastable to xyz[..., 0]) selects the sheared frame (ylm.py:16-21).
    With grad=True also returns dY[..., i, j, 3] w.r.t. the *unsheared* xyz
    (ylm.py:191-222)."""
    lmax = tables.lmax
    L = lmax + 1
    a = np.where(flag, TINY_ANGLE, 0.0)
    x = xyz[..., 0]
    y = xyz[..., 1] - a * xyz[..., 2]
    z = a * xyz[..., 1] + xyz[..., 2]
    # cart_coord_to_trig, ylm.py:26-34
    rxy_sq = x * x + y * y
    rxy = np.sqrt(rxy_sq)
    r = np.sqrt(rxy_sq + z * z)
    sin_theta = rxy / r
    cos_theta = z / r
    sin_phi = y / rxy
    cos_phi = x / rxy
    r2 = r * r
    r_sin_theta = r * sin_theta
    r_cos_theta = r * cos_theta
    alp = [[np.full_like(sin_theta, tables.Yoo)]]
    for l in range(1, L):
        row = [
            tables.al[l][m] * (r_cos_theta * alp[l - 1][m] + r2 * tables.bl[l][m] * alp[l - 2][m]) for m in range(l - 1)
        ]
        row.append(tables.cl[l] * r_cos_theta * alp[l - 1][l - 1])
        row.append(tables.dl[l] * r_sin_theta * alp[l - 1][l - 1])
        alp.append(row)
    sin = [np.zeros_like(sin_phi), sin_phi]
    cos = [np.ones_like(cos_phi), cos_phi]
    for m in range(2, L):
        s = sin_phi * cos[-1] + cos_phi * sin[-1]
        c = cos_phi * cos[-1] - sin_phi * sin[-1]
        sin.append(s)
        cos.append(c)
    Y = np.zeros(x.shape + (L, L))
    for l in range(L):
        for m in range(l + 1):
            Y[..., l, l - m] = alp[l][m] * cos[m]
            if m > 0:
                Y[..., l - m, l] = alp[l][m] * sin[m]
    if not grad:
        return Y
    e = (..., None, None)
    Y_r = tables.l * Y / r[e]
    Y_theta = cos_theta[e] * tables.l * Y / sin_theta[e]
    Y_theta[..., 1:, 1:] -= r[e] * Y[..., :-1, :-1] * tables.coef / sin_theta[e]
    Y_phi = np.swapaxes(Y, -1, -2) * tables.sign * tables.m
    F_r, F_t, F_p = Y_r, Y_theta / r[e], Y_phi / (r * sin_theta)[e]
    st, ct, sp, cp = sin_theta[e], cos_theta[e], sin_phi[e], cos_phi[e]
    # sph_vec_to_cart, ylm.py:37-41
    gx = st * cp * F_r + ct * cp * F_t - sp * F_p
    gy = st * sp * F_r + ct * sp * F_t + cp * F_p
    gz = ct * F_r - st * F_t
    ae = a[e] if np.ndim(a) else a
    dY = np.stack([gx, gy + ae * gz, -ae * gy + gz], axis=-1)  # ylm.py:212-220
    return Y, dY


# --------------------------------------------------------------------------
# a3: descriptor  (descriptor/sesoap.py:161-260 == descriptor/soap.py:765-851)
# --------------------------------------------------------------------------
def nnl_table(lmax, nmax):
    """descriptor/sesoap.py:119-128."""
    a = np.array(
        [[1.0 / ((2 * l + 1) * 2 ** (2 * n + l) * factorial(n) * factorial(n + l)) for l in range(lmax + 1)] for n in range(nmax + 1)]
    )
    return np.sqrt(a[None] * a[:, None])


def _radial(model, units, d, grad):
    """PolyCut(rc)(units*d) * exp(-d^2/2) and its d-derivative
    (descriptor/cutoff.py:20-44, descriptor/sesoap.py:176-183)."""
    dr_true = units * d
    step = np.where(dr_true < model.rc, 1.0, 0.0)
    cut = step * (1.0 - dr_true / model.rc) ** 2
    ex = np.exp(-0.5 * d**2)
    if not grad:
        return cut * ex, None
    dcut = step * (-2.0 * (1.0 - dr_true / model.rc) / model.rc)
    dr = (units * dcut) * ex + cut * (-d * ex)
    return cut * ex, dr


def descriptor_batch(model, species, R, Zb, mask, want_aux=False):
    """Dense restatement of SeSoap.forward / UniversalSoap.forward for a padded
    batch of environments.

    species : sorted table of atomic numbers defining the dense block layout
              (the reference builds blocks only for species present among the
              neighbours, ``torch.unique`` sesoap.py:163; absent blocks are
              exactly zero here, which is equivalent for norms and dot products).
    returns p_hat[B,S,S,n,n,L]  (block [s1,s2] <-> sparse index (Z_s2, Z_s1),
            sesoap.py:165-171) and, if want_aux, the intermediates.
    """
    species = np.asarray(species)
    S = len(species)
    B, nn = mask.shape
    tab = YlmTables(model.lmax)
    nnl = nnl_table(model.lmax, model.nmax)
    units = np.ones((B, nn))
    sidx = np.zeros((B, nn), dtype=np.int64)
    for k, z in enumerate(species):
        sel = Zb == z
        units[sel] = model.unit_of(z)
        sidx[sel] = k
    if mask.any():
        assert np.isin(Zb[mask], species).all(), "species table does not cover all neighbour species"
    xyz = R / units[..., None]
    d = np.sqrt((xyz**2).sum(axis=-1))
    n2 = 2.0 * np.arange(model.nmax + 1)
    rad, drad = _radial(model, units, d, grad=want_aux)
    rad = rad * mask
    if len(model.b_only):  # SubSeSoap sums only over the species of its list (descriptor/sesoap.py:322-326)
        rad = rad * np.isin(Zb, np.asarray(model.b_only, dtype=np.int64))
    f = rad[:, None, :] * d[:, None, :] ** n2[None, :, None]  # [B,n,nn]
    flag = shear_flag(xyz, mask)  # [B]
    if want_aux:
        Y, dY = ylm(xyz, tab, flag[:, None], grad=True)
    else:
        Y = ylm(xyz, tab, flag[:, None])
    onehot = (sidx[..., None] == np.arange(S)) & mask[..., None]  # [B,nn,S]
    # c[s,n,i,j] = sum_{nbr in s} f[n,nbr] Y[i,j,nbr]     sesoap.py:186-194
    c = np.einsum("bjs,bnj,bjik->bsnik", onehot.astype(float), f, Y, optimize=True)
    # nnp[s1,s2,n1,n2] = c[s2,n2] * c[s1,n1];  p = (nnp*Yr).sum(-1) + (nnp*Yi).sum(-2)   sesoap.py:195-203
    L = model.lmax + 1
    Wl = np.zeros((L, L, L))
    for l in range(L):
        Wl[l, l, :] += tab.Yr[l, :]
        Wl[l, :, l] += tab.Yi[:, l]
    p = np.einsum("bsnik,btmik,lik->bstnml", c, c, Wl, optimize=True)
    p = p * nnl
    if model.normalize:
        norm = np.sqrt((p**2).sum(axis=(1, 2, 3, 4, 5))) + EPS  # sesoap.py:249-251
    else:
        norm = np.ones(B)
    p_hat = p / norm[:, None, None, None, None, None]
    if not want_aux:
        return p_hat
    aux = dict(
        xyz=xyz, d=d, units=units, sidx=sidx, f=f, rad=rad, drad=drad * mask, Y=Y, dY=dY, c=c, norm=norm, Wl=Wl, nnl=nnl,
        flag=flag, onehot=onehot, n2=n2,
    )
    return p_hat, aux


def descriptor_backward(model, p_hat, aux, g):
    """Given g = dE/dp_hat [B,S,S,n,n,L] return dE/dR [B,nn,3] -- the analytic
    chain rule the reference delegates to autograd (calculator/active.py:587-599)
    and also spells out in SeSoap.forward(grad=True), sesoap.py:204-246."""
    norm = aux["norm"][:, None, None, None, None, None]
    if model.normalize:
        # p_hat = p / (|p| + eps):  dE/dp = [g - p_hat (p_hat.g) (|p| + eps) / |p|] / (|p| + eps)  -- what autograd gives
        # (calculator/active.py:587-599); sesoap.py:229-235 drops the factor (|p| + eps) / |p| = 1 + O(eps / |p|)
        pg = (p_hat * g).sum(axis=(1, 2, 3, 4, 5), keepdims=True)
        eps = np.finfo(float).eps
        with np.errstate(divide="ignore", invalid="ignore"):
            corr = np.where(norm > 2 * eps, norm / (norm - eps), 1.0)
        dp = (g - p_hat * pg * corr) / norm
    else:
        dp = g
    q = dp * aux["nnl"]
    qs = q + np.transpose(q, (0, 2, 1, 4, 3, 5))
    c = aux["c"]
    # dE/dc[s1,n1,i,k] = sum_{s2,n2,l} (q[s1,s2,n1,n2,l] + q[s2,s1,n2,n1,l]) Wl[l,i,k] c[s2,n2,i,k]
    dc = np.einsum("bstnml,lik,btmik->bsnik", qs, aux["Wl"], c, optimize=True)
    onehot = aux["onehot"].astype(float)
    Dsel = np.einsum("bjs,bsnik->bjnik", onehot, dc, optimize=True)
    f, Y, dY, d, xyz = aux["f"], aux["Y"], aux["dY"], aux["d"], aux["xyz"]
    n2 = aux["n2"]
    # df/dd   sesoap.py:205-208
    with np.errstate(divide="ignore", invalid="ignore"):
        dpow = np.where(n2[None, :, None] > 0, n2[None, :, None] * d[:, None, :] ** (n2[None, :, None] - 1), 0.0)
    df = aux["drad"][:, None, :] * d[:, None, :] ** n2[None, :, None] + aux["rad"][:, None, :] * dpow
    T1 = np.einsum("bjnik,bjik->bjn", Dsel, Y, optimize=True)
    gr = np.einsum("bnj,bjn->bj", df, T1, optimize=True)
    T2 = np.einsum("bjnik,bnj->bjik", Dsel, f, optimize=True)
    gy = np.einsum("bjik,bjikx->bjx", T2, dY, optimize=True)
    gxyz = gr[..., None] * xyz / d[..., None] + gy
    return gxyz / aux["units"][..., None]


# --------------------------------------------------------------------------
# a5/a6: kernel matrix  (similarity/universal.py:100-122, similarity.py:17-43,94-103)
# --------------------------------------------------------------------------
def inducing_descriptors(model, species, chunk=256):
    M = model.M
    S, n, L = len(species), model.nmax + 1, model.lmax + 1
    Zh = np.zeros((M, S, S, n, n, L))
    for k0 in range(0, M, chunk):
        sl = slice(k0, min(M, k0 + chunk))
        R, Zb, mask = pad_environments(model.ind_r[sl], model.ind_b[sl])
        Zh[sl] = descriptor_batch(model, species, R, Zb, mask)
    lone = np.array([len(b) == 0 for b in model.ind_b], dtype=bool)
    return Zh, lone


def _powxi(k, xi):
    return k ** int(xi) if float(xi).is_integer() else k**xi


def kernel_from_descriptors(model, P, Zc, lone_c, Zh, lone_m):
    """K[i,m] = delta(Z_i,Z_m) (p_i . z_m)^xi + lone_weight [both lone, same Z]."""
    B, M = len(P), len(Zh)
    dot = P.reshape(B, -1) @ Zh.reshape(M, -1).T
    same = Zc[:, None] == np.asarray(model.ind_Z)[None, :]
    ok_c = ~model.excluded_centres(Zc) & ~lone_c
    ok_m = ~model.excluded_centres(model.ind_Z) & ~lone_m
    K = np.where(same & ok_c[:, None] & ok_m[None, :], _powxi(dot, model.xi), 0.0)
    K = K + model.lone_weight * (same & lone_c[:, None] & lone_m[None, :])
    return K, dot, same & ok_c[:, None] & ok_m[None, :]


# --------------------------------------------------------------------------
# a7-a9: energy / forces / stress / covloss  (calculator/active.py:548-611,781-804)
# --------------------------------------------------------------------------
def predict(model, pos, cell, pbc, numbers, want_K=False, want_beta=False, chunk=256, atoms=None, nl=None, Zh=None,
            row_weights=None):
    """Full restatement of ``ActiveCalculator.calculate`` in prediction mode.

    returns dict(energy, forces[N,3], stress[6] (xx,yy,zz,yz,xz,xy), virial[3,3],
                 first/j/S neighbour list, optional K[N,M], beta[N]).
    ``atoms``: optional index subset whose local energies are evaluated (the
    reference's per-rank ``atoms.indices``, descriptor/atoms.py:321-341); forces
    and virial are then the partial sums over those environments.
    ``row_weights`` [N, M]: a cotangent dL/dK replacing mu row by row -- ``forces`` / ``virial`` are then the
    vector-Jacobian product that torch.autograd.grad computes through ``cov`` in the reference
    (calculator/active.py:587-599, regression/gppotential.py:905-911).
    """
    if isinstance(model, (list, tuple)):
        return predict_kernel_list(model, pos, cell, pbc, numbers, want_K=want_K, want_beta=want_beta, chunk=chunk, atoms=atoms)
    pos = np.asarray(pos, dtype=float).reshape(-1, 3)
    numbers = np.asarray(numbers, dtype=np.int64)
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    N = len(pos)
    species = model.species_table(extra=numbers)
    first, J, S = neighbor_list(pos, cell, pbc, model.rc, centers=atoms) if nl is None else nl
    Zh, lone_m = inducing_descriptors(model, species, chunk) if Zh is None else Zh
    idx_all = np.arange(N) if atoms is None else np.asarray(atoms, dtype=np.int64)
    F = np.zeros((N, 3))
    W = np.zeros((3, 3))
    e_local = np.zeros(len(idx_all))
    Kout = np.zeros((len(idx_all), model.M)) if (want_K or want_beta) else None
    alpha_all = np.ones(len(idx_all))
    mu = np.asarray(model.mu, dtype=float)
    for k0 in range(0, len(idx_all), chunk):
        idx = idx_all[k0 : k0 + chunk]
        envs_r, envs_b, envs_j = [], [], []
        for i in idx:
            sl = slice(first[i], first[i + 1])
            envs_r.append(displacements(pos, cell, i, J[sl], S[sl]))
            envs_b.append(numbers[J[sl]])
            envs_j.append(J[sl])
        R, Zb, mask = pad_environments(envs_r, envs_b)
        P, aux = descriptor_batch(model, species, R, Zb, mask, want_aux=True)
        lone_c = ~mask.any(axis=1)
        K, dot, valid = kernel_from_descriptors(model, P, numbers[idx], lone_c, Zh, lone_m)
        e_local[k0 : k0 + len(idx)] = K @ mu
        # self kernel k(x,x) (active.py:785-788): 1 for a normalised descriptor, 0 for an
        # excluded centre, 1 for a neighbour-less atom (similarity.py:94-103)
        # self kernel (p_hat . p_hat)^xi: 1 up to O(eps / |p|) for a normalised descriptor, 0 for a zero descriptor (an
        # environment whose neighbours all lie beyond this kernel's cutoff -- possible in kernel lists, which share the
        # neighbour list of the largest cutoff)
        selfk = _powxi((P.reshape(len(idx), -1) ** 2).sum(axis=1), model.xi)
        alpha_all[k0 : k0 + len(idx)] = np.where(lone_c, model.lone_weight,
                                                 np.where(model.excluded_centres(numbers[idx]), 0.0, selfk))
        if Kout is not None:
            Kout[k0 : k0 + len(idx)] = K
        xi = model.xi
        wrow = mu[None, :] if row_weights is None else np.asarray(row_weights, dtype=float)[idx]
        Gm = np.where(valid, xi * _powxi(dot, xi - 1) * wrow, 0.0)
        g = (Gm @ Zh.reshape(model.M, -1)).reshape(P.shape)
        dR = descriptor_backward(model, P, aux, g)
        for b, i in enumerate(idx):
            nn = len(envs_j[b])
            if nn == 0:
                continue
            gb = dR[b, :nn]
            F[i] += gb.sum(axis=0)
            np.subtract.at(F, envs_j[b], gb)
            W += envs_r[b].T @ gb
    energy = e_local.sum()
    if atoms is None:
        for z in np.unique(numbers):
            if int(z) in model.mean_w:
                energy += (numbers == z).sum() * model.mean_w[int(z)]
    try:
        vol = abs(np.linalg.det(cell))
        if vol == 0.0:
            raise ValueError
    except ValueError:
        vol = -2.0  # calculator/active.py:606-609
    stress = (W / vol).reshape(-1)[[0, 4, 8, 5, 2, 1]]
    out = dict(energy=energy, forces=F, stress=stress, virial=W, first=first, j=J, S=S, e_local=e_local, alpha=alpha_all)
    if want_K:
        out["K"] = Kout
    if want_beta:
        out["beta"] = covloss(model, Kout, numbers[idx_all], alpha=alpha_all)
    return out


def predict_kernel_list(models, pos, cell, pbc, numbers, want_K=False, want_beta=False, chunk=256, atoms=None):
    """A model whose EnergyForceKernel sums several similarity kernels with DIFFERENT hyper-parameters
    (regression/gppotential.py:81-84): the kernels share mu, choli, the inducing LCEs and ONE neighbour list, built with
    the largest cutoff (model.cutoff, descriptor/atoms.py:348-355) -- so "neighbour-less" (similarity.py:94-103) means no
    neighbour within the largest cutoff, and every kernel adds the lone-atoms term.  ``models``: one OracleModel per
    kernel, each with lone_weight 1; the mean is taken from the first."""
    pos = np.asarray(pos, dtype=float).reshape(-1, 3)
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    numbers = np.asarray(numbers, dtype=np.int64)
    nl = neighbor_list(pos, cell, pbc, max(m.rc for m in models), centers=atoms)
    tot = None
    for k, m in enumerate(models):
        if k > 0:
            m = dataclass_replace(m, mean_w={})
        out = predict(m, pos, cell, pbc, numbers, want_K=True, chunk=chunk, atoms=atoms, nl=nl)
        if tot is None:
            tot = dict(out)
        else:
            for key in ("energy", "forces", "virial", "stress", "e_local", "K", "alpha"):
                tot[key] = tot[key] + out[key]
    if want_beta:
        idx = np.arange(len(numbers)) if atoms is None else np.asarray(atoms, dtype=np.int64)
        tot["beta"] = covloss(models[0], tot["K"], numbers[idx], alpha=tot["alpha"])
    return tot


def dataclass_replace(obj, **kw):
    import dataclasses

    return dataclasses.replace(obj, **kw)


def kernel_jacobian(model, pos, cell, pbc, numbers, chunk=256, scatter_quirk=False):
    """Training-time kernels of ONE structure against the inducing set, as ``EnergyForceKernel`` builds them
    (regression/gppotential.py:63-77 over similarity/universal.py:109-183):

      Ke [M]     = energy_energy([atoms], X)  = sum_i K[i,m]
      Kf [3N, M] = forces_energy([atoms], X)  = -leftgrad = -d Ke[m] / d xyz
      Kv [6, M]  = virial_energy([atoms], X)  = sum_pairs r (x) dk/dr, picked [0,4,8,5,2,1] (xx,yy,zz,yz,xz,xy),
                   NOT divided by the volume (universal.py:155-183)

    i.e. the prediction backward pass with mu replaced by each unit vector e_m in turn.

    ``scatter_quirk=True`` reproduces the reference's hand-written ``leftgrad`` instead of the true derivative:
    ``g[j] += f`` (universal.py:148) is an index assignment, so when the same atom j occurs more than once in an
    environment (periodic images in a cell narrower than 2 rc) only the LAST of its contributions survives.
    The reference's autograd forces (calculator/active.py:587-611) do not have this defect."""
    pos = np.asarray(pos, dtype=float).reshape(-1, 3)
    numbers = np.asarray(numbers, dtype=np.int64)
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    N, M = len(pos), model.M
    species = model.species_table(extra=numbers)
    first, J, S = neighbor_list(pos, cell, pbc, model.rc)
    Zh, lone_m = inducing_descriptors(model, species, chunk)
    Ke = np.zeros(M)
    G = np.zeros((M, N, 3))     # d Ke[m] / d xyz
    W = np.zeros((M, 3, 3))
    xi = model.xi
    for k0 in range(0, N, chunk):
        idx = np.arange(k0, min(N, k0 + chunk))
        envs_r, envs_b, envs_j = [], [], []
        for i in idx:
            sl = slice(first[i], first[i + 1])
            envs_r.append(displacements(pos, cell, i, J[sl], S[sl]))
            envs_b.append(numbers[J[sl]])
            envs_j.append(J[sl])
        R, Zb, mask = pad_environments(envs_r, envs_b)
        P, aux = descriptor_batch(model, species, R, Zb, mask, want_aux=True)
        lone_c = ~mask.any(axis=1)
        K, dot, valid = kernel_from_descriptors(model, P, numbers[idx], lone_c, Zh, lone_m)
        Ke += K.sum(axis=0)
        Gm = np.where(valid, xi * _powxi(dot, xi - 1), 0.0)
        Zflat = Zh.reshape(M, -1)
        for m in range(M):
            if not Gm[:, m].any():
                continue
            g = (Gm[:, m : m + 1] * Zflat[m][None, :]).reshape(P.shape)
            dR = descriptor_backward(model, P, aux, g)       # dk/dr per pair
            for b, i in enumerate(idx):
                nn = len(envs_j[b])
                if nn == 0:
                    continue
                gb = dR[b, :nn]
                G[m, i] -= gb.sum(axis=0)                     # universal.py:147-148: g[i] -= f ; g[j] += f
                if scatter_quirk:
                    G[m][envs_j[b]] += gb                     # non-accumulating, like torch's g[j] += f
                else:
                    np.add.at(G[m], envs_j[b], gb)
                W[m] += envs_r[b].T @ gb
    Kf = -G.reshape(M, 3 * N).T
    Kv = W.reshape(M, 9)[:, [0, 4, 8, 5, 2, 1]].T
    return Ke, Kf, Kv


def covloss(model, K, numbers, alpha=None):
    """calculator/active.py:781-804.  ``alpha`` = self kernel k(x,x) per atom: the
    reference divides by it whenever it is not identically 1 (active.py:784-791), which
    yields NaN for centres excluded through ``a``/``a_not``."""
    b = model.choli @ K.T
    c = (b * b).sum(axis=0)
    if alpha is not None and not np.allclose(alpha, 1.0):
        with np.errstate(invalid="ignore", divide="ignore"):
            c = c / alpha
    beta = np.sqrt(np.clip(1 - c, 0.0, None))
    vs = np.array([model.vscale.get(int(z), np.inf) for z in numbers], dtype=float)
    with np.errstate(invalid="ignore"):
        return beta * np.sqrt(vs)


# --------------------------------------------------------------------------
# the reference's own known-answer test  (descriptor/soap.py:124-187,488-532)
# --------------------------------------------------------------------------
def abs_series_soap(coo, lmax, nmax, rc, unit=None):
    """AbsSeriesSoap(lmax, nmax, PolyCut(rc)).forward(coo, grad=False)
    (descriptor/soap.py:124-187): no Gaussian, no nnl, no normalisation."""
    unit = rc / 3 if unit is None else unit
    tab = YlmTables(lmax)
    xyz = np.asarray(coo, dtype=float) / unit
    d = np.sqrt((xyz**2).sum(axis=-1))
    n2 = 2.0 * np.arange(nmax + 1)
    dr_true = unit * d
    r = np.where(dr_true < rc, 1.0, 0.0) * (1.0 - dr_true / rc) ** 2
    f = r[None] * d[None] ** n2[:, None]
    Y = ylm(xyz, tab, shear_flag(xyz))
    c = np.einsum("nj,jik->nik", f, Y)
    nnp = c[None] * c[:, None]
    return (nnp * tab.Yr).sum(axis=-1) + (nnp * tab.Yi).sum(axis=-2)


SOAP_GOLDEN_XYZ = np.array(
    [
        [0.175, 0.884, -0.87, 0.354, -0.082, 3.1],
        [-0.791, 0.116, 0.19, -0.832, 0.184, 0.0],
        [0.387, 0.761, 0.655, -0.528, 0.973, 0.0],
    ]
).T  # descriptor/soap.py:493-500
SOAP_GOLDEN_TARGET = np.array(
    [
        [[0.36174603, 0.39013356, 0.43448023], [0.39013356, 0.42074877, 0.46857549], [0.43448023, 0.46857549, 0.5218387]],
        [[0.2906253, 0.30558356, 0.33600938], [0.30558356, 0.3246583, 0.36077952], [0.33600938, 0.36077952, 0.40524778]],
        [[0.16241845, 0.18307552, 0.20443194], [0.18307552, 0.22340802, 0.26811937], [0.20443194, 0.26811937, 0.34109511]],
    ]
)  # descriptor/soap.py:502-520  (p.permute(2,0,1): [l, n1, n2])
