"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (/root/reference/theforce) on seeded synthetic inputs.

Runs only in the build container (needs /root/reference + oracle/shims):

    python tests/golden/make_golden.py [case ...]

Each ``<case>.npz`` holds the inputs (structure, kernel hyper-parameters, the
inducing LCEs as CSR, mu, mean weights, choli, vscale) and the reference's own
outputs: results['energy'/'forces'/'stress'] of ActiveCalculator.calculate
(calculator/active.py:425-611), the kernel matrix ``calc.cov``, ``get_covloss()``
(active.py:781-804), the neighbour list the reference consumed, and dense copies
of a few cached descriptors (``loc.kern_0_value``).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_runner as rr  # noqa: E402
from oracle import sgpr_oracle as o  # noqa: E402


def fcc(rep, Zs, sigma, seed, a0=3.61, cell_scale=None):
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    rep = (rep,) * 3 if np.isscalar(rep) else rep
    cells = np.array([[i, j, k] for i in range(rep[0]) for j in range(rep[1]) for k in range(rep[2])])
    pos = (cells[:, None, :] + base[None]).reshape(-1, 3) * a0
    if sigma > 0:
        pos = pos + rng.normal(0, sigma, pos.shape)
    numbers = rng.choice(np.asarray(Zs), len(pos))
    cell = np.diag([a0 * r for r in rep]).astype(float)
    return pos, cell, numbers


def pick_inducing(src_pos, src_cell, src_pbc, src_num, rc, M, seed):
    """Per-species balanced draw of M environments from a source structure."""
    rng = np.random.default_rng(seed)
    first, J, S = o.neighbor_list(src_pos, src_cell, src_pbc, rc)
    Zs = np.unique(src_num)
    sel = []
    for k, z in enumerate(Zs):
        idx = np.nonzero(src_num == z)[0]
        m = M // len(Zs) + (1 if k < M % len(Zs) else 0)
        sel += list(rng.choice(idx, min(m, len(idx)), replace=False))
    envs = []
    for i in sel:
        sl = slice(first[i], first[i + 1])
        envs.append((int(src_num[i]), o.displacements(src_pos, src_cell, i, J[sl], S[sl]), src_num[J[sl]].copy()))
    return envs


def case_inputs(name):
    """-> dict(kernel=..., pos, cell, pbc, numbers, envs, mu, mean_w, choli, vscale)."""
    rng = np.random.default_rng(abs(hash(name)) % (2**31) if False else sum(map(ord, name)))
    if name == "cu108_sesoap":
        kern = dict(kind="sesoap", lmax=3, nmax=3, xi=4, rc=6.0)
        pos, cell, num = fcc(3, [29], 0.1, 0)
        sp, sc, sn = fcc(3, [29], 0.15, 100)
        envs = pick_inducing(sp, sc, True, sn, 6.0, 20, 1)
        pbc = [True] * 3
    elif name == "cu108_perfect":
        # perfect lattice: the z-axis shear (ylm.py:10-23) fires in every environment
        kern = dict(kind="sesoap", lmax=3, nmax=3, xi=4, rc=6.0)
        pos, cell, num = fcc(3, [29], 0.0, 0)
        sp, sc, sn = fcc(3, [29], 0.15, 100)
        envs = pick_inducing(sp, sc, True, sn, 6.0, 8, 2)
        pbc = [True] * 3
    elif name == "lipso108":
        kern = dict(kind="sesoap", lmax=3, nmax=3, xi=4, rc=6.0)
        pos, cell, num = fcc(3, [3, 15, 16, 8], 0.1, 3)
        sp, sc, sn = fcc(3, [3, 15, 16, 8], 0.15, 103)
        envs = pick_inducing(sp, sc, True, sn, 6.0, 24, 4)
        pbc = [True] * 3
    elif name == "tric_oh":
        # triclinic cell, H (radius 0.5, DefaultRadii) + O, some atoms outside the cell
        kern = dict(kind="sesoap", lmax=2, nmax=3, xi=4, rc=4.5)
        pos, cell, num = fcc((2, 2, 3), [1, 8], 0.12, 5, a0=3.2)
        cell = cell + np.array([[0.0, 0.0, 0.0], [0.9, 0.0, 0.0], [0.5, -0.7, 0.0]])
        pos[::5] += cell[0]
        pos[1::7] -= 2 * cell[2]
        sp, sc, sn = fcc(3, [1, 8], 0.2, 105, a0=3.2)
        envs = pick_inducing(sp, sc, True, sn, 4.5, 12, 6)
        pbc = [True] * 3
    elif name == "universal_2sp":
        kern = dict(kind="universal", lmax=3, nmax=2, xi=4, rc=5.0)
        pos, cell, num = fcc((2, 3, 2), [14, 8], 0.1, 7, a0=3.4)
        sp, sc, sn = fcc(3, [14, 8], 0.15, 107, a0=3.4)
        envs = pick_inducing(sp, sc, True, sn, 5.0, 12, 8)
        pbc = [True] * 3
    elif name == "cluster_lone":
        # non-periodic cluster, zero cell, one isolated atom + one neighbour-less inducing LCE
        kern = dict(kind="sesoap", lmax=3, nmax=3, xi=4, rc=4.0)
        pos, cell, num = fcc(2, [29, 47], 0.1, 9, a0=3.8)
        pos = np.vstack([pos, [[40.0, 40.0, 40.0]], [[-30.0, 5.0, 2.0]]])
        num = np.concatenate([num, [29, 47]])
        cell = np.zeros((3, 3))
        pbc = [False] * 3
        sp, sc, sn = fcc(2, [29, 47], 0.15, 109, a0=3.8)
        envs = pick_inducing(sp, np.zeros((3, 3)), False, sn, 4.0, 10, 10)
        envs.append((29, np.zeros((0, 3)), np.zeros(0, dtype=np.int64)))
    elif name == "slab_ttf":
        kern = dict(kind="sesoap", lmax=3, nmax=3, xi=4, rc=5.0)
        pos, cell, num = fcc((2, 2, 2), [13], 0.1, 11, a0=4.05)
        cell[2, 2] += 12.0
        pbc = [True, True, False]
        sp, sc, sn = fcc(2, [13], 0.15, 111, a0=4.05)
        envs = pick_inducing(sp, sc, True, sn, 5.0, 8, 12)
    elif name == "highres_l6n8":
        kern = dict(kind="sesoap", lmax=6, nmax=8, xi=4, rc=7.0)
        pos, cell, num = fcc(2, [29], 0.1, 13)
        sp, sc, sn = fcc(2, [29], 0.15, 113)
        envs = pick_inducing(sp, sc, True, sn, 7.0, 6, 14)
        pbc = [True] * 3
    elif name == "anot_xi2":
        # species 8 excluded as a centre (EqAll exceptions), exponent 2
        kern = dict(kind="sesoap", lmax=3, nmax=3, xi=2, rc=5.0, a_not=[8])
        pos, cell, num = fcc((2, 2, 2), [3, 8], 0.1, 15, a0=3.9)
        sp, sc, sn = fcc(2, [3, 8], 0.15, 115, a0=3.9)
        envs = pick_inducing(sp, sc, True, sn, 5.0, 10, 16)
        pbc = [True] * 3
    elif name == "subse_3sp":
        # default_kernel(species=[3, 8]): one SubSeSoapKernel per central species; S atoms (16) are present in
        # the structure but are neither centres nor counted as neighbours
        kern = dict(kind="subsesoap", lmax=3, nmax=2, xi=4, rc=5.0, species=[3, 8])
        pos, cell, num = fcc((2, 2, 3), [3, 8, 16], 0.1, 17, a0=3.9)
        sp, sc, sn = fcc(3, [3, 8, 16], 0.15, 117, a0=3.9)
        envs = [e for e in pick_inducing(sp, sc, True, sn, 5.0, 15, 18) if e[0] in (3, 8)]
        pbc = [True] * 3
    elif name == "subse_lone":
        # SubSeSoapKernel list + neighbour-less atoms: every kernel of the list adds the lone-lone term
        # (similarity.py:41-43,94-103), so k(lone, lone) = number of kernels
        kern = dict(kind="subsesoap", lmax=2, nmax=2, xi=4, rc=4.0, species=[3, 8])
        pos, cell, num = fcc((2, 2, 2), [3, 8], 0.1, 19, a0=3.9)
        cell = cell + np.diag([30.0, 0.0, 0.0])
        pos = np.vstack([pos, [[20.0, 1.0, 1.0]], [[30.0, 5.0, 2.0]]])
        num = np.concatenate([num, [3, 8]])
        sp, sc, sn = fcc(2, [3, 8], 0.15, 119, a0=3.9)
        envs = [e for e in pick_inducing(sp, sc, True, sn, 4.0, 10, 20)]
        envs.append((3, np.zeros((0, 3)), np.zeros(0, dtype=np.int64)))
        envs.append((8, np.zeros((0, 3)), np.zeros(0, dtype=np.int64)))
        pbc = [True] * 3
    elif name == "hetero_2sp":
        # HeterogeneousSoapKernel(DotProd()**4, a, b, ...) per central species, descriptor NormalizedSoap(HeteroSoap):
        # one length unit rc/3 for all species (descriptor/soap.py:13-26)
        kern = dict(kind="heterosoap", lmax=3, nmax=2, xi=4, rc=4.5, species=[14, 8])
        pos, cell, num = fcc((2, 2, 3), [14, 8], 0.1, 21, a0=3.4)
        sp, sc, sn = fcc(3, [14, 8], 0.15, 121, a0=3.4)
        envs = pick_inducing(sp, sc, True, sn, 4.5, 12, 22)
        pbc = [True] * 3
    elif name == "two_kernels":
        # two similarity kernels with different lmax / nmax / exponent / cutoff summed in one model; an isolated atom
        # and an atom whose only neighbour lies between the two cutoffs
        kern = dict(kind="multi", lmax=3, nmax=3, xi=4, rc=5.5,
                    kernels=[dict(kind="sesoap", lmax=3, nmax=3, xi=4, rc=5.5), dict(kind="sesoap", lmax=2, nmax=2, xi=2, rc=3.8)])
        pos, cell, num = fcc((2, 2, 2), [3, 8], 0.1, 23, a0=3.9)
        cell = cell + np.diag([40.0, 0.0, 0.0])
        pos = np.vstack([pos, [[25.0, 1.0, 1.0]], [[35.0, 4.0, 2.0]], [[39.5, 4.0, 2.0]]])
        num = np.concatenate([num, [3, 8, 3]])
        sp, sc, sn = fcc(2, [3, 8], 0.15, 123, a0=3.9)
        envs = pick_inducing(sp, sc, True, sn, 5.5, 10, 24)
        envs.append((3, np.zeros((0, 3)), np.zeros(0, dtype=np.int64)))
        envs.append((8, np.array([[4.5, 0.0, 0.0]]), np.array([3], dtype=np.int64)))
        pbc = [True] * 3
    elif name == "unnorm_2sp":
        # un-normalised descriptor (normalize=False): the power spectrum itself enters the kernel (sesoap.py:249-251 off)
        kern = dict(kind="sesoap", lmax=2, nmax=2, xi=2, rc=4.5, normalize=False)
        pos, cell, num = fcc((2, 2, 2), [3, 8], 0.1, 25, a0=3.9)
        sp, sc, sn = fcc(2, [3, 8], 0.15, 125, a0=3.9)
        envs = pick_inducing(sp, sc, True, sn, 4.5, 10, 26)
        pbc = [True] * 3
    elif name == "afixed_2sp":
        # a kernel restricted to ONE central species (a=8): other centres give zero rows (universal.py:100-107)
        kern = dict(kind="sesoap", lmax=3, nmax=2, xi=4, rc=5.0, a=8)
        pos, cell, num = fcc((2, 2, 2), [3, 8], 0.1, 27, a0=3.9)
        sp, sc, sn = fcc(2, [3, 8], 0.15, 127, a0=3.9)
        envs = pick_inducing(sp, sc, True, sn, 5.0, 10, 28)
        pbc = [True] * 3
    else:
        raise KeyError(name)
    M = len(envs)
    Zs = np.unique(np.concatenate([num] + [e[2] for e in envs] + [[e[0] for e in envs]]).astype(np.int64))
    mu = rng.normal(0, 1, M) * 0.1
    if not kern.get("normalize", True):
        mu = mu * 1e9   # the un-normalised power spectrum is tiny (nnl ~ 1e-3 ... 1e-6 per entry): K ~ 1e-9
    A = rng.normal(0, 1, (M, M)) * 0.05
    choli = 0.5 * np.eye(M) + np.tril(A)
    mean_w = {int(z): -3.0 + 0.1 * k for k, z in enumerate(Zs)}
    vscale = {int(z): 1.0 + 0.25 * k for k, z in enumerate(np.unique([e[0] for e in envs]))}
    return dict(kernel=kern, pos=pos, cell=cell, pbc=pbc, numbers=num, envs=envs, mu=mu, mean_w=mean_w, choli=choli, vscale=vscale)


CASES = [
    "cu108_sesoap", "cu108_perfect", "lipso108", "tric_oh", "universal_2sp", "cluster_lone", "slab_ttf", "highres_l6n8", "anot_xi2",
    "subse_3sp", "subse_lone", "hetero_2sp", "two_kernels", "unnorm_2sp", "afixed_2sp",
]


def species_dense(desc, species, kind):
    """sparse [120,120,D] (index (Z_s2, Z_s1), sesoap.py:165-171) -> [S,S,D] blocks [s1,s2]."""
    S = len(species)
    out = np.zeros((S, S, desc.shape[-1]))
    for a, z1 in enumerate(species):
        for b, z2 in enumerate(species):
            out[a, b] = desc[z2, z1]
    return out


def run_case(name):
    c = case_inputs(name)
    k = c["kernel"]
    kern = rr.make_kernel(k["kind"], k["lmax"], k["nmax"], k["xi"], k["rc"], a_not=k.get("a_not", ()),
                          normalize=k.get("normalize", True), a=k.get("a"),
                          radii={"species": k["species"]} if k["kind"] in ("subsesoap", "heterosoap") else
                          {"kernels": k["kernels"]} if k["kind"] == "multi" else None)
    model = rr.synth_model(kern, c["envs"], c["mu"], c["mean_w"], c["choli"], c["vscale"])
    t0 = time.time()
    want = [0, len(c["pos"]) // 2, len(c["pos"]) - 1]
    ref = rr.ref_predict(model, c["pos"], c["cell"], c["pbc"], c["numbers"], want_descriptors=want)
    # second evaluation with large weights (stress test for precision)
    import torch

    mu_big = c["mu"] * 1000.0
    model.mu = torch.as_tensor(mu_big)
    ref_big = rr.ref_predict(model, c["pos"], c["cell"], c["pbc"], c["numbers"])
    dt = time.time() - t0
    envs = c["envs"]
    ind_first = np.cumsum([0] + [len(e[2]) for e in envs]).astype(np.int64)
    species = np.unique(np.concatenate([c["numbers"]] + [e[2] for e in envs] + [[e[0] for e in envs]]).astype(np.int64))
    descs = {}
    zdesc = []
    if k["kind"] in ("subsesoap", "heterosoap", "multi"):
        zdesc = [np.zeros((1,))] * len(model.X)   # dense per-kernel caches: not stored (the oracle is pinned through K)
    else:
        for a, d in ref["descriptors"].items():
            if d is not None:
                descs[f"desc_{a}"] = species_dense(d, species, k["kind"])
        # descriptors of the inducing LCEs cached by the reference (loc.kern_0_value)
        for loc in model.X:
            v = loc.__dict__.get("kern_0_value")
            zdesc.append(np.zeros((len(species), len(species), kern.dim)) if v is None else species_dense(v.detach().to_dense().numpy(), species, k["kind"]))
    meta = dict(kernel=k, pbc=[bool(x) for x in c["pbc"]], mean_w={str(z): w for z, w in c["mean_w"].items()},
                vscale={str(z): v for z, v in c["vscale"].items()}, species=[int(z) for z in species],
                unit=(float(kern.descriptor.unit) if k["kind"] == "universal" else
                      float(kern[0].descriptor.soap.unit) if k["kind"] == "heterosoap" else None),
                a_only=(k["species"] if k["kind"] in ("subsesoap", "heterosoap") else [k["a"]] if k.get("a") is not None else []),
                b_only=(k["species"] if k["kind"] in ("subsesoap", "heterosoap") else []),
                lone_weight=(len(k["species"]) if k["kind"] in ("subsesoap", "heterosoap") else
                             len(k["kernels"]) if k["kind"] == "multi" else 1),
                generator="tests/golden/make_golden.py", reference="theforce v2021.09", ref_seconds=round(dt, 2))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        meta=json.dumps(meta),
        pos=c["pos"], cell=c["cell"], numbers=c["numbers"].astype(np.int64),
        ind_first=ind_first, ind_Z=np.array([e[0] for e in envs], dtype=np.int64),
        ind_r=np.concatenate([np.asarray(e[1]).reshape(-1, 3) for e in envs]),
        ind_b=np.concatenate([np.asarray(e[2], dtype=np.int64).reshape(-1) for e in envs]),
        mu=c["mu"], choli=c["choli"],
        energy=ref["energy"], forces=ref["forces"], stress=ref["stress"], K=ref["K"], covloss=ref["covloss"],
        energy_big=ref_big["energy"], forces_big=ref_big["forces"], stress_big=ref_big["stress"], mu_big=mu_big,
        nl_first=ref["nl_first"], nl_j=ref["nl_j"].astype(np.int32), nl_S=ref["nl_S"].astype(np.int8),
        ind_desc=np.stack(zdesc), **descs,
    )
    print(f"{name}: N={len(c['pos'])} M={len(envs)} pairs={len(ref['nl_j'])} E={float(ref['energy']):.6f} "
          f"|F|max={np.abs(ref['forces']).max():.4f} ref_time={dt:.1f}s", flush=True)


if __name__ == "__main__":
    rr.import_reference()
    for name in sys.argv[1:] or CASES:
        run_case(name)
