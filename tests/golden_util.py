"""Load the golden fixtures written by tests/golden/make_golden.py."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.endswith("_train") and not n.startswith("bench_")]


def train_cases():
    """Cases with training-time kernels (tests/golden/make_golden_train.py)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*_train.npz")))
    return [n[: -len("_train")] for n in names]


def load_train(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + "_train.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["meta"] = json.loads(str(g["meta"]))
    first = g["ind_first"]
    g["envs_r"] = [g["ind_r"][first[m] : first[m + 1]] for m in range(len(first) - 1)]
    g["envs_b"] = [g["ind_b"][first[m] : first[m + 1]] for m in range(len(first) - 1)]
    return g


def golden_radii(meta):
    """Z -> length unit, as the reference's kernel had it (DefaultRadii: H 0.5, else 1.0,
    descriptor/sesoap.py:84-99; UniversalSoap: one unit for all, soap.py:724-728)."""
    if meta["kernel"]["kind"] in ("universal", "heterosoap"):   # HeteroSoap: one unit as well (descriptor/soap.py:20-23)
        return {int(z): float(meta["unit"]) for z in meta["species"]}, float(meta["unit"])
    return {1: 0.5}, 1.0


def oracle_model(g, big=False, kernel=None):
    """OracleModel of a golden case; for a case with a list of different kernels (kind "multi") a list of them."""
    from oracle.sgpr_oracle import OracleModel

    k = g["meta"]["kernel"] if kernel is None else kernel
    if k["kind"] == "multi":
        return [oracle_model(g, big, dict(kk)) for kk in k["kernels"]]
    radii, default = golden_radii(g["meta"])
    return OracleModel(
        lmax=k["lmax"], nmax=k["nmax"], xi=k["xi"], rc=k["rc"], radii=radii, default_radius=default,
        ind_Z=g["ind_Z"], ind_r=g["envs_r"], ind_b=g["envs_b"], mu=g["mu_big"] if big else g["mu"],
        mean_w={int(z): w for z, w in g["meta"]["mean_w"].items()}, choli=g["choli"],
        vscale={int(z): v for z, v in g["meta"]["vscale"].items()}, a_not=tuple(k.get("a_not", ())),
        normalize=bool(k.get("normalize", True)),
        a_only=tuple(g["meta"].get("a_only", ())), b_only=tuple(g["meta"].get("b_only", ())),
        lone_weight=float(g["meta"].get("lone_weight", 1)) if kernel is None else 1.0,
    )


def b200_model(g, big=False, kernel=None):
    """The same golden model as an autoforce_b200.SgprModel (product-side container); a list of them for a case with
    a list of different kernels (kind "multi")."""
    import autoforce_b200 as ab

    k = g["meta"]["kernel"] if kernel is None else kernel
    if k["kind"] == "multi":
        return [b200_model(g, big, dict(kk)) for kk in k["kernels"]]
    radii, default = golden_radii(g["meta"])
    return ab.SgprModel(
        lmax=k["lmax"], nmax=k["nmax"], xi=k["xi"], rc=k["rc"], kind="universal" if k["kind"] in ("universal", "heterosoap") else "sesoap",
        radii=radii, default_radius=default, a_not=tuple(k.get("a_not", ())), a_only=tuple(g["meta"].get("a_only", ())),
        normalize=bool(k.get("normalize", True)),
        b_only=tuple(g["meta"].get("b_only", ())), ind_Z=g["ind_Z"], ind_first=g["ind_first"], ind_r=g["ind_r"], ind_b=g["ind_b"],
        mu=g["mu_big"] if big else g["mu"], mean_w={int(z): w for z, w in g["meta"]["mean_w"].items()},
        choli=g["choli"], vscale={int(z): v for z, v in g["meta"]["vscale"].items()},
        lone_weight=float(g["meta"].get("lone_weight", 1)) if kernel is None else 1.0)
