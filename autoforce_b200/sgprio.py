"""ASE-free reader / writer for the inducing-LCE blocks of an AutoForce ``.sgpr`` tape.

Format (theforce/io/sgprio.py:16-54,57-143): an append-only text log of blocks

    start: local            start: atoms                include: other.sgpr
    <Z>                     <extxyz frame>              (relative to the including file,
    <Z_j> <x> <y> <z>       end: atoms                   recursive, each file read once)
    ...
    end: local

``read_lces`` returns the local chemical environments in tape order as (Z, r[nn,3], b[nn]) -- exactly
what ``SgprModel.from_envs`` takes -- and skips ``atoms``/``params`` blocks (training data stays with
the reference's trainer).  The tape does not carry the weights mu: they come from the reference's fit.
"""
from __future__ import annotations

import os

import numpy as np


def write_lce(f, Z, r, b):
    """Same text as theforce.io.sgprio.write_lce (sgprio.py:16-22) inside a ``local`` block."""
    f.write("\nstart: local\n")
    f.write(f"{int(Z):4d}\n")
    for s, x in zip(np.asarray(b).reshape(-1), np.asarray(r, dtype=float).reshape(-1, 3)):
        f.write("{:4d} {:16.8f} {:16.8f} {:16.8f}\n".format(int(s), *x.tolist()))
    f.write("end: local\n")


def read_lces(path, _exclude=None):
    path = os.path.abspath(os.path.expanduser(os.path.expandvars(path)))
    exclude = [] if _exclude is None else _exclude
    if path in exclude or not os.path.isfile(path):
        return []
    exclude.append(path)
    envs, on, typ, blk = [], False, None, []
    with open(path) as f:
        for line in f:
            if not on:
                if line.startswith("start:"):
                    on, typ, blk = True, line.split()[-1], []
                elif line.startswith("include:"):
                    inc = os.path.expanduser(os.path.expandvars(line.split()[-1]))
                    if not os.path.isabs(inc):
                        inc = os.path.join(os.path.dirname(path), inc)
                    envs.extend(read_lces(inc, exclude))
            elif line.startswith("end:"):
                if line.split()[-1] != typ:
                    raise ValueError(f"{path}: block '{typ}' closed by '{line.strip()}'")
                on = False
                if typ == "local":
                    Z = int(blk[0].strip())
                    rows = [ln.split() for ln in blk[1:] if ln.strip()]
                    b = np.array([int(t[0]) for t in rows], dtype=np.int32)
                    r = np.array([[float(v) for v in t[1:4]] for t in rows], dtype=np.float64).reshape(-1, 3)
                    envs.append((Z, r, b))
            else:
                blk.append(line)
    return envs
