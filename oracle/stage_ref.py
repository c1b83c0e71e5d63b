"""Stage the UNMODIFIED reference package for runs outside the build container.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is pure Python
(`setup.py:4-17` is `find_packages()` only), so "building" it is a file copy:

    /root/reference/theforce/**.py  ->  oracle/_ref/theforce/

`oracle/_ref/` is git-ignored (never part of the history) but NOT gpurun-ignored, so it
travels to the GPU box next to the built `.so` files.  Together with `oracle/shims`
(minimal `ase` / `mpi4py` stand-ins, ASE is not in the image) it lets
  * `bench.py --impl reference` time the reference's own `ActiveCalculator.calculate`
    on the GPU box's host cores (`cpu_baseline.kind = "reference"`), and
  * the `-m gpu` tests run the reference's on-the-fly control flow side by side with the
    B200 plugin (`tests/test_reference_plugin.py`).
Nothing under `autoforce_b200/` reads this directory; `__graft_entry__.build()` calls
`stage()` when `/root/reference` is present.  A manifest with the sha256 of every staged
file is written so that "unmodified" can be checked.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(HERE, "_ref")
SOURCE_ROOT = os.environ.get("AUTOFORCE_REFERENCE", "/root/reference")
SKIP_DIRS = {"__pycache__", "deprecated"}   # dead code, imported only inside the reference's own test functions


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def staged_available():
    return os.path.isfile(os.path.join(STAGED_ROOT, "theforce", "__init__.py"))


def stage(force=False):
    """Copy the reference package into oracle/_ref/ (no-op when the source tree is absent).  Returns the staged
    root or None."""
    src = os.path.join(SOURCE_ROOT, "theforce")
    if not os.path.isdir(src):
        return STAGED_ROOT if staged_available() else None
    dst = os.path.join(STAGED_ROOT, "theforce")
    manifest_path = os.path.join(STAGED_ROOT, "MANIFEST.json")
    manifest = {}
    for root, dirs, files in os.walk(src):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        rel = os.path.relpath(root, src)
        for f in sorted(files):
            if f.endswith(".py"):
                manifest[os.path.normpath(os.path.join(rel, f))] = _sha(os.path.join(root, f))
    if not force and os.path.isfile(manifest_path):
        try:
            if json.load(open(manifest_path)).get("files") == manifest and staged_available():
                return STAGED_ROOT
        except Exception:
            pass
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for rel in manifest:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
    with open(manifest_path, "w") as f:
        json.dump({"source": SOURCE_ROOT, "package": "theforce", "files": manifest}, f, indent=1, sort_keys=True)
    return STAGED_ROOT


def verify():
    """True when every staged file still has the sha256 recorded at staging time."""
    mp = os.path.join(STAGED_ROOT, "MANIFEST.json")
    if not os.path.isfile(mp):
        return False
    files = json.load(open(mp))["files"]
    return all(os.path.isfile(os.path.join(STAGED_ROOT, "theforce", rel)) and _sha(os.path.join(STAGED_ROOT, "theforce", rel)) == sha
               for rel, sha in files.items())


if __name__ == "__main__":
    print(stage(force=True), "verified" if verify() else "NOT verified")
