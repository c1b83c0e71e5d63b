"""Reference results for bench.py's in-run parity block (run in the build container, where /root/reference is).

For each workload family of BASELINE.json a small periodic cell (fcc rep=4, 256 atoms) of the same lattice, species
mix and rattle is evaluated by the UNMODIFIED reference (oracle/ref_bench.py: ActiveCalculator.calculate in prediction
mode on 8 processes) with the same frozen synthetic model bench.py builds (autoforce_b200/synth.py, + choli = 0.5 I).
Only the structure and the reference's outputs are stored (the model is regenerated from its seed):

    tests/golden/bench_<workload>_sample.npz : pos, cell, numbers, energy, forces, stress, covloss

bench.py evaluates the same cell on the GPU(s) after the timed region of every run and prints the deltas.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_bench  # noqa: E402

if __name__ == "__main__":
    for wl in sys.argv[1:] or ["c2", "c3", "c4", "c5"]:
        work = tempfile.mkdtemp(prefix=f"fixture_{wl}_")
        out = ref_bench.run(wl, steps=1, warmup=0, procs=min(8, os.cpu_count() or 1), rep=4, variants=1, keep_dir=work)
        z = np.load(os.path.join(work, "last_step.npz"))
        dst = os.path.join(ROOT, "tests", "golden", f"bench_{wl}_sample.npz")
        np.savez_compressed(dst, **{k: z[k] for k in z.files}, rep=np.array(4), M=np.array(out["M"]))
        print(wl, out["value"], "atom-steps/s on", out["procs"], "processes ->", dst, flush=True)
        shutil.rmtree(work, ignore_errors=True)
