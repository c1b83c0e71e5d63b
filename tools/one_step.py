"""Developer tool: a few steps of one workload (for ncu captures)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autoforce_b200 as ab
from autoforce_b200 import synth
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = synth.WORKLOADS[wl]
model = synth.synth_model(w["Zs"], w["M"], 1, lmax=w["lmax"], nmax=w["nmax"], rc=w["rc"], with_choli=bool(os.environ.get("BETA")))
pos, cell, numbers = synth.fcc(w["rep"], w["Zs"], 0.1, 0)
eng = ab.SgprEngine(model, species=w["Zs"])
for it in range(steps):
    E = eng.predict(pos, numbers, cell, True, want_beta=bool(os.environ.get('BETA')))[0]
print(wl, len(pos), E)
