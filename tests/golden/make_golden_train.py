"""Golden vectors of the TRAINING-time kernels from the reference itself.

For a few of the make_golden.py cases this writes ``<case>_train.npz`` with, for the one structure against
the inducing set X (regression/gppotential.py:63-77):

  Ke  [M]     energy_energy([atoms], X)
  Kf_analytic [3N, M], Kv_analytic [6, M]
              forces_energy / virial_energy as the reference evaluates them: its hand-written descriptor
              gradients (similarity/universal.py:124-183, descriptor/sesoap.py:203-238, descriptor/ylm.py:192-224)
  Kf_autograd [3N, M], Kv_autograd [6, M]
              the same derivatives taken by torch.autograd through the reference's forward pass, exactly as
              ActiveCalculator.grads does for the forces (calculator/active.py:587-611), one inducing LCE at a time.

The two differ by up to ~1e-4 relative: the reference's analytic Ylm gradient uses a coefficient table rounded
to float32 (descriptor/ylm.py:103-106, ``.float()``) whose error is amplified near the z axis; its autograd
derivative is exact to rounding.  Run in the build container only:  python tests/golden/make_golden_train.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import ref_runner as rr  # noqa: E402

CASES = ["lipso108", "cu108_sesoap", "universal_2sp", "cluster_lone", "tric_oh"]


def run_case(name):
    import make_golden as mg
    import torch
    from theforce.descriptor.atoms import TorchAtoms

    c = mg.case_inputs(name)
    k = c["kernel"]
    kern = rr.make_kernel(k["kind"], k["lmax"], k["nmax"], k["xi"], k["rc"], a_not=k.get("a_not", ()))
    model = rr.synth_model(kern, c["envs"], c["mu"], c["mean_w"], c["choli"], c["vscale"])
    M, N = len(model.X), len(c["numbers"])

    def fresh():
        return TorchAtoms(ase_atoms=rr.ase_atoms(c["pos"], c["cell"], c["pbc"], c["numbers"]), cutoff=model.cutoff,
                          descriptors=model.descriptors)

    ta = fresh()                                  # descriptors with the analytic gradients cached
    Ke = model.gp.kern([ta], model.X, cov="energy_energy").detach().numpy().reshape(-1)
    Kf = model.gp.kern([ta], model.X, cov="forces_energy").detach().numpy()
    Kv = model.gp.kern([ta], model.X, cov="virial_energy").detach().numpy()
    tb = fresh()
    tb.update(posgrad=True, cellgrad=True, forced=True, dont_save_grads=True)
    K = model.gp.kern(tb, model.X)                # [N, M], differentiable
    Kf_a = np.zeros((3 * N, M))
    Kv_a = np.zeros((6, M))
    for m in range(M):
        col = K[:, m].sum()
        if not col.requires_grad:
            continue
        gx, gl = torch.autograd.grad(col, [tb.xyz, tb.lll], retain_graph=True, allow_unused=True)
        gx = torch.zeros_like(tb.xyz) if gx is None else gx
        gl = torch.zeros_like(tb.lll) if gl is None else gl
        Kf_a[:, m] = -gx.numpy().reshape(-1)
        # pair virial = sum_i x_i (x) dk/dx_i + sum_k L_k (x) dk/dL_k   (calculator/active.py:603-605 without -1/V)
        w = (tb.xyz.detach()[:, :, None] * gx[:, None, :]).sum(dim=0) + (tb.lll.detach()[:, :, None] * gl[:, None, :]).sum(dim=0)
        Kv_a[:, m] = w.numpy().reshape(-1)[[0, 4, 8, 5, 2, 1]]
    np.savez_compressed(os.path.join(HERE, name + "_train.npz"), Ke=Ke, Kf_analytic=Kf, Kv_analytic=Kv, Kf_autograd=Kf_a,
                        Kv_autograd=Kv_a)
    sf, sv = np.abs(Kf_a).max(), max(np.abs(Kv_a).max(), 1e-300)
    print(f"{name}: N={N} M={M} |Kf|max={sf:.4f} analytic-vs-autograd: Kf {np.abs(Kf - Kf_a).max():.2e} "
          f"Kv {np.abs(Kv - Kv_a).max():.2e} (|Kv|max={sv:.3f})", flush=True)


if __name__ == "__main__":
    rr.import_reference()
    for name in sys.argv[1:] or CASES:
        run_case(name)
