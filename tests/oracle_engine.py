"""TEST INFRASTRUCTURE: an ``SgprEngine`` look-alike on the numpy oracle (CPU).

It exists so that the `not gpu` suite can drive the host logic of
``autoforce_b200.reference_plugin.B200ActiveCalculator`` (engine synchronisation with the
reference model, the ``model.gp.kern`` dispatch, lazy ``TorchAtoms``, autograd seam) in the
build container, where there is no GPU.  It is never importable from the product: the
tests monkey-patch ``reference_plugin.SgprEngine`` with it.  On the GPU box the same tests
run with the real engine.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import sgpr_oracle as o


def to_oracle(model):
    first = model.ind_first
    return o.OracleModel(
        lmax=model.lmax, nmax=model.nmax, xi=model.xi, rc=model.rc, radii=model.radii, default_radius=model.default_radius,
        ind_Z=model.ind_Z.astype(np.int64), ind_r=[model.ind_r[first[m]:first[m + 1]] for m in range(model.M)],
        ind_b=[model.ind_b[first[m]:first[m + 1]].astype(np.int64) for m in range(model.M)], mu=model.mu,
        mean_w=model.mean_w, choli=model.choli, vscale=model.vscale, normalize=model.normalize, a_not=model.a_not,
        a_only=model.a_only, b_only=model.b_only, lone_weight=max(model.lone_weight, 0.0))


class _Cov(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, lll, eng, numbers, pbc):
        out = o.predict(to_oracle(eng.model), xyz.detach().numpy(), lll.detach().numpy(), pbc, numbers, want_K=True)
        ctx.eng, ctx.args = eng, (xyz.detach().numpy().copy(), lll.detach().numpy().copy(), pbc, numbers)
        return torch.as_tensor(out["K"])

    @staticmethod
    def backward(ctx, gK):
        pos, cell, pbc, numbers = ctx.args
        m = to_oracle(ctx.eng.model)
        m.mean_w = {}
        out = o.predict(m, pos, cell, pbc, numbers, row_weights=gK.detach().numpy())
        F, W = out["forces"], out["virial"]
        Y = W + pos.T @ F                       # = cell^T (sum_pairs S (x) g), see csrc/api.cu sgpr_kernel_backward
        cc = o.complete_cell(cell)
        C = np.linalg.solve(cc.T, Y)
        C[~np.broadcast_to(np.asarray(pbc), (3,))] = 0.0
        return torch.as_tensor(-F), torch.as_tensor(C), None, None, None


class OracleEngine:
    instances = 0

    def __init__(self, model, species=None, device=0):
        self.model = model
        self.species = sorted(set(model.species()) | set(int(z) for z in (species or [])))
        self.device = device
        self.calls = {"cov": 0, "predict": 0, "neighbors": 0, "append": 0, "set_weights": 0}
        OracleEngine.instances += 1

    def close(self):
        pass

    def neighbors(self, pos, numbers, cell, pbc):
        self.calls["neighbors"] += 1
        return o.neighbor_list(np.asarray(pos, dtype=float), o.complete_cell(np.asarray(cell, dtype=float).reshape(3, 3)),
                               np.asarray(pbc), self.model.rc)

    def cov(self, xyz, lll, numbers, pbc):
        self.calls["cov"] += 1
        return _Cov.apply(xyz, lll, self, np.asarray(numbers), np.asarray(pbc))

    def predict(self, pos, numbers, cell, pbc, rank=0, world=1, want_beta=False, out_forces=None):
        self.calls["predict"] += 1
        out = o.predict(to_oracle(self.model), pos, cell, pbc, numbers, want_beta=want_beta)
        res = (float(out["energy"]), out["forces"], out["virial"], np.ones(len(numbers), bool))
        return res + (out["beta"],) if want_beta else res

    def append_inducing(self, envs, mu, choli=None):
        self.calls["append"] += 1
        m = self.model
        first = np.cumsum([0] + [len(e[2]) for e in envs])
        m.ind_first = np.concatenate([m.ind_first, m.ind_first[-1] + first[1:]])
        m.ind_Z = np.concatenate([m.ind_Z, np.array([e[0] for e in envs], dtype=np.int32)])
        m.ind_r = np.concatenate([m.ind_r] + [np.asarray(e[1], dtype=float).reshape(-1, 3) for e in envs])
        m.ind_b = np.concatenate([m.ind_b] + [np.asarray(e[2], dtype=np.int32).reshape(-1) for e in envs])
        m.mu = np.asarray(mu, dtype=float).copy()
        m.choli = None if choli is None else np.asarray(choli, dtype=float).copy()

    def set_weights(self, mu=None, mean_w=None, choli=None, vscale=None):
        self.calls["set_weights"] += 1
        m = self.model
        if mu is not None:
            m.mu = np.asarray(mu, dtype=float).copy()
        if mean_w is not None:
            m.mean_w = dict(mean_w)
        if choli is not None:
            m.choli = np.asarray(choli, dtype=float).copy()
        if vscale is not None:
            m.vscale = dict(vscale)
