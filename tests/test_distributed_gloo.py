"""CPU, world_size 2, gloo: the host-side exchange step of the sharded path.

The CUDA library cannot run here, so each rank's engine is replaced by a stub that
returns the partial sums a rank would produce (owned forces, partial E and W); the test
checks the calculator's reduction logic: one all-reduce of 10 doubles for E + virial,
owner-computed forces, optional force gather, stress/energy identical on all ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class StubEngine:
    device = 0
    species = {29}

    def __init__(self, N, seed=0):
        rng = np.random.default_rng(seed)
        self.F = rng.normal(size=(N, 3))
        self.e_atom = rng.normal(size=N)
        self.w_atom = rng.normal(size=(N, 3, 3))

        self.beta = np.abs(rng.normal(size=N))
        self.beta[3] = np.nan        # centre excluded through a / a_not
        self.beta[N - 2] = np.inf    # species without a vscale entry

    def predict(self, pos, numbers, cell, pbc, rank=0, world=1, want_beta=False):
        N = len(numbers)
        c0, c1 = N * rank // world, N * (rank + 1) // world
        owned = np.zeros(N, dtype=bool)
        owned[c0:c1] = True
        F = np.where(owned[:, None], self.F, 0.0)
        out = (float(self.e_atom[owned].sum()), F, self.w_atom[owned].sum(axis=0), owned)
        return out + (np.where(owned, self.beta, 0.0),) if want_beta else out

    def close(self):
        pass


class Atoms:
    def __init__(self, N):
        rng = np.random.default_rng(1)
        self.positions = rng.uniform(0, 10, (N, 3))
        self.cell = np.eye(3) * 10.0
        self.pbc = [True] * 3
        self.numbers = np.full(N, 29)


def _worker(rank, world, port, gather, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import autoforce_b200 as ab

    N = 37
    calc = ab.B200Calculator(ab.SgprModel.from_envs([], lmax=3, nmax=3, xi=4.0, rc=6.0), gather_forces=gather)
    calc._engine = StubEngine(N)
    res = calc.calculate(Atoms(N))
    torch.save({"energy": float(res["energy"]), "forces": res["forces"], "stress": res["stress"], "owned": calc.owned},
               os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("gather", [True, False])
def test_two_rank_reduction(tmp_path, gather):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, gather, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(str(tmp_path), f"r{k}.pt"), weights_only=False) for k in range(2)]
    ref = StubEngine(37)
    assert r[0]["energy"] == r[1]["energy"]
    assert abs(r[0]["energy"] - ref.e_atom.sum()) < 1e-12
    W = ref.w_atom.sum(axis=0)
    stress = (W / 1000.0).reshape(-1)[[0, 4, 8, 5, 2, 1]]
    for k in range(2):
        assert np.abs(r[k]["stress"] - stress).max() < 1e-14
    if gather:
        for k in range(2):
            assert np.abs(r[k]["forces"] - ref.F).max() < 1e-15
    else:
        assert np.abs(r[0]["forces"] + r[1]["forces"] - ref.F).max() < 1e-15
        assert np.all(r[0]["forces"][~r[0]["owned"]] == 0)
    assert np.all(r[0]["owned"] ^ r[1]["owned"])


def _worker_beta(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import autoforce_b200 as ab

    N = 37
    model = ab.SgprModel.from_envs([], lmax=3, nmax=3, xi=4.0, rc=6.0, choli=np.zeros((0, 0)))
    seen = []
    calc = ab.B200Calculator(model, covloss=True, ediff=0.5, on_uncertain=lambda a, b: seen.append(rank))
    calc._engine = StubEngine(N)
    calc.calculate(Atoms(N))
    torch.save({"beta": calc.beta, "covlog": calc.covlog, "seen": seen}, os.path.join(out, f"b{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_covloss_gather(tmp_path):
    """Each rank fills the betas of the atoms it owns; the calculator merges them (NaN and inf survive) and every
    rank logs the same maximum; only rank 0 reports the uncertain structure (calculator/active.py:492-499)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_beta, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(str(tmp_path), f"b{k}.pt"), weights_only=False) for k in range(2)]
    ref = StubEngine(37).beta
    for k in range(2):
        assert np.array_equal(np.isnan(r[k]["beta"]), np.isnan(ref)) and np.array_equal(np.isinf(r[k]["beta"]), np.isinf(ref))
        fin = np.isfinite(ref)
        assert np.abs(r[k]["beta"][fin] - ref[fin]).max() < 1e-15
    assert r[0]["covlog"] == r[1]["covlog"] == "nan"      # max() propagates NaN like torch.max
    assert r[0]["seen"] == [] and r[1]["seen"] == []       # nan > ediff is False


def test_reference_arm_under_torchrun():
    """`bench.py --impl reference` launched the way the driver launches every N > 1 run: under torchrun.  Rank 0 alone
    runs the CPU arm and prints ONE JSON line; its worker processes must not inherit the launcher's rendezvous variables
    (with TORCHELASTIC_USE_AGENT_STORE set they would wait on the agent's store forever)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c2",
           "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(x) for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    line = lines[0]
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
