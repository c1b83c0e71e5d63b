"""The reference's own CPU path as bench.py's baseline (TEST / BENCH INFRASTRUCTURE ONLY, oracle/__init__.py).

What runs: the UNMODIFIED ``theforce.calculator.active.ActiveCalculator.calculate`` (calculator/active.py:425-611) in
prediction mode (``calculator=None``) -- neighbour list, per-atom descriptors, the N x M kernel loop, autograd forces and
stress, covloss -- imported from /root/reference or from the copy staged under oracle/_ref, on P worker processes
with ``torch.set_num_threads(1)`` and OMP/MKL/OPENBLAS_NUM_THREADS=1 (set in the environment BEFORE the interpreter
starts).  The P processes use the reference's own decomposition over atoms (``Distributer`` / ``index_distribute``,
descriptor/atoms.py:228-259,321-341) and its own all-reduces of energy, forces and cell gradient
(calculator/active.py:562,601-602); MPI is not installed, so ``theforce.distributed`` falls back to mpi4py
(distributed.py:4-10) and oracle/shims/mpi4py carries the collectives over torch.distributed/gloo.

A full-size step is infeasible on a CPU (c3: ~4.6 core-hours), so the bounded sample is a SMALLER
periodic cell of the same workload family -- same lattice, species mix, rattle, per-step perturbation, kernel and the
same M inducing LCEs -- sized to ~``atoms_per_proc`` atoms per process and evaluated IN FULL by the reference.  The
reference's cost is linear in the number of atoms at fixed M (3.8 ms N + 0.34 ms N M_sameZ per process, BASELINE.md
section 2), so atom-steps/s of the sample is the rate at the full size.  Because the sample is a complete structure its
energy / forces / stress / covloss are real reference results: bench.py compares the GPU path against them.

ASE is not in the image: the neighbour list comes from oracle/shims/ase (cell-list restatement of ASE's semantics).
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def sample_rep(procs, atoms_per_proc=16, min_rep=3):
    """fcc repetitions of the sample cell: 4 rep^3 atoms ~ procs * atoms_per_proc."""
    return max(min_rep, int(round((procs * atoms_per_proc / 4.0) ** (1.0 / 3.0))))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def thread_env(env=None):
    env = dict(os.environ if env is None else env)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        env[k] = "1"
    return env


def run(workload, steps, warmup, procs=None, rep=None, variants=4, keep_dir=None, timeout=3000):
    """Launch the reference on ``procs`` processes; returns the dict rank 0 wrote (value in atom-steps/s, timings,
    the last step's results) plus the structure / model files for a parity check."""
    sys.path.insert(0, ROOT)
    from autoforce_b200 import synth
    from oracle.sgpr_oracle import neighbor_list

    procs = procs or min(os.cpu_count() or 1, 64)
    w = synth.WORKLOADS[workload]
    rep = rep or sample_rep(procs)
    work = keep_dir or tempfile.mkdtemp(prefix="sgpr_ref_bench_")
    os.makedirs(work, exist_ok=True)
    # the same frozen model as the GPU arm's (autoforce_b200/synth.py), + choli = 0.5 I: the reference evaluates
    # covloss on every prediction step (calculator/active.py:492-499) and needs it
    model = synth.synth_model(w["Zs"], w["M"], 1, lmax=w["lmax"], nmax=w["nmax"], rc=w["rc"], with_choli=True,
                              neighbors_fn=neighbor_list)
    model.save(os.path.join(work, "model.npz"))
    pos, cell, numbers = synth.fcc(rep, w["Zs"], 0.1, 0)
    rng = np.random.default_rng(123)
    pos_variants = np.stack([pos + rng.normal(0, 0.03, pos.shape) for _ in range(max(1, variants))])
    np.savez(os.path.join(work, "structure.npz"), pos_variants=pos_variants, cell=cell, numbers=numbers)
    port = _free_port()
    env = thread_env()
    env.update(SGPR_SHIM_WORLD=str(procs), SGPR_SHIM_PORT=str(port), PYTHONPATH=ROOT + os.pathsep + env.get("PYTHONPATH", ""))
    # under torchrun the parent carries the launcher's rendezvous variables; with TORCHELASTIC_USE_AGENT_STORE set,
    # init_process_group("tcp://...") in the workers would connect to the agent's store instead of hosting their own
    # and wait forever
    for k in list(env):
        if k.startswith(("TORCHELASTIC_", "TORCH_NCCL_", "NCCL_")) or k in (
                "RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE", "ROLE_RANK",
                "ROLE_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k)
    env["CUDA_VISIBLE_DEVICES"] = ""          # the reference arm must not touch a GPU
    ps = []
    for r in range(procs):
        e = dict(env, SGPR_SHIM_RANK=str(r))
        log = open(os.path.join(work, f"worker{r}.log"), "w")
        ps.append((subprocess.Popen([sys.executable, "-m", "oracle.ref_bench", "--worker", "--work", work, "--steps", str(steps),
                                     "--warmup", str(warmup)], env=e, cwd=work, stdout=log, stderr=subprocess.STDOUT), log))
    t0 = time.time()
    rc = 0
    for p, log in ps:
        try:
            rc |= p.wait(timeout=max(1.0, timeout - (time.time() - t0)))
        except subprocess.TimeoutExpired:
            p.kill()
            rc |= 1
        log.close()
    if not os.path.isfile(os.path.join(work, "result.json")):
        tails = "\n".join(f"--- worker{r}.log\n" + open(os.path.join(work, f"worker{r}.log")).read()[-600:] for r in range(min(procs, 4)))
        raise RuntimeError(f"reference workers failed (rc={rc}):\n{tails}")
    # (rank 0 writes result.json atomically after the timed region and all collectives; a worker that trips over a
    # closed socket while shutting down does not invalidate the measurement)
    out = json.load(open(os.path.join(work, "result.json")))
    out["worker_exit_codes_ok"] = rc == 0
    out.update(work=work, procs=procs, rep=rep, M=int(model.M), workload=workload)
    return out


def _worker(args):
    import torch

    torch.set_num_threads(1)
    sys.path.insert(0, ROOT)
    from oracle import ref_runner

    ref_runner.import_reference()
    from autoforce_b200.model import SgprModel   # the flat container only (no CUDA anywhere in this process)
    from theforce.calculator.active import ActiveCalculator
    from theforce.util.parallel import mpi_init

    import theforce.distributed as distrib

    group = mpi_init(seed=12345)
    rank, world = distrib.get_rank(), distrib.get_world_size()
    flat = SgprModel.load(os.path.join(args.work, "model.npz"))
    z = np.load(os.path.join(args.work, "structure.npz"))
    pos_variants, cell, numbers = z["pos_variants"], z["cell"], z["numbers"]
    kernel = ref_runner.make_kernel("sesoap", flat.lmax, flat.nmax, int(flat.xi) if float(flat.xi).is_integer() else flat.xi, flat.rc)
    first = flat.ind_first
    envs = [(int(flat.ind_Z[m]), flat.ind_r[first[m]:first[m + 1]], flat.ind_b[first[m]:first[m + 1]]) for m in range(flat.M)]
    model = ref_runner.synth_model(kernel, envs, flat.mu, flat.mean_w, flat.choli, flat.vscale)
    logfile = os.path.join(args.work, "active.log") if rank == 0 else None
    calc = ActiveCalculator(covariance=model, calculator=None, process_group=group, pckl=None, tape=None, logfile=logfile,
                            report_timings=True)
    atoms = ref_runner.ase_atoms(pos_variants[0], cell, True, numbers)
    atoms.calc = calc

    def step(k):
        atoms.set_positions(pos_variants[k % len(pos_variants)])
        e = atoms.get_potential_energy()
        return e, atoms.get_forces(), atoms.get_stress()

    for k in range(args.warmup):
        step(k)
    distrib.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e, f, s = step(args.warmup + k)   # never the positions of the previous call: ase caches an unchanged structure
    distrib.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt])
    distrib.all_reduce(t, distrib.ReduceOp.MAX)
    beta = calc.get_covloss().detach().numpy()     # a collective (gather, active.py:769-779): every rank calls it
    if rank == 0:
        nodes = []
        for line in open(logfile):
            if " timings:" in line:
                nodes.append([float(v) for v in line.split("timings:")[1].split("total:")[0].split()])
        nodes = np.array(nodes[-args.steps:]) if nodes else np.zeros((1, 5))
        N = len(numbers)
        k_last = (args.warmup + args.steps - 1) % len(pos_variants)
        np.savez(os.path.join(args.work, "last_step.npz"), pos=pos_variants[k_last], cell=cell, numbers=numbers, energy=np.array(e),
                 forces=np.array(f), stress=np.array(s), covloss=beta)
        res = dict(value=N * args.steps / float(t[0]), ms_per_step=float(t[0]) / args.steps * 1e3, atoms=int(N), world=int(world),
                   steps=args.steps, warmup=args.warmup, torch_threads=torch.get_num_threads(),
                   omp=os.environ.get("OMP_NUM_THREADS"), reference=ref_runner.reference_kind(),
                   node_seconds=dict(zip(["nl_desc", "kernel", "results", "covloss", "post"], nodes.mean(axis=0).tolist())))
        with open(os.path.join(args.work, "result.json.tmp"), "w") as fh:
            json.dump(res, fh)
        os.replace(os.path.join(args.work, "result.json.tmp"), os.path.join(args.work, "result.json"))
    # orderly shutdown of the gloo world behind the mpi4py stand-in: nobody leaves while a peer still talks, and the
    # interpreter exits without running the process-group destructors (they abort if a peer's socket is already gone)
    distrib.barrier()
    try:
        import torch.distributed as tdist

        if tdist.is_initialized():
            tdist.destroy_process_group()
    except Exception:
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--work", default=None)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--rep", type=int, default=0)
    ap.add_argument("--variants", type=int, default=4)
    a = ap.parse_args()
    if a.worker:
        _worker(a)
    else:
        print(json.dumps(run(a.workload, a.steps, a.warmup, a.procs or None, rep=a.rep or None, variants=a.variants, keep_dir=a.work)))
