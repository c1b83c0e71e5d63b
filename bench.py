#!/usr/bin/env python
"""bench.py -- atom-steps/sec of SGPR E+F+stress prediction (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

A step = one full evaluation (neighbour list -> descriptors -> kernel GEMMs -> energy,
forces, virial) of one synthetic structure.  Workload (default c3, the configuration the
metric's target is quoted on): 4-species Li/P/S/O-like 97,556-atom fcc-derived solid,
random-init frozen model with M = 2000 inducing LCEs, SeSoap lmax=3 nmax=3 rc=6.
Prints ONE JSON line (see DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "atom-steps/sec (E+F+stress) SGPR MD prediction"
UNIT = "atom-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="atoms in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variants", type=int, default=4, help="pre-generated perturbed position sets cycled over the steps")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "halo"],
                    help="multi-GPU force exchange: peer-memory adds over NVLink (default) or halo recompute")
    return ap.parse_args()


def workload_inputs(name, variants):
    from autoforce_b200 import synth

    w = synth.WORKLOADS[name]
    pos, cell, numbers = synth.fcc(w["rep"], w["Zs"], 0.1, 0)
    rng = np.random.default_rng(123)
    # every step sees perturbed positions, N(0, 0.03 A); generated before timing
    pos_variants = [pos + rng.normal(0, 0.03, pos.shape) for _ in range(max(1, variants))]
    return w, pos_variants, cell, numbers


def config_of(name, w, N, M, world, exchange="none"):
    return {
        "workload": f"{name}: fcc-derived {N}-atom {'/'.join(str(z) for z in w['Zs'])} solid, frozen random-init SGPR model "
                    f"M={M}, SeSoap lmax={w['lmax']} nmax={w['nmax']} rc={w['rc']}, xi=4",
        "atoms": N, "inducing": M, "species": len(w["Zs"]),
        "parallelism": f"atoms sharded over {world} GPU(s); positions replicated; all-reduce of E + 3x3 virial only"
                       + ("" if world == 1 else f"; forces: {exchange}"),
        "cache": "working set (pair list + descriptor/gradient matrices, >500 MB at c3) exceeds the 126 MB L2; positions change every step",
    }


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dgemm_peak_tflops():
    """FP64 GEMM peak of this GPU measured in-run (cuBLAS DGEMM 8192^3, best of 5):
    MEASURED_PEAKS.json only lists bf16, and the kernel GEMMs run on the FP64 pipe."""
    import torch

    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n**3 / (best * 1e-3) / 1e12


def cpu_arm(args, world, rank):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the
    Python reference itself cannot travel to the GPU box)."""
    if rank != 0:
        return
    from autoforce_b200 import synth
    from oracle.cpu_bench import CpuBench

    w, pos_variants, cell, numbers = workload_inputs(args.workload, args.variants)
    from oracle.sgpr_oracle import neighbor_list

    model = synth.synth_model(w["Zs"], w["M"], 1, lmax=w["lmax"], nmax=w["nmax"], rc=w["rc"], neighbors_fn=neighbor_list)
    procs = min(os.cpu_count() or 1, 64)
    sample = args.cpu_sample or 32 * procs
    cb = CpuBench(model, pos_variants, cell, numbers, sample, procs)
    value, sec = cb.run(args.steps, args.warmup)
    cb.close()
    N = len(numbers)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(args.workload, w, N, model.M, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{cb.sample} of {N} atoms per step (full neighbour environments, all {model.M} inducing LCEs), "
                                   f"{procs} worker processes; oracle/sgpr_oracle.py (vectorised numpy restatement of the reference)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        cpu_arm(args, world, rank)
        return
    import torch
    import torch.distributed as dist

    import autoforce_b200 as ab
    from autoforce_b200 import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w, pos_variants, cell, numbers = workload_inputs(args.workload, args.variants)
    N = len(numbers)
    model = synth.synth_model(w["Zs"], w["M"], 1, lmax=w["lmax"], nmax=w["nmax"], rc=w["rc"])
    eng = ab.SgprEngine(model, species=w["Zs"], device=local_rank)
    pbc = True
    pos_d = [torch.as_tensor(p, device=dev) for p in pos_variants]
    z_d = torch.as_tensor(numbers.astype(np.int32), device=dev)
    out = (torch.empty(1, dtype=torch.float64, device=dev), torch.empty((N, 3), dtype=torch.float64, device=dev),
           torch.empty(9, dtype=torch.float64, device=dev))
    ew = torch.zeros(10, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    exchange = "none"
    px = None
    if world > 1:
        exchange = args.exchange
        if exchange == "p2p":
            try:
                px = eng.peer_exchange(N)
            except Exception as ex:  # symmetric memory unavailable -> halo recompute
                if rank == 0:
                    print(f"bench: peer-memory exchange unavailable ({ex}); falling back to halo recompute", file=sys.stderr)
                exchange = "halo"
        ok = torch.tensor([1 if exchange == "p2p" else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            exchange, px = "halo", None
    # host buffers of the end-to-end leg are page-locked (inputs and the force output)
    pin_variants = [torch.from_numpy(p).pin_memory() for p in pos_variants]
    pin_F = torch.empty((N, 3), dtype=torch.float64).pin_memory()
    pin_F_np = pin_F.numpy()
    pin_Z = torch.from_numpy(np.ascontiguousarray(numbers, dtype=np.int32)).pin_memory()
    dev_pos = torch.empty((N, 3), dtype=torch.float64, device=dev) if px is not None else None

    def step_device(it):
        if px is not None:
            px.step(pos_d[it % len(pos_d)], z_d, cell, pbc)
            return
        E, F, W = eng.predict_device(pos_d[it % len(pos_d)], z_d, cell, pbc, rank=rank, world=world, out=out)
        if world > 1:  # the only collective of the path: 10 doubles
            ew[0:1].copy_(E)
            ew[1:].copy_(W)
            dist.all_reduce(ew)

    def step_host(it):
        if px is not None:   # host buffers in and out around the peer-memory step
            dev_pos.copy_(pin_variants[it % len(pin_variants)], non_blocking=True)
            E, F, W, owned = px.step(dev_pos, z_d, cell, pbc)
            pin_F.copy_(F, non_blocking=True)
            Eh = float(E.item())   # D2H of the reduced energy: synchronises the step
            torch.cuda.current_stream().synchronize()
            return Eh
        E, F, W, owned = eng.predict(pin_variants[it % len(pin_variants)].numpy(), pin_Z.numpy(), cell, pbc, rank=rank,
                                     world=world, out_forces=pin_F_np)
        if world > 1:
            t = torch.tensor([E] + list(W.reshape(-1)), dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            E = float(t[0].item())
        return E

    def timed(fn, steps, warmup):
        for it in range(warmup):
            fn(it)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for it in range(steps):
            fn(it)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    # ---- device-resident leg (value) with per-stage CUDA-event timing inside the library
    eng.enable_timing(True)
    sampler = ClockSampler(local_rank)
    stage = {"ms_nl": 0.0, "ms_desc": 0.0, "ms_gemm": 0.0, "ms_force": 0.0, "ms_total": 0.0, "gemm_flops": 0.0, "i8_ops": 0.0,
             "launches": 0}

    def step_device_acc(it):
        step_device(it)
        s = eng.stats()
        if it >= 0:
            for k in ("ms_nl", "ms_desc", "ms_gemm", "ms_force", "ms_total"):
                stage[k] += s[k]
            stage["gemm_flops"] += s["gemm_flops"]
            stage["i8_ops"] += s["i8_ops"]
            stage["launches"] += s["kernel_launches"]
            stage["n_active"] = s["n_active"]
            stage["n_pairs"] = s["n_pairs"]

    for it in range(args.warmup):
        step_device(it)
    for k in list(stage):
        stage[k] = 0 if k == "launches" else 0.0
    sampler.start()
    ms_dev, wall_dev = timed(step_device_acc, args.steps, 0)
    clocks = sampler.stop()
    eng.enable_timing(False)
    # ---- end-to-end leg: host buffers through the public host API (H2D + D2H inside)
    ms_e2e, wall_e2e = timed(step_host, args.steps, args.warmup)

    if rank == 0:
        K = args.steps
        value = N * K / (ms_dev * 1e-3)
        e2e = N * K / (wall_e2e * 1e-3)
        gemm_s = stage["ms_gemm"] * 1e-3
        fp64_equiv = stage["gemm_flops"] / gemm_s / 1e12 if gemm_s > 0 else None
        try:
            dgemm_peak = dgemm_peak_tflops()
        except Exception:  # pragma: no cover
            dgemm_peak = None
        use_i8 = stage["i8_ops"] > 0
        if use_i8:
            # the GEMMs run on tcgen05 as int8 digit-slice products: roofline in int8 tensor operations
            achieved = stage["i8_ops"] / gemm_s / 1e12
            peak, peak_src = 4500.0, "fallback: nominal dense int8 tensor peak of B200 (MEASURED_PEAKS.json lists bf16 only)"
            try:
                mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                peak = 2.0 * float(mp["bf16_tflops"])
                peak_src = "2 x measured dense bf16 peak of MEASURED_PEAKS.json (int8 tensor rate = 2 x bf16; burst figure)"
            except Exception:
                pass
            kernel_name = "i8gemm_kernel (tcgen05.mma kind::i8, TMA + TMEM): FP64-accurate kernel GEMM + back projection as int8 digit-slice products"
            unit = "TOP/s (int8)"
        else:
            achieved = fp64_equiv
            peak = dgemm_peak if dgemm_peak else 40.0
            peak_src = "measured in-run: cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json has no FP64 entry)"
            kernel_name = "gemm_tn_kernel (FP64 DMMA kernel-matrix GEMM + back projection)"
            unit = "TFLOP/s"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload, {}).get("i8gemm_dram_bytes_per_launch" if use_i8 else "gemm_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(args.workload, w, N, model.M, world, {"p2p": "peer-memory adds over NVLink into the owner's buffer (no halo recompute)", "halo": "owner-computes with one-cutoff halo recompute", "none": ""}[exchange]),
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": wall_e2e / K,
                    "h2d_bytes_per_step": int(N * 24 + N * 4), "d2h_bytes_per_step": int(N * 24 + 16 * 8 + N)},
            "gpu_launches": int(stage["launches"]),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": kernel_name,
                         "achieved": achieved, "peak": peak, "unit": unit, "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "peak_source": peak_src,
                         "fp64_equivalent_tflops": fp64_equiv, "cublas_dgemm_tflops_in_run": dgemm_peak,
                         "int8_ops_per_step": stage["i8_ops"] / K,
                         "flops_per_step": stage["gemm_flops"] / K, "gemm_ms_per_step": stage["ms_gemm"] / K},
            "stages_ms_per_step": {k[3:]: stage[k] / K for k in ("ms_nl", "ms_desc", "ms_gemm", "ms_force", "ms_total")},
            "pairs": int(stage.get("n_pairs", 0)), "active_envs_rank0": int(stage.get("n_active", 0)),
        }
        if world == 1:
            # covloss (calculator/active.py:781-804) runs every prediction step in the reference but is not part
            # of the metric: reported separately, same structure, model with choli = 0.5 I
            try:
                model_c = synth.synth_model(w["Zs"], w["M"], 1, lmax=w["lmax"], nmax=w["nmax"], rc=w["rc"], with_choli="tril")
                eng_c = ab.SgprEngine(model_c, species=w["Zs"], device=local_rank)
                eng_c.enable_timing(True)
                tb, fl = [], 0.0
                for it in range(4):
                    eng_c.predict(pos_variants[it % len(pos_variants)], numbers, cell, pbc, want_beta=True)
                    st_c = eng_c.stats()
                    if it > 0:
                        tb.append(st_c["ms_beta"])
                        fl = st_c["covloss_flops"]
                eng_c.close()
                mb = sum(tb) / len(tb)
                line["covloss"] = {"ms_per_step": mb, "tflops": fl / (mb * 1e-3) / 1e12, "flops_per_step": fl,
                                   "note": "extra device time per step when beta is requested; tcgen05 int8 digit-slice GEMM K.choli^T (26 slice products; dense lower-triangular synthetic choli like the reference's L^-1, whose zero K-chunks are skipped; flops counted for the dense product) + row sum of squares"}
            except Exception as ex:  # pragma: no cover
                line["covloss"] = {"error": str(ex)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle.cpu_bench import CpuBench

                procs = min(os.cpu_count() or 1, 64)
                sample = args.cpu_sample or 16 * procs
                cb = CpuBench(model, pos_variants, cell, numbers, sample, procs)
                v, sec = cb.run(2, 1)
                cb.close()
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
                                        "sample": f"{cb.sample} of {N} atoms x 2 steps (full environments, all {model.M} inducing LCEs), "
                                                  f"{procs} processes, oracle/sgpr_oracle.py"}
            except Exception as ex:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
