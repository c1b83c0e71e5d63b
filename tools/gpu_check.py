"""Developer diagnostic (GPU box): run every golden case through the CUDA path and
print per-stage deviations from the oracle / golden vectors without asserting."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from golden_util import golden_cases, golden_radii, load_golden, oracle_model  # noqa: E402
from oracle import sgpr_oracle as o  # noqa: E402

import autoforce_b200 as ab  # noqa: E402


def model_from_golden(g, big=False):
    from golden_util import b200_model

    return b200_model(g, big)


def sorted_rows(first, J, S):
    I = np.repeat(np.arange(len(first) - 1), np.diff(first))
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], J, I))
    return I[order], J[order].astype(np.int64), S[order].astype(np.int64)


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    for name in sys.argv[1:] or golden_cases():
        try:
            g = load_golden(name)
            m = model_from_golden(g)
            species = g["meta"]["species"]
            eng = ab.SgprEngine(m, species=species)
            pbc = g["meta"]["pbc"]
            N = len(g["numbers"])
            first, J, S = eng.neighbors(g["pos"], g["numbers"], g["cell"], pbc)
            I1, J1, S1 = sorted_rows(first, J, S)
            I0, J0, S0 = sorted_rows(g["nl_first"], g["nl_j"], g["nl_S"])
            nl_ok = len(J1) == len(J0) and np.array_equal(I1, I0) and np.array_equal(J1, J0) and np.array_equal(S1, S0)
            Zh = eng.inducing_descriptors()
            om = oracle_model(g)
            Zo, lone = o.inducing_descriptors(om, np.array(species))
            dZ = np.abs(Zh - Zo).max()
            P = eng.descriptors(g["pos"], g["numbers"], g["cell"], pbc)
            ref = o.predict(om, g["pos"], g["cell"], pbc, g["numbers"], want_K=True)
            dP = 0.0
            for key in [k for k in g if k.startswith("desc_")]:
                a = int(key.split("_")[1])
                dP = max(dP, np.abs(P[a] - g[key].reshape(P[a].shape)).max())
            K = eng.kernel_matrix(g["pos"], g["numbers"], g["cell"], pbc).cpu().numpy()
            dK = np.abs(K - g["K"]).max()
            E, F, W, owned = eng.predict(g["pos"], g["numbers"], g["cell"], pbc)
            vol = abs(np.linalg.det(g["cell"])) or -2.0
            stress = (W / vol).reshape(-1)[[0, 4, 8, 5, 2, 1]]
            print(f"{name:16s} N={N:4d} nl_ok={nl_ok} pairs={len(J)}/{len(g['nl_j'])} dZ={dZ:.2e} dP={dP:.2e} dK={dK:.2e} "
                  f"dE/N={abs(E - g['energy']) / N:.2e} dF={np.abs(F - g['forces']).max():.2e} dS={np.abs(stress - g['stress']).max():.2e} "
                  f"launches={eng.stats()['kernel_launches']}", flush=True)
            eng.set_weights(mu=g["mu_big"])
            E, F, W, owned = eng.predict(g["pos"], g["numbers"], g["cell"], pbc)
            stress = (W / vol).reshape(-1)[[0, 4, 8, 5, 2, 1]]
            print(f"{'':16s} big-mu: dE/N={abs(E - g['energy_big']) / N:.2e} dF={np.abs(F - g['forces_big']).max():.2e} "
                  f"(|F|max {np.abs(g['forces_big']).max():.1f}) dS={np.abs(stress - g['stress_big']).max():.2e}", flush=True)
            eng.close()
        except Exception:
            print(f"{name}: EXCEPTION", flush=True)
            traceback.print_exc()


def box_speed():
    """cuBLAS DGEMM 4096^3 as a box-speed indicator (boxes differ by several percent)."""
    a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    torch.matmul(a, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        torch.matmul(a, a)
    e1.record()
    torch.cuda.synchronize()
    print(f"box speed: cuBLAS DGEMM 4096^3 {5 * 2 * 4096**3 / e0.elapsed_time(e1) / 1e9:.2f} TFLOP/s", flush=True)


def speed(workload="c2", steps=5):
    from autoforce_b200 import synth

    w = synth.WORKLOADS[workload]
    t0 = time.time()
    model = synth.synth_model(w["Zs"], w["M"], 1, lmax=w["lmax"], nmax=w["nmax"], rc=w["rc"])
    pos, cell, numbers = synth.fcc(w["rep"], w["Zs"], 0.1, 0)
    eng = ab.SgprEngine(model, species=w["Zs"])
    eng.enable_timing(True)
    print(f"{workload}: N={len(pos)} M={model.M} setup {time.time() - t0:.1f}s", flush=True)
    for it in range(steps):
        t = time.time()
        E, F, W, owned = eng.predict(pos, numbers, cell, True)[:4]
        dt = time.time() - t
        s = eng.stats()
        print(f"  step {it}: {dt * 1e3:.2f} ms wall  E={E:.6f} |F|max={np.abs(F).max():.4f} pairs={s['n_pairs']} "
              f"nl={s['ms_nl']:.3f} desc={s['ms_desc']:.3f} gemm={s['ms_gemm']:.3f} force={s['ms_force']:.3f} total={s['ms_total']:.3f} ms "
              f"gemm TF/s={s['gemm_flops'] / max(s['ms_gemm'], 1e-9) / 1e9:.2f}", flush=True)
    if os.environ.get("BETA"):
        model2 = synth.synth_model(w["Zs"], w["M"], 1, lmax=w["lmax"], nmax=w["nmax"], rc=w["rc"], with_choli=os.environ.get("CHOLI", "tril") if os.environ.get("CHOLI", "tril") != "eye" else True)
        eng2 = ab.SgprEngine(model2, species=w["Zs"])
        eng2.enable_timing(True)
        for it in range(3):
            t = time.time()
            out = eng2.predict(pos, numbers, cell, True, want_beta=True)
            s = eng2.stats()
            print(f"  beta step {it}: {1e3 * (time.time() - t):.2f} ms wall, beta stage {s['ms_beta']:.3f} ms, "
                  f"{s['covloss_flops'] / max(s['ms_beta'], 1e-9) / 1e9:.2f} TF/s, beta max {np.nanmax(out[4]):.4f}", flush=True)
        eng2.close()
    eng.close()


if __name__ == "__main__":
    main()
    box_speed()
    for wl in os.environ.get("SPEED", "c2,c3").split(","):
        if wl:
            try:
                speed(wl)
            except Exception:
                traceback.print_exc()
