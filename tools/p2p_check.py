"""Developer/CI tool (>= 2 GPUs, torchrun): the peer-memory force exchange against the unsharded result.
   torchrun --nproc-per-node 2 tools/p2p_check.py [workload] [p2p|nccl] [steps]
   default: the fused exchange step (sgpr_p2p_step: mailboxes + stamped flags, no NCCL in the step);
   "nccl": sgpr_predict_p2p + NCCL all-reduce + sgpr_p2p_collect."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import autoforce_b200 as ab
from autoforce_b200 import synth

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = synth.WORKLOADS[wl]
model = synth.synth_model(w["Zs"], w["M"], 1, lmax=w["lmax"], nmax=w["nmax"], rc=w["rc"])
pos, cell, numbers = synth.fcc(w["rep"], w["Zs"], 0.1, 0)
N = len(pos)
fused = not (len(sys.argv) > 2 and sys.argv[2] == "nccl")
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
torch.cuda.set_stream(torch.cuda.Stream(device=dev))             # warm steps replay as CUDA graphs (not on stream 0)
ref = ab.SgprEngine(model, species=w["Zs"], device=local)        # unsharded reference on this GPU
eng = ab.SgprEngine(model, species=w["Zs"], device=local)
eng.set_async(True)
px = eng.peer_exchange(N, fused=fused)
z_t = torch.as_tensor(numbers.astype(np.int32), device=dev)
ok = True
rng = np.random.default_rng(5)
variants = [pos, pos + rng.normal(0, 0.02, pos.shape), pos + rng.normal(0, 0.03, pos.shape)]
variants_d = [torch.as_tensor(p, device=dev) for p in variants]
refs = [ref.predict(p, numbers, cell, True) for p in variants]
for it in range(n_steps):                                        # sizing, warm, graph replay, both buffer parities
    v = (it // 2) % len(variants) if it < 8 else int(rng.integers(0, len(variants)))
    if it >= 8:
        # soak: a burst of steps without any host synchronisation (ranks drift apart by up to a step; the stamped
        # mailboxes and the two buffer parities must keep them apart); the same random sequence on every rank
        for _ in range(it % 7):
            px.step(variants_d[int(rng.integers(0, len(variants)))], z_t, cell, True)
    Er, Fr, Wr, _ = refs[v]
    E, F, W, owned = px.step(variants_d[v], z_t, cell, True)
    torch.cuda.synchronize()
    eng.check()
    owned = owned.cpu().numpy().astype(bool)
    F = F.cpu().numpy()
    dE = abs(float(E) - Er) / N
    dW = np.abs(W.cpu().numpy() - Wr).max()
    dF = np.abs(F[owned] - Fr[owned]).max()
    cnt = torch.tensor([int(owned.sum())], device=dev)
    dist.all_reduce(cnt)
    good = dE < 1e-11 and dW < 1e-8 and dF < 1e-10 and int(cnt.item()) == N and np.all(F[~owned] == 0)
    ok &= good
    if it < 8 or not good or it == n_steps - 1:
        print(f"rank {rank} step {it}: dE/N={dE:.2e} dW={dW:.2e} dF={dF:.2e} owned={int(owned.sum())} total_owned={int(cnt.item())} {'OK' if good else 'FAIL'}", flush=True)
eng.close()
ref.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
