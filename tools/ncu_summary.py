#!/usr/bin/env python
"""Turn the two ncu captures of bench.py (B200_PROFILING.md recipe) into the tracked evidence under profiles/:

    tools/ncu_summary.py <round tag> <launches.csv> <full.ncu-rep> [workload]

* <launches.csv>: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` of a short bench run -> the per-kernel
  launch list of one steady-state step (profiles/<tag>_launches_<wl>.csv holds the whole capture);
* <full.ncu-rep>: `ncu --set full` of the dominant kernels -> DRAM bytes, L2 -> SM bytes, pipe utilisation, issue-slot
  utilisation, occupancy per kernel (read here with `ncu -i ... --page raw --csv`).
Writes profiles/<tag>_ncu_summary_<wl>.md and the per-stage entries of profiles/roofline_traffic.json that bench.py
attaches to its `roofline` / `roofline_stages` objects.
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HBM_PEAK = 6536.0   # GB/s, MEASURED_PEAKS.json


def launch_list(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[start + 2:] if len(r) > vi]
    idx = [i for i, (n, _) in enumerate(seq) if "bin_count_kernel" in n]
    # a steady-state step of the device-resident leg: the 4th step of the capture
    a, b = idx[3], idx[4]
    return seq[a:b]


def full_metrics(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    want = {
        "gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
        "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm", "launch__registers_per_thread": "regs",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct", "smsp__inst_executed.sum": "warp_inst",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    }
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    res = {}
    for r in rows[2:]:
        name = r[ki]
        d = {}
        for i, h in enumerate(hdr):
            if h in want:
                v = float(r[i].replace(",", ""))
                if units[i] in scale:
                    v *= scale[units[i]]
                if units[i] in ("ms", "msecond"):
                    v *= 1e3
                if units[i] in ("ns", "nsecond"):
                    v *= 1e-3
                d[want[h]] = v
        res.setdefault(name, d)   # first instance of each kernel
    return res


def short(name):
    n = name.replace("sgpr::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
    return n[: n.index("(")] if "(" in n else n


def main():
    tag, launches, full = sys.argv[1:4]
    wl = sys.argv[4] if len(sys.argv) > 4 else "c3"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    shutil.copyfile(launches, os.path.join(ROOT, "profiles", f"{tag}_launches_{wl}.csv"))
    step = launch_list(launches)
    total = sum(v for _, v in step)
    fm = full_metrics(full)
    lines = [f"# {tag} — ncu evidence, workload {wl} (1 x B200)", "",
             "Commands (the recipe of /opt/skills/guides/B200_PROFILING.md, run through `gpurun`):", "",
             "```",
             "ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline",
             'SGPR_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"desc_forward|desc_backward|i8gemm|neighbor_bin" -s 12 -c 6 -o full python bench.py --steps 2 --warmup 3 --no-cpu-baseline',
             "```", "",
             f"## Launch list of one steady-state step ({len(step)} launches, {total / 1e6:.3f} ms serialised under ncu; "
             f"`profiles/{tag}_launches_{wl}.csv` is the whole capture)", "",
             "ncu serialises launches and runs them cold-cache, so the absolute times are upper bounds; the SHARES are what must "
             "agree with the stage split bench.py measures with CUDA events.", "",
             "| kernel | us | share |", "|---|---|---|"]
    for n, v in step:
        if v >= 2000:
            lines.append(f"| `{short(n)}` | {v / 1e3:.1f} | {100 * v / total:.1f} % |")
    small = sum(v for _, v in step if v < 2000)
    lines.append(f"| {sum(1 for _, v in step if v < 2000)} launches < 2 us each | {small / 1e3:.1f} | {100 * small / total:.1f} % |")
    lines += ["", "## `--set full` per kernel (one launch each)", "",
              "| kernel | us | DRAM read + write (MB) | DRAM GB/s (frac of 6536) | L2 -> SM (MB) | L2 -> SM TB/s | FP64 pipe % | tensor pipe % | "
              "issue slots % | warps active % | regs |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    stage_of = {"desc_forward": "stage_desc", "desc_backward": "stage_force", "neighbor_bin_kernel<0": "stage_nl_count",
                "neighbor_bin_kernel<1": "stage_nl_fill", "Epi1": "stage_gemm_kernel_matrix", "Epi2": "stage_gemm_back_projection"}
    prof = {}
    for name, d in fm.items():
        t = d["time_us"]
        dram = d["dram_rd"] + d["dram_wr"]
        gbs = dram / (t * 1e-6) / 1e9
        lines.append(f"| `{short(name)}` | {t:.1f} | {dram / 1e6:.0f} | {gbs:.0f} ({gbs / HBM_PEAK:.2f}) | {d['l2_to_sm'] / 1e6:.0f} | "
                     f"{d['l2_to_sm'] / (t * 1e-6) / 1e12:.2f} | {d['fp64_pipe_pct']:.1f} | {d['tensor_pipe_pct']:.1f} | "
                     f"{d['issue_active_pct']:.1f} | {d['warps_active_pct']:.1f} | {int(d['regs'])} |")
        for key, st in stage_of.items():
            if key in name:
                prof[st] = {"kernel": short(name), "us": t, "dram_bytes": dram, "dram_gbs": gbs, "frac_hbm": gbs / HBM_PEAK,
                            "l2_to_sm_bytes": d["l2_to_sm"], "fp64_pipe_pct": d["fp64_pipe_pct"], "tensor_pipe_pct": d["tensor_pipe_pct"],
                            "issue_active_pct": d["issue_active_pct"], "warps_active_pct": d["warps_active_pct"], "registers": int(d["regs"]),
                            "source": f"profiles/{tag}_ncu_summary_{wl}.md"}
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary_{wl}.md"), "w").write("\n".join(lines) + "\n")
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    allp = json.load(open(tpath)) if os.path.exists(tpath) else {}
    w = allp.setdefault(wl, {})
    if "stage_gemm_kernel_matrix" in prof:
        w["i8gemm_dram_bytes_per_launch"] = prof["stage_gemm_kernel_matrix"]["dram_bytes"]
    w["stage_desc"] = prof.get("stage_desc")
    w["stage_force"] = prof.get("stage_force")
    w["stage_nl"] = {"count": prof.get("stage_nl_count"), "fill": prof.get("stage_nl_fill")}
    w["stage_gemm"] = {"kernel_matrix": prof.get("stage_gemm_kernel_matrix"), "back_projection": prof.get("stage_gemm_back_projection")}
    w["captured"] = tag
    json.dump(allp, open(tpath, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
