#!/usr/bin/env python
"""Turn the raw ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/ncu_summary.py --tag r01 [--workload G256_K16384_T50]

Reads  gpurun_out/launches.csv        (ncu --metrics gpu__time_duration.sum ... python bench.py ...)
       gpurun_out/prof_rollout.ncu-rep (ncu --set full -k regex:rollout_kernel ...)
Writes profiles/<tag>_launches_summary.txt, profiles/<tag>_rollout_ncu_summary.txt and updates
       profiles/rollout_traffic.json (per-launch DRAM bytes of the rollout kernel, read by bench.py).
"""
import argparse
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys  # noqa: E402

sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__cycles_active.max",
    "sm__inst_executed.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


EXTRA_METRICS = [
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
]


def to_bytes(value: str, unit: str) -> float:
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value) * mult


def summarize_launches(path: str, out_path: str, tag: str, cmd: str) -> None:
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    ci = {n: i for i, n in enumerate(rows[0])}
    d = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ci["Kernel Name"]]].append(float(r[ci["Metric Value"]]))
        except (ValueError, IndexError):
            pass
    tot = sum(sum(v) for v in d.values())
    with open(out_path, "w") as f:
        f.write(f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none ... {cmd}\n")
        f.write("# per-launch device time, cold-cache and serialised by ncu: compare SHARES, not absolutes\n")
        for k, v in sorted(d.items(), key=lambda x: -sum(x[1])):
            f.write(f"{k[:96]:96s} launches={len(v):4d} avg_us={sum(v) / len(v) / 1000:8.2f} share={sum(v) / tot * 100:5.1f}%\n")
        f.write("# at::FillFunctor = bench.py's 256 MiB L2 flush between timed steps (outside the timed event pairs)\n")
        ours = {k: v for k, v in d.items() if "bnv::" in k or "rollout_kernel" in k or "normalize_weights" in k}
        tot_ours = sum(sum(v) for v in ours.values()) or 1.0
        for k, v in ours.items():
            f.write(f"# share of the step (our kernels only): {k[:70]} {sum(v) / tot_ours * 100:5.1f}%\n")
    print(open(out_path).read())


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--workload", default="G256_K16384_T50")
    ap.add_argument("--cmd", default="python bench.py --steps 150 --warmup 10")
    ap.add_argument("--out-dir", default=None, help="write the summaries here instead of profiles/ (on the GPU box: "
                                                    "gpurun_out/summaries, the only directory that travels back)")
    args = ap.parse_args()
    global PROF
    if args.out_dir:
        PROF = os.path.abspath(args.out_dir)
    os.makedirs(PROF, exist_ok=True)

    for csv_name, suffix, cmd in (("launches.csv", "launches_summary", args.cmd),
                                  ("launches_c2.csv", "launches_c2_summary", "python bench.py --config c2 --steps 30 --warmup 5")):
        summarize_launches(os.path.join(OUT, csv_name), os.path.join(PROF, f"{args.tag}_{suffix}.txt"), args.tag, cmd)
    path = os.path.join(OUT, "launches.csv.disabled")
    if os.path.exists(path):
        rows = [r for r in csv.reader(open(path)) if len(r) > 5]
        ci = {n: i for i, n in enumerate(rows[0])}
        d = collections.defaultdict(list)
        for r in rows[1:]:
            try:
                d[r[ci["Kernel Name"]]].append(float(r[ci["Metric Value"]]))
            except (ValueError, IndexError):
                pass
        tot = sum(sum(v) for v in d.values())
        with open(os.path.join(PROF, f"{args.tag}_launches_summary.txt"), "w") as f:
            f.write(f"# {args.tag}: ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 {args.cmd}\n")
            f.write("# per-launch device time, cold-cache and serialised by ncu: compare SHARES, not absolutes\n")
            for k, v in sorted(d.items(), key=lambda x: -sum(x[1])):
                f.write(f"{k[:96]:96s} launches={len(v):4d} avg_us={sum(v) / len(v) / 1000:8.2f} "
                        f"share={sum(v) / tot * 100:5.1f}%\n")
            f.write("# at::FillFunctor = bench.py's 256 MiB L2 flush between timed steps (outside the timed event pairs)\n")
            ours = {k: v for k, v in d.items() if "bnv::" in k}
            tot_ours = sum(sum(v) for v in ours.values())
            for k, v in ours.items():
                f.write(f"# share of the step (our kernels only): {k[:60]} {sum(v) / tot_ours * 100:5.1f}%\n")
        print(open(os.path.join(PROF, f"{args.tag}_launches_summary.txt")).read())

    # full captures (ncu --set full): one summary file each, and the per-launch DRAM traffic of the rollout kernel into
    # profiles/rollout_traffic.json under the key bench.py looks up, stamped with the kernel source hash
    import bench

    caps = {
        "prof_rollout": ("rollout_ncu_summary", "c1_G256_K16384_T50",
                         "latency variant, bench.py --config c1 (G=256, K=16384, T=50): python bench.py --steps 40 --warmup 10"),
        "prof_wide": ("wide_ncu_summary", "c2_G512_K131072_T50",
                      "wide variant, bench.py --config c2 on one GPU (G=512, K=131072, T=50)"),
        "prof_batch": ("batch_ncu_summary", None, "rollout_kernel<.., kBatch>: 8 environments x K=4096 x T=30 (config 3 per-GPU share)"),
        "prof_stoch": ("stoch_ncu_summary", "c4_G256_K32768_T50", "rollout_kernel<.., kStoch>: G=256, K=32768, T=50 (config 4)"),
        "prof_risk": ("risk_mc_ncu_summary", None, "risk_mc_kernel: G=256, 1000 draws per cell, CVaR"),
    }
    tpath = os.path.join(PROF, "rollout_traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for stem, (name, workload, what) in caps.items():
        rep_x = os.path.join(OUT, stem + ".ncu-rep")
        if not os.path.exists(rep_x):
            continue
        raw = subprocess.run(["ncu", "-i", rep_x, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        traffic = []
        if "dram__bytes_read.sum" in hdr:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in data]
        with open(os.path.join(PROF, f"{args.tag}_{name}.txt"), "w") as f:
            f.write(f"# {args.tag}: ncu --set full --clock-control none --import-source on -k regex:<kernel>, {what}\n")
            f.write("# one column per captured launch; ncu flushes caches between replays (cold)\n")
            if "Kernel Name" in hdr:
                f.write("# kernel: " + data[0][hdr.index("Kernel Name")][:150] + "\n")
            for m in METRICS + EXTRA_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"{m:78s} [{units[i]:>16s}] " + "  ".join(r[i] for r in data) + "\n")
            if traffic:
                f.write(f"# DRAM traffic per launch (read+write): {[int(t) for t in traffic]} bytes\n")
        print(open(os.path.join(PROF, f"{args.tag}_{name}.txt")).read())
        if traffic and workload:
            tj[workload] = {"dram_bytes_per_launch": int(sum(traffic) / len(traffic)), "source": f"profiles/{args.tag}_{name}.txt",
                            "launches": len(traffic), "kernel_source_hash": bench.kernel_source_hash()}
    json.dump(tj, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()
