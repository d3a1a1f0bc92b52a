"""Turn the .ncu-rep files of profiles/capture.sh into the text summaries committed under profiles/.

    python profiles/summarise.py TAG     # reads gpurun_out/prof_TAG_{guide,alpha}.ncu-rep, gpurun_out/launches_TAG.csv
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def metrics(rep, title, out):
    # the .ncu-rep, or the raw page already exported on the GPU box (tools/ncu_export.sh: gpurun brings back <= 64 MiB)
    text = open(rep).read() if rep.endswith(".csv") else ncu(rep, "--page", "raw", "--csv")
    rows = list(csv.reader(l for l in text.splitlines() if l.startswith('"')))
    d = {h: (rows[2][i], rows[1][i]) for i, h in enumerate(rows[0])}
    lines = [title]
    for k in KEEP + sorted(k for k in d if "issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k):
        if k in d:
            if "issue_stalled" in k and float(d[k][0].replace(",", "") or 0) < 0.05:
                continue
            lines.append(f"{k} [{d[k][1]}] = {d[k][0]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:6]))


def main(tag):
    g = os.path.join(ROOT, "gpurun_out")
    what = {"guide": "svi_guide_kernel, c5 (1M guides x 8 x 4), step 500 of profiles/steady_state.py 600 (steady state)",
            "alpha": "svi_alpha_kernel, c5 (1M guides x 8 x 4), step 500 of profiles/steady_state.py 600 (steady state)",
            "surv_guide": "surv_guide_kernel, c4 at scale (1M guides x 3 x 3), step 200 of profiles/survival_steady.py 300",
            "tiling_guide": "tiling_guide_kernel, c3 (800 guides x <= 16 alleles x 4 replicates), step 100 of profiles/tiling_steady.py 200"}
    for kern in ("guide", "alpha", "surv_guide", "tiling_guide"):
        rep = f"{g}/prof_{tag}_{kern}.ncu-rep"
        exported = f"{g}/prof_{tag}_{kern}_raw.csv"
        if not os.path.exists(rep) and not os.path.exists(exported):
            continue
        title = f"ncu --set full --clock-control none, {what[kern]}, capture {tag}"
        src = f"{g}/prof_{tag}_{kern}_src.csv"
        if os.path.exists(rep):
            metrics(rep, title, f"{ROOT}/profiles/{tag}_{kern}_metrics.txt")
            open(src, "w").write(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))
        else:
            metrics(exported, title, f"{ROOT}/profiles/{tag}_{kern}_metrics.txt")
        top = subprocess.run([sys.executable, f"{ROOT}/profiles/agg_source.py", src, "40"], capture_output=True, text=True).stdout
        open(f"{ROOT}/profiles/{tag}_{kern}_source_top.txt", "w").write(top)
    launches = f"{g}/launches_{tag}.csv"
    if os.path.exists(launches):
        rows = [r for r in csv.reader(open(launches)) if r and r[0].isdigit()]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows:
            agg[r[4].split("(")[0][:70]][0] += 1
            agg[r[4].split("(")[0][:70]][1] += float(r[-1].replace(",", "")) * 1e-6
        open(f"{ROOT}/profiles/{tag}_launches.csv", "w").write(open(launches).read())
        tot = sum(v[1] for v in agg.values())
        summary = [f"{v[1]:10.3f} ms {v[0]:5d} launches {v[1] / tot * 100:5.1f}%  {k}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]]
        open(f"{ROOT}/profiles/{tag}_launch_summary.txt", "w").write("\n".join(summary) + "\n")
        print("\n".join(summary))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
