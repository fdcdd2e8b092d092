"""Turn the raw ncu output of profiles/run_profile_r2.sh (in gpurun_out/) into the tracked artefacts under profiles/:

  r2_launches.csv          launch list of the bench command (first 700 launches)
  r2_traffic_ncu.csv       DRAM bytes + duration of every launch of one C2 train step
  r2_traffic.json          per-kernel averages of the above (bench.py reads roofline.traffic from it)
  r2_top_kernels_ncu.csv   selected --set full metrics of the top kernels at the widest layer (D = 78)

and print the tables for r2_summary.md.      python profiles/summarize_profile_r2.py
"""
import collections
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
TO_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# bench.py's roofline candidates -> kernel name pattern
BENCH_KEYS = {
    "forward iteration (rows_tma_kernel<FWD> / tile_fwd)": r"rows_tma_kernel<\(int\)0>|rows_tma_kernel<0>",
    "dW (dw_tma_kernel)": r"dw_tma_kernel",
    "dX (rows_tma_kernel<DX>)": r"rows_tma_kernel<\(int\)1>|rows_tma_kernel<1>",
    "dz (dz_kernel)": r"dz_kernel",
    "Adj^T s (agg_stats_kernel)": r"agg_stats_kernel<\(int\)2, \(bool\)0|agg_stats_kernel<2, 0",
}
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_tc.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def read_metric_csv(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    col = {k: i for i, k in enumerate(rows[hi])}
    out = []
    for r in rows[hi + 1:]:
        if len(r) < len(rows[hi]):
            continue
        out.append((int(r[col["ID"]]), r[col["Kernel Name"]], r[col["Metric Name"]], r[col["Metric Unit"]],
                    float(r[col["Metric Value"]].replace(",", ""))))
    return out


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "")[:64]


def launch_table(fname, title):
    per = collections.OrderedDict()
    for id_, name, m, u, v in read_metric_csv(os.path.join(SRC, fname)):
        k = per.setdefault(id_, {"name": name, "us": 0.0, "bytes": 0.0})
        if "time" in m:
            k["us"] = v * TO_US[u]
        else:
            k["bytes"] += v * TO_B[u]
    agg = collections.OrderedDict()
    for k in per.values():
        a = agg.setdefault(short(k["name"]), [0, 0.0, 0.0])
        a[0] += 1; a[1] += k["us"]; a[2] += k["bytes"]
    tot = sum(a[1] for a in agg.values())
    print(f"\n### {title}: {len(per)} launches, {tot / 1e3:.2f} ms of kernel time under ncu\n")
    print("| kernel | launches | total us | share | avg us | avg DRAM MB |\n|---|---|---|---|---|---|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:26]:
        print(f"| `{n}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[1] / a[0]:.1f} | {a[2] / a[0] / 1e6:.1f} |")
    return per


def main():
    for f in ("r2_launches.csv", "r2_traffic_ncu.csv"):
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
    launch_table("r2_launches.csv", "launch list of `bench.py --steps 2 --warmup 3` (first 700 launches)")
    per = launch_table("r2_traffic_ncu.csv", "one C2 train step (profiles/prof_step.py 8192 1), DRAM bytes per launch")
    tj = {}
    for key, pat in BENCH_KEYS.items():
        ks = [k for k in per.values() if re.search(pat, k["name"])]
        if ks:
            tj[key] = {"launches": len(ks), "avg_dram_bytes_per_launch": sum(k["bytes"] for k in ks) / len(ks),
                       "avg_us_under_ncu": sum(k["us"] for k in ks) / len(ks),
                       "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r2_traffic_ncu.csv"}
    json.dump(tj, open(os.path.join(DST, "r2_traffic.json"), "w"), indent=1)
    print("\nr2_traffic.json:", json.dumps({k: round(v["avg_dram_bytes_per_launch"] / 1e6, 1) for k, v in tj.items()}))
    # --set full captures
    out = [["capture", "kernel"] + KEEP]
    for f in sorted(os.listdir(SRC)):
        if not (f.startswith("r2_top_") and f.endswith("_raw.csv")):
            continue
        rows = list(csv.reader(open(os.path.join(SRC, f))))
        if len(rows) < 3:
            continue
        h, v = rows[0], rows[2]
        col = {k: i for i, k in enumerate(h)}
        out.append([f[7:-8], short(v[col["Kernel Name"]])] + [v[col[k]] if k in col else "" for k in KEEP])
    with open(os.path.join(DST, "r2_top_kernels_ncu.csv"), "w", newline="") as fh:
        csv.writer(fh).writerows(out)
    print("\n### --set full captures (widest layer)\n")
    print("| capture | time us | DRAM MB (r+w) | DRAM % | L2 % | L1 % | L1 hit % | warps active % | issue/cycle | long-scoreboard stall |\n|---|---|---|---|---|---|---|---|---|---|")
    for r in out[1:]:
        d = dict(zip(out[0], r))
        g = lambda k: float(d[k].replace(",", "")) if d.get(k) else float("nan")
        print(f"| {d['capture']} `{d['kernel'][:40]}` | {g('gpu__time_duration.sum'):.1f} | {g('dram__bytes_read.sum') + g('dram__bytes_write.sum'):.1f} | "
              f"{g('dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('l1tex__t_sector_hit_rate.pct'):.1f} | "
              f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {g('smsp__issue_active.avg.per_cycle_active'):.2f} | "
              f"{g('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):.1f} |")


if __name__ == "__main__":
    main()
