"""Turn the raw ncu output of profiles/run_profile.sh (in gpurun_out/) into the tracked artefacts under profiles/:

  r1_final_launches.csv        copy of the launch list (one C2 train step)
  r1_gemm_traffic_ncu.csv      copy of the per-launch DRAM bytes of every GEMM launch
  r1_gemm_traffic.json         per-kernel averages of the above (bench.py reads roofline.traffic from it)
  r1_final_top_kernels_ncu.csv selected --set full metrics of the top kernels (widest layer)

and print the launch table for r1_final_summary.md.   python profiles/summarize_profile.py
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


def read_metric_csv(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    col = {k: i for i, k in enumerate(hdr)}
    out = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        out.append((int(r[col["ID"]]), r[col["Kernel Name"]], r[col["Metric Name"]], r[col["Metric Unit"]], v))
    return out


def to_us(unit, v):
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[unit]


def to_bytes(unit, v):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def launch_table():
    data = read_metric_csv(os.path.join(SRC, "r1_final_launches.csv"))
    agg = collections.OrderedDict()
    for _, name, _, unit, v in data:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += to_us(unit, v)
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print(f"{n} launches, {tot / 1e3:.2f} ms of kernel time under ncu\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        print(f"| `{k[:72]}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[1] / a[0]:.1f} |")


def traffic_json():
    data = read_metric_csv(os.path.join(SRC, "r1_gemm_traffic_ncu.csv"))
    per = collections.OrderedDict()
    for i, name, metric, unit, v in data:
        d = per.setdefault(i, {"name": name})
        if metric.startswith("dram__bytes"):
            d["bytes"] = d.get("bytes", 0.0) + to_bytes(unit, v)
        else:
            d["us"] = to_us(unit, v)
    groups = collections.OrderedDict()
    seen_bn = set()
    for d in per.values():
        nm = d["name"]
        m = re.search(r"gemm_rows_tc_kernel<(?:\(int\))?(\d+), (?:\(bool\))?(\d)", nm)
        if m:
            bn, fwd = int(m.group(1)), int(m.group(2))
            if fwd:
                key = "gemm_rows_tc_kernel<fwd>"
            elif bn not in seen_bn:      # first backward launch of a layer = net_output's dX (before any iteration of the state net)
                seen_bn.add(bn)
                key = "net_output/gemm_rows<bwd dX>"
            else:
                key = "gemm_rows_tc_kernel<bwd dX>"
        elif "gemm_dw_tc_kernel" in nm:
            key = "gemm_dw_tc_kernel"
        elif "gemm_dw_kernel" in nm:
            key = "net_output/gemm_dw"
        elif "gemm_rows_kernel" in nm:
            key = "net_output/gemm_rows<fwd>"
        else:
            continue
        g = groups.setdefault(key, [0, 0.0, 0.0])
        g[0] += 1
        g[1] += d.get("bytes", 0.0)
        g[2] += d.get("us", 0.0)
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one C2 "
                     "train step (profiles/run_profile.sh), B200, tensor-core kernels, paired dX launches"}
    for k, g in groups.items():
        ent = {"launches": g[0], "avg_dram_bytes_per_launch": g[1] / g[0], "avg_us_cold_cache": g[2] / g[0]}
        if k.startswith("net_output/"):
            out.setdefault("net_output", {})[k.split("/", 1)[1]] = ent
        else:
            out[k] = ent
    json.dump(out, open(os.path.join(DST, "r1_gemm_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


WANT = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def top_kernels():
    rows_out, hdr_out = [], None
    for tag in ("fwd80", "dx80", "dx64x2", "dw", "agg", "dz"):
        path = os.path.join(SRC, f"top_{tag}_raw.csv")
        if not os.path.exists(path) or os.path.getsize(path) == 0:
            print("missing", path)
            continue
        r = list(csv.reader(open(path)))
        h, u = r[0], r[1]
        cols = [i for i, k in enumerate(h) if k in WANT or ("issue_stalled" in k and k.endswith("per_issue_active.ratio"))]
        if hdr_out is None:
            hdr_out = ["capture"] + [h[i] for i in cols]
        for x in r[2:]:      # every cell carries its unit (ncu picks units per report)
            rows_out.append([tag] + [(x[i] + " " + u[i]).strip() for i in cols])
    if not hdr_out:
        return
    with open(os.path.join(DST, "r1_final_top_kernels_ncu.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(hdr_out)
        w.writerows(rows_out)
    ix = {k: i for i, k in enumerate(hdr_out)}
    print("\n| capture | kernel | time | issue slots busy | DRAM | read | write | warp instr | regs |\n|---|---|---|---|---|---|---|---|---|")
    for x in rows_out:
        g = lambda k: x[ix[k]]
        print(f"| {x[0]} | `{g('Kernel Name')[:60]}` | {g('gpu__time_duration.sum')} | "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {g('dram__bytes.sum.per_second')} | "
              f"{g('dram__bytes_read.sum')} | {g('dram__bytes_write.sum')} | {g('smsp__inst_executed.sum')} | "
              f"{g('launch__registers_per_thread')} |")


if __name__ == "__main__":
    for f in ("r1_final_launches.csv", "r1_gemm_traffic_ncu.csv"):
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    launch_table()
    traffic_json()
    top_kernels()
