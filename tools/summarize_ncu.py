#!/usr/bin/env python
"""Summarise ncu captures into profiles/ (tracked):
    python tools/summarize_ncu.py full  <report.ncu-rep> <tag>    -> profiles/<tag>_summary.csv (+ profiles/igemm_traffic.json)
    python tools/summarize_ncu.py list  <launches.csv>   <tag>    -> profiles/<tag>_per_step.md   (per-kernel share of one PC step)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers"]


def short(k):
    return k.replace("<unnamed>::", "").replace("void ", "").split("(")[0]


def full(rep, tag):
    if rep.endswith(".csv"):          # raw page already exported on the GPU box (tools/round_gpu_run.sh)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = [hdr.index(k) for k in KEEP if k in hdr]
    out = os.path.join(ROOT, "profiles", f"{tag}_summary.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in ix])
        w.writerow([units[i] for i in ix])
        for r in body:
            w.writerow([short(r[i]) if hdr[i] == "Kernel Name" else r[i] for i in ix])
    kn, rd, wr, du = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    assert units[rd] == "Mbyte" and units[wr] == "Mbyte", (units[rd], units[wr])
    agg = collections.defaultdict(list)
    for r in body:
        agg["igemm" if "igemm" in r[kn] else ("gn_apply" if "gn_apply" in r[kn] else "other")].append(
            (float(r[rd]) + float(r[wr])) * 1e6)
    js = {"source": f"profiles/{tag}_summary.csv: ncu --set full --clock-control none, mean of (dram__bytes_read.sum + dram__bytes_write.sum) over "
                    f"the {len(agg['igemm'])} igemm launches captured at the start of one score-network forward (stem, level-0 and first "
                    f"level-1 res-blocks: the largest tensors), batch 128",
          "dram_bytes_per_launch": sum(agg["igemm"]) / max(len(agg["igemm"]), 1),
          "igemm_launches_captured": len(agg["igemm"]),
          "gn_apply_dram_bytes_per_launch": sum(agg["gn_apply"]) / max(len(agg["gn_apply"]), 1),
          "gn_apply_launches_captured": len(agg["gn_apply"])}
    with open(os.path.join(ROOT, "profiles", "igemm_traffic.json"), "w") as f:
        json.dump(js, f, indent=1)
    print(out, js)


def launch_list(path, tag):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if l.startswith('"')))
    names = [r["Kernel Name"] for r in rows]
    ns = [float(r["Metric Value"]) for r in rows]
    marks = [i for i, k in enumerate(names) if "predictor_update_kernel" in k]
    lines = [f"# {tag}: per-kernel device time of ONE PC sampling step (score-network forward + fused predictor update)",
             "",
             f"Source: `profiles/{os.path.basename(path)}` = `ncu --metrics gpu__time_duration.sum --clock-control none` over "
             "`python bench.py --steps 1 --warmup 1 --num-scales 2 --skip-train --skip-cpu --skip-extras` (the bench command with a 2-step schedule). "
             f"{len(rows)} launches captured before the time limit; steps are delimited by `predictor_update_kernel`. ncu serialises "
             "launches and runs them cold-cache: compare SHARES with bench.py's live `share_of_forward_device_time`, not absolutes.",
             ""]
    for a, b in list(zip(marks[:-1], marks[1:]))[-1:]:
        seg = range(a + 1, b + 1)
        tot = sum(ns[i] for i in seg)
        agg, cnt = collections.Counter(), collections.Counter()
        for i in seg:
            agg[short(names[i])] += ns[i]
            cnt[short(names[i])] += 1
        classes = collections.Counter()
        for k, v in agg.items():
            classes["igemm_kernel + igemm_halo_kernel (all instantiations)" if k.startswith("igemm") else
                    ("gn_apply_kernel (all)" if k.startswith("gn_apply") else "everything else")] += v
        lines += [f"One step = {len(list(seg))} launches, {tot / 1e3:.1f} us summed kernel time.", "", "| class | us | share |", "|---|---|---|"]
        lines += [f"| {k} | {v / 1e3:.1f} | {100 * v / tot:.1f} % |" for k, v in classes.most_common()]
        lines += ["", "| kernel | launches | us | share |", "|---|---|---|---|"]
        lines += [f"| `{k}` | {cnt[k]} | {v / 1e3:.1f} | {100 * v / tot:.1f} % |" for k, v in agg.most_common()]
    out = os.path.join(ROOT, "profiles", f"{tag}_per_step.md")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
