"""Turns the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
    python tools/summarise_profiles.py <round-tag>
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles")
GO = os.path.join(REPO, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_tensor_op_utcmma.sum", "sm__pipe_tensor_subpipe_utcmma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "lts__t_bytes.sum"]


def launches(tag):
    src = os.path.join(GO, "launches_%s.csv" % tag)
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or r[ik] == "Kernel Name":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
        name = r[ik].split("(")[0].replace("void ", "").replace("idg::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, "launches_%s.md" % tag), "w") as f:
        f.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none` over bench.py train steps + eval\n\n" % tag)
        f.write("Times are cold-cache and serialised (profiler replay): compare shares, not absolutes.\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.2f | %.1f%% |\n" % (k, n, t, t / n, 100 * t / tot))
    print("wrote launches_%s.md" % tag)


def report(rep, name):
    src = os.path.join(GO, rep)
    if not os.path.exists(src):
        return None
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = r[hdr.index(k)] + " " + units[hdr.index(k)]
        out.append(d)
    json.dump(out, open(os.path.join(OUT, name), "w"), indent=1)
    print("wrote", name)
    return out


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    if tag == "r1":
        reps = (("prof_spmm_ab_v2.ncu-rep", "spmm_amazon-book_%s_ncu_full.json" % tag), ("prof_eval_ab.ncu-rep", "eval_amazon-book_%s_ncu_full.json" % tag),
                ("prof_spmm_ab.ncu-rep", "spmm_amazon-book_%s_first_kernel_ncu_full.json" % tag),
                ("prof_pair.ncu-rep", "pairloss_lightccf_B4096_%s_ncu_full.json" % tag),
                ("prof_pair_tc.ncu-rep", "pairloss_tc_lightccf_B4096_%s_ncu_full.json" % tag))
    else:
        reps = (("prof_spmm_%s.ncu-rep" % tag, "spmm_amazon-book_%s_ncu_full.json" % tag), ("prof_eval_%s.ncu-rep" % tag, "eval_amazon-book_%s_ncu_full.json" % tag),
                ("prof_nce_%s.ncu-rep" % tag, "infonce_n1923_%s_ncu_full.json" % tag))
    for rep, name in reps:
        report(rep, name)
