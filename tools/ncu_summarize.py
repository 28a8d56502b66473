#!/usr/bin/env python
"""Summarise gpurun_out/ ncu exports into profiles/ (tracked): per-kernel share of the launch list and the key
counters of every full-set capture.  usage: tools/ncu_summarize.py r1"""
import csv, os, re, sys, collections
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = []
lf = os.path.join(G, "launches_%s.csv" % rnd)
if os.path.exists(lf):
    rows = [r for r in csv.reader(open(lf, errors="ignore")) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = re.sub(r"\(.*", "", r[ik])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append("# ncu launch list of ONE C60 job (serialised, cold-cache: compare SHARES)  total %.1f ms in %d launches\n" % (tot, sum(v[0] for v in agg.values())))
    out.append("%-70s %8s %10s %7s\n" % ("kernel", "launches", "ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-70s %8d %10.2f %6.1f%%\n" % (k[:70], v[0], v[1], 100 * v[1] / tot))
    import shutil
    shutil.copy(lf, os.path.join(P, "%s_ncu_launches.csv" % rnd))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for f in sorted(os.listdir(G)):
    m = re.match(r"prof_(\w+)_%s_raw.csv" % rnd, f)
    if not m:
        continue
    rows = list(csv.reader(open(os.path.join(G, f), errors="ignore")))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = dict(zip(hdr, zip(vals, units)))
    out.append("\n# ncu --set full: %s   kernel %s\n" % (m.group(1), d.get("Kernel Name", ("?",))[0][:110]))
    for k in keys:
        if k in d:
            out.append("  %-82s %16s %s\n" % (k, d[k][0], d[k][1]))
    import shutil
    shutil.copy(os.path.join(G, "prof_%s_%s_details.txt" % (m.group(1), rnd)), os.path.join(P, "%s_ncu_%s_details.txt" % (rnd, m.group(1))))
# DRAM traffic of the dominant kernel's captured launch (bench.py's roofline.traffic): ncu capture `reg_psps` = the
# (skip+1)-th launch of eri_reg_kernel<1,0,1,0,2,2> in one pass; its algorithmic bytes come from the launch list
try:
    import json
    import numpy as np
    f = os.path.join(G, "prof_reg_psps_%s_raw.csv" % rnd)
    rows = list(csv.reader(open(f, errors="ignore")))
    d = dict(zip(rows[0], zip(rows[-1], rows[1])))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * scale[d["dram__bytes_read.sum"][1]]
    wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * scale[d["dram__bytes_write.sum"][1]]
    tr = {"kernel": "eri_reg_kernel<1,0,1,0,2,2>", "capture": "ncu --set full --clock-control none, 9th launch of this kernel in one C60 pass (-s 8 -c 1)",
          "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
          "duration_us": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["gpu__time_duration.sum"][1], 1.0)}
    lr = os.path.join(G, "launch_rows.npy")
    if os.path.exists(lr):
        L = np.load(lr)
        mine = [r for r in L if tuple(int(v) for v in r[1:7]) == (1, 0, 1, 0, 2, 2)]
        if len(mine) > 8:
            tr["algorithmic_store_bytes"] = float(mine[8][10]) * 8
            tr["model_flops"] = float(mine[8][11])
    json.dump(tr, open(os.path.join(P, "%s_traffic.json" % rnd), "w"), indent=1)
    out.append("\n# roofline.traffic source: %s\n" % json.dumps(tr))
except Exception as e:
    out.append("\n# no traffic figure: %r\n" % (e,))
open(os.path.join(P, "%s_ncu_summary.txt" % rnd), "w").write("".join(out))
print("".join(out))
