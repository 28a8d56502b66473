#!/usr/bin/env python
"""Summarise the round-2 ncu exports of gpurun_out/ (tools/ncu_profile_r2.sh) into profiles/ (tracked):
  r2_ncu_launches.csv / r2_ncu_summary.txt   launch list of one C60 job incl. the J/K consumer kernels, per-kernel shares
  r2_class_pipes.txt                          per kernel class: time, FP64-pipe utilisation, warps active, registers (all launches of one pass)
  r2_ncu_<name>_details.txt                   `ncu --page details` of the full captures, key counters appended to the summary
  r2_traffic.json                             DRAM bytes of the captured launch of the dominant kernel (bench.py's roofline.traffic)"""
import csv, os, re, sys, collections, json, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rnd = "r2"


def rows_of(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
    return rows[0], rows[1:]


def ms_of(v, unit):
    return float(v.replace(",", "")) * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(unit, 1e-6)


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "")


out = []
hdr, rows = rows_of(os.path.join(G, "launches_%s.csv" % rnd))
ik, iv, iu, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[im] != "gpu__time_duration.sum":
        continue
    agg[short(r[ik])][0] += 1
    agg[short(r[ik])][1] += ms_of(r[iv], r[iu])
tot = sum(v[1] for v in agg.values())
eri = sum(v[1] for k, v in agg.items() if k.startswith("eri_"))
out.append("# ncu launch list of ONE C60 J/K job (cintb200_int2e_sph_jk, 80 GB chunks, graphs off; serialised, cold-cache: compare SHARES)\n")
out.append("# total %.1f ms in %d launches; ERI kernels %.1f ms, tile consumers (J/K digestion) %.1f ms\n" % (tot, sum(v[0] for v in agg.values()), eri, tot - eri))
out.append("%-62s %8s %10s %7s\n" % ("kernel", "launches", "ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    out.append("%-62s %8d %10.2f %6.1f%%\n" % (k[:62], v[0], v[1], 100 * v[1] / tot))
shutil.copy(os.path.join(G, "launches_%s.csv" % rnd), os.path.join(P, "%s_ncu_launches.csv" % rnd))

# per-launch counters of one plain pass
hdr, rows = rows_of(os.path.join(G, "pipes_%s.csv" % rnd))
ik, iv, iu, im, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name"), hdr.index("ID")
per = collections.defaultdict(dict)
names = {}
for r in rows:
    names[r[iid]] = short(r[ik])
    try:
        per[r[iid]][r[im]] = (float(r[iv].replace(",", "")), r[iu])
    except ValueError:
        pass
cls = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0])
for lid, m in per.items():
    if "gpu__time_duration.sum" not in m:
        continue
    t = ms_of("%r" % m["gpu__time_duration.sum"][0], m["gpu__time_duration.sum"][1])
    c = cls[names[lid]]
    c[0] += 1
    c[1] += t
    c[2] += t * m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", (0, ""))[0]
    c[3] += t * m.get("sm__warps_active.avg.pct_of_peak_sustained_active", (0, ""))[0]
    c[4] += t * m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", (0, ""))[0]
    c[5] = int(m.get("launch__registers_per_thread", (0, ""))[0])
tot2 = sum(c[1] for c in cls.values())
pipes = ["# Per kernel class over ALL launches of one plain C60 pass (ncu, serialised): time-weighted FP64-pipe utilisation\n",
         "# (sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active), resident warps and shared-memory wavefronts; total %.1f ms, job average FP64 pipe %.1f %%\n"
         % (tot2, sum(c[2] for c in cls.values()) / tot2),
         "%-58s %4s %9s %6s %9s %8s %8s %5s\n" % ("kernel", "n", "ms", "share", "fp64pipe%", "warps%", "smemwf%", "regs")]
for k, c in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    pipes.append("%-58s %4d %9.2f %5.1f%% %9.1f %8.1f %8.1f %5d\n" % (k[:58], c[0], c[1], 100 * c[1] / tot2, c[2] / c[1], c[3] / c[1], c[4] / c[1], c[5]))
open(os.path.join(P, "%s_class_pipes.txt" % rnd), "w").write("".join(pipes))

keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
for name in ("reg_psps", "coop_dpdp", "coop_dsdp", "coop_dddd", "jk_rows3", "jk_cols", "rowsum", "wide_ffff"):
    raw = os.path.join(G, "prof_%s_%s_raw.csv" % (name, rnd))
    det = os.path.join(G, "prof_%s_%s_details.txt" % (name, rnd))
    if os.path.exists(det):
        shutil.copy(det, os.path.join(P, "%s_ncu_%s_details.txt" % (rnd, name)))
    if not os.path.exists(raw):
        continue
    rr = list(csv.reader(open(raw, errors="ignore")))
    if len(rr) < 3:
        continue
    h, units, vals = rr[0], rr[1], rr[2]
    d = {h[i]: (vals[i], units[i]) for i in range(min(len(h), len(vals)))}
    out.append("\n# full capture %s: %s\n" % (name, d.get("Kernel Name", ("?",))[0][:110]))
    for k in keys:
        if k in d:
            out.append("  %-86s %16s %s\n" % (k, d[k][0], d[k][1]))
    if name == "reg_psps":
        rd, wr = float(d["dram__bytes_read.sum"][0].replace(",", "")), float(d["dram__bytes_write.sum"][0].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd *= scale.get(d["dram__bytes_read.sum"][1], 1)
        wr *= scale.get(d["dram__bytes_write.sum"][1], 1)
        dur = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        dur *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(d["gpu__time_duration.sum"][1], 1e-3)
        json.dump({"kernel": "eri_reg_kernel<1,0,1,0,2,2>", "capture": "ncu --set full --clock-control none, 5th launch of this kernel in one C60 pass (-s 4 -c 1), round 2",
                   "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr, "duration_us": dur},
                  open(os.path.join(P, "%s_traffic.json" % rnd), "w"), indent=1)
open(os.path.join(P, "%s_ncu_summary.txt" % rnd), "w").write("".join(out))
print("".join(out[:40]))
print("".join(pipes[:30]))
