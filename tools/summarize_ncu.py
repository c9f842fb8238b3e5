#!/usr/bin/env python
"""Turns the ncu outputs of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/:
  profiles/<tag>_ncu_launches.csv     the per-launch duration list (gpu__time_duration.sum) of `bench.py --steps 2 --warmup 3`
  profiles/<tag>_ncu_full.json        selected metrics of the full capture (one bounce-ray launch + one primary-ray launch)
  profiles/ncu_summary.json           what bench.py reads for roofline.traffic (DRAM bytes of the bounce-ray launch)
usage: python tools/summarize_ncu.py <tag>"""
import csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
def src(name):   # tools/gpu_round2.sh prefixes its outputs with the tag; tools/gpu_round.sh does not
    t = os.path.join(G, f"{tag}_{name}")
    return t if os.path.exists(t) else os.path.join(G, name)
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
rows = list(csv.reader(open(src("prof.raw.csv"))))
hdr, units = rows[0], rows[1]
launches = []
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    launches.append({"kernel": d["Kernel Name"], **{k: {"value": d[k], "unit": u[k]} for k in KEEP if k in d}})
cmd = open(src("ncu_full.log")).read()
out = {"command": "ncu --set full --clock-control none --import-source on -k regex:k_trace -s 7 -c 2 python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --sequential",
       "launches": launches}
json.dump(out, open(os.path.join(P, f"{tag}_ncu_full.json"), "w"), indent=1)
def num(x): return float(x["value"].replace(",", ""))
def to_bytes(m):
    v = num(m); u = m["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
b = max(launches, key=lambda l: num(l["gpu__time_duration.sum"]))   # of a step's two trace launches the bounce-ray one is the longer
summary = {"source": f"profiles/{tag}_ncu_full.json (ncu --set full --clock-control none, bench.py --steps 2 --warmup 3 --no-cpu --no-extras --sequential; the whole-batch bounce-ray launch of a step)",
           "kernel": b["kernel"].split("(")[0],
           "bounce_trace_dram_bytes_per_launch": int(to_bytes(b["dram__bytes_read.sum"]) + to_bytes(b["dram__bytes_write.sum"])),
           "dram_read_bytes": int(to_bytes(b["dram__bytes_read.sum"])), "dram_write_bytes": int(to_bytes(b["dram__bytes_write.sum"])),
           "duration_ms_under_ncu": num(b["gpu__time_duration.sum"]),
           "l2_hit_rate_pct": num(b["lts__t_sector_hit_rate.pct"]), "l1_hit_rate_pct": num(b["l1tex__t_sector_hit_rate.pct"]),
           "issue_slots_busy_per_cycle": num(b["smsp__issue_active.avg.per_cycle_active"]),
           "avg_threads_per_instruction": num(b["smsp__thread_inst_executed_per_inst_executed.ratio"]),
           "achieved_occupancy_pct": num(b["sm__warps_active.avg.pct_of_peak_sustained_active"]),
           "l1_data_pipe_wavefronts_pct_of_peak": num(b["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]),
           "registers_per_thread": int(num(b["launch__registers_per_thread"]))}
json.dump(summary, open(os.path.join(P, "ncu_summary.json"), "w"), indent=1)
# launch list: keep only the CSV part
with open(src("launches.csv")) as f, open(os.path.join(P, f"{tag}_ncu_launches.csv"), "w") as o:
    for line in f:
        if line.startswith('"'):
            o.write(line)
# share of the step per kernel, from the launch list
tot = {}
for r in csv.DictReader(open(os.path.join(P, f"{tag}_ncu_launches.csv"))):
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    tot[name] = tot.get(name, 0.0) + float(r["Metric Value"].replace(",", ""))
s = sum(tot.values())
shares = {k: round(v / s, 4) for k, v in sorted(tot.items(), key=lambda x: -x[1])}
json.dump({"note": "share of summed device time per kernel over the whole ncu launch list (cold-cache, serialised)", "shares": shares},
          open(os.path.join(P, f"{tag}_ncu_launch_shares.json"), "w"), indent=1)
for name in ("bench_full", "bench_reference", "bench_v0", "bench_v1", "bench_v2", "bench_cwbvh8"):
    f = src(name + ".json")
    if os.path.exists(f):
        shutil.copy(f, os.path.join(P, f"{tag}_{name}.json"))
print(json.dumps(summary, indent=1)); print(json.dumps(shares, indent=1))
