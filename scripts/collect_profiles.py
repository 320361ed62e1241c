"""Turns the files scripts/final_profile.sh leaves in gpurun_out/ into the tracked summaries under profiles/:
ncu_r1_final_details.csv (selected metrics of the --set full captures), scan_traffic.json (DRAM traffic of one scan launch, read by
bench.py for roofline.traffic), launches_r1_final.csv, bench lines.  Prints the launch table of the timed step for r1_summary.md."""
import csv, json, shutil, subprocess, sys, collections
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
METRICS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio""".split()

def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

out_rows = [("capture", "metric", "value", "unit")]
scan_bytes = None
for rep, label in ((G / "prof_scan_final.ncu-rep", "scan, 1000 x 5 Mb"), (G / "prof_dp_final.ncu-rep", "DP kernels, 250 x 5 Mb")):
    hdr, units, launches = raw(rep)
    for vals in launches:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        out_rows.append((label, "Kernel Name", d["Kernel Name"], ""))
        for m in METRICS:
            if m in d: out_rows.append((label, m, d[m], u[m]))
        if "kb_scan" in d["Kernel Name"]:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            scan_bytes = (float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]],
                          float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]])
with open(P / "ncu_r1_final_details.csv", "w", newline="") as f:
    csv.writer(f).writerows(out_rows)
if scan_bytes:
    json.dump({"assemblies_per_launch": 1000, "asm_len": 5000000, "dram_bytes_read": int(scan_bytes[0]), "dram_bytes_write": int(scan_bytes[1]),
               "source": "ncu --set full --clock-control none -k regex:kb_scan_kernel -s 1 -c 1 python bench.py --steps 1 --warmup 1 (profiles/ncu_r1_final_details.csv)"},
              open(P / "scan_traffic.json", "w"), indent=1)
shutil.copy(G / "launches_r1_final.csv", P / "launches_r1_final.csv")
shutil.copy(G / "final_bench.json", P / "bench_r1_final.json")
shutil.copy(G / "final_bench_reference.json", P / "bench_r1_final_reference.json")
# launch table of the timed step: the run is warmup 1 + step 1 + e2e (warmup + step on 8 assemblies); the timed main step is the
# second quarter of the kernel list
rows = [r for r in csv.reader(open(G / "launches_r1_final.csv")) if len(r) > 10 and r[0].isdigit()]
names = [r[4] for r in rows]; t = [float(r[-1]) for r in rows]; unit = rows[0][-2] if rows else ""
n = len(rows) // 4
seg = list(zip(names[n:2 * n], t[n:2 * n]))
acc = collections.OrderedDict()
for k, v in seg:
    k = k.split("(")[0][:50]
    acc.setdefault(k, [0, 0.0]); acc[k][0] += 1; acc[k][1] += v
tot = sum(v for _, v in acc.values())
print("unit", unit, "total", tot)
for k, (c, v) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"| {k} | {c} | {v / (1e6 if unit in ('ns', 'nsecond') else 1e3 if unit in ('us', 'usecond') else 1):.3f} | {100 * v / tot:.1f} % |")
