"""Copy one measurement round (scripts/measure_round.sh, tag T) from gpurun_out/ into profiles/: bench lines, launch list,
ncu --set full summary (selected metrics per kernel) and the DRAM bytes per launch bench.py quotes as `roofline.traffic`."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(f"{G}/{T}_bench.json", f"{P}/r1_bench_line.json")
shutil.copy(f"{G}/{T}_ref.json", f"{P}/r1_bench_line_reference_arm.json")
shutil.copy(f"{G}/{T}_launches.csv", f"{P}/r1_launches_bench.csv")
open(f"{P}/r1_launch_list_bench.txt", "w").write(subprocess.run([sys.executable, f"{ROOT}/scripts/launch_list.py", f"{G}/{T}_launches.csv"], capture_output=True, text=True).stdout)
raw = subprocess.run(["ncu", "-i", f"{G}/prof_{T}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); h, u = r[0], r[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [h.index(w) for w in want]
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
tr = {}
with open(f"{P}/r1_ncu_full_top_kernels.csv", "w", newline="") as f:
    w = csv.writer(f); w.writerow(want); w.writerow([u[i] for i in idx])
    for row in r[2:]:
        w.writerow([row[i] for i in idx])
        name = row[h.index("Kernel Name")].split("(")[0]
        by = sum(float(row[h.index(m)]) * scale[u[h.index(m)]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tr[name] = by
        print(name, row[h.index("gpu__time_duration.sum")], "us", row[h.index("launch__registers_per_thread")], "regs", round(by / 1e6, 2), "MB",
              row[h.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")], "% dram")
json.dump({"source": "ncu --set full --clock-control none of scripts/prof_iter.py 100000 1 (LRR_PROF_RANGE, LRR_PROF_SETTLE=15: one steady-state "
           "finest-level iteration), profiles/r1_ncu_full_top_kernels.csv; cold caches under ncu replay", "dram_bytes_per_launch": tr},
          open(f"{P}/ncu_traffic.json", "w"), indent=1)
