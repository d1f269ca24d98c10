"""Copy one measurement round (scripts/measure_round.sh, tag T) from gpurun_out/ into profiles/: bench lines, launch list,
ncu --set full summary (selected metrics per kernel) and the DRAM bytes per launch bench.py quotes as `roofline.traffic`."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = sys.argv[1]
R = sys.argv[2] if len(sys.argv) > 2 else T[:2]   # round prefix of the files under profiles/ (r1, r2, ...)
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(f"{G}/{T}_bench.json", f"{P}/{R}_bench_line.json")
shutil.copy(f"{G}/{T}_ref.json", f"{P}/{R}_bench_line_reference_arm.json")
shutil.copy(f"{G}/{T}_launches.csv", f"{P}/{R}_launches_bench.csv")
open(f"{P}/{R}_launch_list_bench.txt", "w").write(subprocess.run([sys.executable, f"{ROOT}/scripts/launch_list.py", f"{G}/{T}_launches.csv"], capture_output=True, text=True).stdout)
raw = subprocess.run(["ncu", "-i", f"{G}/prof_{T}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); h, u = r[0], r[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [h.index(w) for w in want]
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
tr = {}
with open(f"{P}/{R}_ncu_full_top_kernels.csv", "w", newline="") as f:
    w = csv.writer(f); w.writerow(want); w.writerow([u[i] for i in idx])
    for row in r[2:]:
        w.writerow([row[i] for i in idx])
        name = row[h.index("Kernel Name")].split("(")[0]
        by = sum(float(row[h.index(m)]) * scale[u[h.index(m)]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tr[name] = by
        print(name, row[h.index("gpu__time_duration.sum")], "us", row[h.index("launch__registers_per_thread")], "regs", round(by / 1e6, 2), "MB",
              row[h.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")], "% dram")
tr_fixed = dict(tr)
# the in-registration capture (LRR_PROF_INREG=2: finest-level iterations 2.. of the ramped registration, transform still moving)
inreg = f"{G}/prof_{T}_inreg.ncu-rep"
tr_in = {}
if os.path.exists(inreg):
    raw = subprocess.run(["ncu", "-i", inreg, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines())); h, u = r[0], r[1]
    idx = [h.index(w) for w in want]
    agg = {}
    with open(f"{P}/{R}_ncu_full_inreg_kernels.csv", "w", newline="") as f:
        w = csv.writer(f); w.writerow(want); w.writerow([u[i] for i in idx])
        for row in r[2:]:
            w.writerow([row[i] for i in idx])
            name = row[h.index("Kernel Name")].split("(")[0]
            by = sum(float(row[h.index(m)]) * scale[u[h.index(m)]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += by; a[2] += float(row[h.index("gpu__time_duration.sum")])
    print("in-registration capture:")
    for name, (cnt, by, us) in agg.items():
        tr_in[name] = by / cnt
        print(f"  {name}: {cnt} launches, {us / cnt:.1f} us, {by / cnt / 1e6:.2f} MB per launch")
first = f"{G}/prof_{T}_first.ncu-rep"
tr_first = {}
if os.path.exists(first):   # the first finest-level iteration of the registration (LRR_PROF_INREG=1 -c 10): first launch of every kernel
    raw = subprocess.run(["ncu", "-i", first, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines())); h, u = r[0], r[1]
    idx = [h.index(w) for w in want]
    with open(f"{P}/{R}_ncu_full_first_finest_iteration.csv", "w", newline="") as f:
        w = csv.writer(f); w.writerow(want); w.writerow([u[i] for i in idx])
        for row in r[2:]:
            w.writerow([row[i] for i in idx])
            name = row[h.index("Kernel Name")].split("(")[0]
            if name not in tr_first:
                tr_first[name] = sum(float(row[h.index(m)]) * scale[u[h.index(m)]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                print(f"  first finest-level iteration: {name} {row[h.index('gpu__time_duration.sum')]} us, {tr_first[name] / 1e6:.2f} MB")
json.dump({"source": f"ncu --set full --clock-control none of scripts/prof_iter.py 100000 1; cold caches under ncu replay. in_registration: LRR_PROF_INREG=2, "
           f"finest-level iterations 2.. of the ramped registration (profiles/{R}_ncu_full_inreg_kernels.csv); fixed_point: LRR_PROF_RANGE, LRR_PROF_SETTLE=15 "
           f"(profiles/{R}_ncu_full_top_kernels.csv); first_finest_iteration: LRR_PROF_INREG=1 -c 10, the first finest-level iteration of the registration, no "
           f"certificates yet (profiles/{R}_ncu_full_first_finest_iteration.csv)",
           "dram_bytes_per_launch": tr_in or tr_fixed, "dram_bytes_per_launch_fixed_point": tr_fixed, "dram_bytes_per_launch_first_finest_iteration": tr_first},
          open(f"{P}/ncu_traffic.json", "w"), indent=1)
