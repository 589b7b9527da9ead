"""Turn an `ncu --set full` report into the text summary committed under profiles/ and update profiles/ncu_traffic.json.

  python scripts/ncu_summary.py gpurun_out/prof_field.ncu-rep field_fwd_kernel 262144 profiles/r01_v6_ncu_field_fwd.txt
"""
import csv, io, json, os, subprocess, sys

rep, kernel, units, out = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
m = {n: (v[i], u[i]) for i, n in enumerate(h)}
def f(name):
    val, unit = m[name]
    x = float(val.replace(",", ""))
    return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(unit, 1)
dram = f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct"]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
with open(out, "w") as fo:
    fo.write(f"# ncu --set full --clock-control none, kernel {v[h.index('Kernel Name')]} ({units} units per launch)\n")
    fo.write(f"# dram bytes per launch (read+write) = {dram:.0f}\n\n## selected raw metrics\n")
    for k in keep:
        if k in m:
            fo.write(f"{k:70s} {m[k][0]} {m[k][1]}\n")
    fo.write("\n## details page\n")
    fo.write("\n".join(l for l in det.splitlines() if l.strip()))
tj = os.path.join(R, "profiles", "ncu_traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t[f"{kernel}@{units}"] = {"dram_bytes": dram, "duration_us": f("gpu__time_duration.sum") / (1e3 if m["gpu__time_duration.sum"][1] == "ns" else 1), "report": os.path.basename(out)}
json.dump(t, open(tj, "w"), indent=1)
print(out, dram)
