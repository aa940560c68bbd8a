"""Turn an `ncu --set full` report of the bench command into the committed profile artefacts:
profiles/<prefix>_trace_kernel_summary.txt, <prefix>_trace_kernel_raw.csv, <prefix>_metrics.json (bench.py reads the last one).

usage: make_final_profile.py <report.ncu-rep> <launches.csv> <prefix> "<command line that was profiled>"
"""
import csv, json, os, subprocess, sys
rep, launches, prefix, command = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
summary = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, "40"], capture_output=True, text=True).stdout
open(os.path.join(out, f"{prefix}_trace_kernel_summary.txt"), "w").write("\n".join(l[:200] for l in summary.splitlines()) + "\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(out, f"{prefix}_trace_kernel_raw.csv"), "w").write(raw)
rows = list(csv.reader(raw.splitlines())); h, d = rows[0], rows[2]
g = lambda k: float(d[h.index(k)].replace(",", ""))
unit = lambda k: rows[1][h.index(k)]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# share of the trace kernel in the launch list
tot = trace = 0.0
for r in csv.reader(l for l in open(launches) if l.startswith('"')):
    if len(r) > 14 and r[12] == "gpu__time_duration.sum":
        t = float(r[14].replace(",", "")); tot += t
        if "trace_kernel" in r[4] or "trace_sorted" in r[4] or "trace_pool" in r[4]: trace += t
name = d[h.index("Kernel Name")] if "Kernel Name" in h else "trace_kernel"
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_digest: the sha256 bench.py checks before it reports these counters)
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
dirty = bool(subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "lsc-pm_solar_miniplant_b200/csrc", "include"],
                            capture_output=True, text=True).stdout.strip())
meta = {
    "kernel": name, "command": command,
    "source_sha256": bench.source_digest(), "git_head": head + ("+uncommitted kernel sources" if dirty else ""),
    "photons_per_launch": 64000000,
    "gpu__time_duration_ms": g("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[unit("gpu__time_duration.sum")],
    "dram__bytes_read.sum": g("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")],
    "dram__bytes_write.sum": g("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")],
    "launch__registers_per_thread": g("launch__registers_per_thread"),
    "blocks_per_sm": g("launch__occupancy_limit_registers"),
    "sm__warps_active.avg.pct_of_peak_sustained_active": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "smsp__inst_executed.sum": g("smsp__inst_executed.sum"),
    "smsp__thread_inst_executed.sum": round(g("smsp__inst_executed.sum") * g("smsp__thread_inst_executed_per_inst_executed.ratio")),
    "smsp__thread_inst_executed_per_inst_executed.ratio": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": g("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "smsp__issue_active.avg.pct_of_peak_sustained_active": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "sm__pipe_fma_cycles_active.pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "sm__pipe_alu_cycles_active.pct": g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    "sm__inst_executed_pipe_xu.pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    "trace_kernel_share_of_gpu_time_in_launch_list": round(trace / tot, 4) if tot else None,
}
json.dump(meta, open(os.path.join(out, f"{prefix}_metrics.json"), "w"), indent=1)
print(json.dumps(meta, indent=1))
