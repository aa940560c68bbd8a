"""Summarise an .ncu-rep here (no GPU needed): headline metrics + per-source-line instruction shares."""
import collections, csv, subprocess, sys, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed.sum','sm__issue_active.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','sm__cycles_elapsed.avg','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_fmaheavy.sum','sm__inst_executed_pipe_fmalite.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active']
for w in want:
    for i, h in enumerate(hdr):
        if h == w: print(f"{w} [{units[i]}] = {[r[i] for r in data]}")
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
        v = float(data[0][i] or 0)
        if v > 3: print(f"stall {h.split('stalled_')[1].split('_per_warp')[0]}: {v:.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0]); cur = None; h2 = None
for r in csv.reader(src.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": h2 = r; ii = r.index("Instructions Executed"); it = r.index("Thread Instructions Executed"); isamp = r.index("# Samples"); continue
    if h2 is None: continue
    try: line = int(r[0])
    except ValueError: continue
    try:
        agg[(cur, line)][0] += int(r[ii] or 0); agg[(cur, line)][1] += int(r[it] or 0); agg[(cur, line)][2] += int(r[isamp] or 0)
    except ValueError: pass
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[2] for v in agg.values()) or 1
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lsc-pm_solar_miniplant_b200", "csrc")
srcs = {f: open(os.path.join(root, f)).read().split('\n') for f in ('lscpm.cu', 'lscpm_device.cuh', 'lscpm_pool.cuh')}
print(f"total warp-inst {tot}, thread-inst {sum(v[1] for v in agg.values())}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = srcs[f][l - 1].strip()[:84] if f in srcs and l - 1 < len(srcs[f]) else ''
    print(f"{f}:{l:4d} inst {100*v[0]/tot:5.2f}% thr/inst {v[1]/max(v[0],1):5.1f} samp {100*v[2]/tots:5.2f}% | {text}")
