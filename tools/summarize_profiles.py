"""Turn the ncu reports brought back in gpurun_out/ into small tracked summaries under profiles/ (round tag as argv[1])."""
import csv, io, json, os, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")  # argv[2]: output directory
os.makedirs(P, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "dram__sectors_read.sum", "dram__sectors_write.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
md = [f"# ncu summaries ({tag})", "", "Source: `ncu --set full --clock-control none --import-source on` on a B200 via gpurun; numbers are per launch.", ""]
for name in ("prof_nhood", "prof_moran", "prof_cooc", "prof_ripley", "prof_misc"):
    rep = os.path.join(G, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out_rows = [["metric", "unit"] + [r[idx["Kernel Name"]].split("(")[0] for r in rows[2:]]]
    for k in KEYS + stall_cols:
        if k in idx:
            out_rows.append([k, units[idx[k]]] + [r[idx[k]] for r in rows[2:]])
    with open(os.path.join(P, f"{tag}_{name}_metrics.csv"), "w", newline="") as f:
        csv.writer(f).writerows(out_rows)
    md.append(f"## {name}")
    for r in rows[2:]:
        kn = r[idx["Kernel Name"]].split("(")[0]
        g = lambda k: r[idx[k]] if k in idx else "n/a"
        dram = (float(g("dram__bytes_read.sum")) + float(g("dram__bytes_write.sum"))) if "dram__bytes_read.sum" in idx else float("nan")
        stalls = sorted(((float(r[idx[c]]), c.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for c in stall_cols), reverse=True)[:4]
        md.append(f"* `{kn}`: {g('gpu__time_duration.sum')} {units[idx['gpu__time_duration.sum']]}, grid {g('launch__grid_size')} x {g('launch__block_size')}, "
                  f"{g('launch__registers_per_thread')} regs, DRAM read+write {dram:.3f} {units[idx['dram__bytes_read.sum']]} "
                  f"({g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} % of peak), L2 hit {g('lts__t_sector_hit_rate.pct')} %, "
                  f"issue active {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} %, warp instructions {float(g('smsp__inst_executed.sum')):.3e}, "
                  f"tensor pipe {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} %, top stalls: "
                  + ", ".join(f"{n} {v:.2f}" for v, n in stalls))
    md.append("")
for f in ("bench.json", "tune_nhood.json", "gpu_info.txt", "summary.txt"):
    src = os.path.join(G, f)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{tag}_{f}"))
# launch list: keep kernel name + duration, aggregate per kernel
ll = os.path.join(G, "launches.csv")
if os.path.exists(ll):
    txt = [l for l in open(ll) if l.startswith('"')]
    rows = list(csv.reader(txt))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[1:]:
        try:
            k = r[idx["Kernel Name"]].split("(")[0][:80]
            v = float(r[idx["Metric Value"]])
            u = r[idx["Metric Unit"]]
        except Exception:
            continue
        v = v / 1e3 if u in ("usecond", "us") else (v / 1e6 if u in ("nsecond", "ns") else v)  # -> ms
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_launches_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms (ncu, serialised, cold cache)", "share"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], f"{a[1]:.3f}", f"{a[1] / tot:.3f}"])
    md.append("## launch list (`ncu --metrics gpu__time_duration.sum`, bench.py --steps 2 --warmup 1): see " + f"`{tag}_launches_summary.csv`")
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md))
