"""Source-level hot spots of one kernel from an `ncu --set full --import-source on` report: per CUDA source line the share
of warp-stall samples and of executed instructions, with the dominant stall reasons.
usage: python tools/ncu_hotspots.py gpurun_out/prof_nhood.ncu-rep nhood_apply_list [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Line No"][0]
name = [r[1] for r in rows[:hi] if r and r[0] == "Function Name"]
hdr = rows[hi]
samp, inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg, tot_s, tot_i, allst = collections.OrderedDict(), 0, 0, collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[1].strip() == "":
        continue  # SASS-view rows (no CUDA source text): the same samples again
    try:
        s, n = int(r[samp]), int(r[inst])
    except ValueError:
        continue
    a = agg.setdefault((int(r[0]), r[1].strip()[:110]), [0, 0, collections.Counter()])
    a[0] += s
    a[1] += n
    for c in stall_cols:
        try:
            v = int(r[c])
        except ValueError:
            continue
        a[2][hdr[c]] += v
        allst[hdr[c]] += v
    tot_s += s
    tot_i += n
print(f"kernel: {name[0] if name else kern}")
print(f"warp-stall samples: {tot_s}, warp instructions executed (sum over source lines): {tot_i}")
print("stall reasons, % of samples: " + ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}" for k, v in allst.most_common(8)))
print(f"{'line':>5} {'samples':>8} {'instr':>7}  source | top stalls")
for (ln, src), (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:>5} {100 * s / tot_s:7.1f}% {100 * n / tot_i:6.1f}%  {src} | " + ", ".join(f"{k[6:]}:{100 * v / tot_s:.1f}" for k, v in st.most_common(3)))
