"""Pretty-print gpurun_out/tune_nhood.log (one JSON object per line)."""
import json, sys
for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tune_nhood.log"):
    try:
        d = json.loads(l)
    except Exception:
        print(l[:300].rstrip()); continue
    if "shuffle_algo" in d:
        print(d["shuffle_algo"], d["shuffle_threads"], d["shuffle_r"], "q", d["shuffle_q"], "wf", d["shuffle_wfactor_x100"], "ctas", d["shuffle_ctas"],
              "sg", d["shuffle_stagger_us"], "low", d.get("shuffle_low"), "sym", d.get("count_sym"), "| shuffle %.2f misc %.2f count %.2f ok %s" % (d["shuffle"], d["misc"], d["count"], d["ok"]))
    else:
        print(json.dumps(d)[:600])
