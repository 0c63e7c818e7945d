"""Print the headline fields of a bench.py JSON line. usage: python tools/bench_brief.py gpurun_out/x.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print("value", round(d["value"]), "pages/s  ms/step", round(d["ms_per_step"], 4), " stages(us)",
      {k: round(v["ms"] * 1e3, 1) for k, v in r.get("stages", {}).items()}, " dom frac", round(r.get("frac", 0), 3),
      " step frac", round(r.get("step", {}).get("frac", 0), 3), " e2e", round(d["e2e"]["value"]),
      " single-stream ms", round(d.get("config_detail", d["config"]).get("ms_per_step_single_stream", 0), 4))
