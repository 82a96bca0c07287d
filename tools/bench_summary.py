"""Print the interesting parts of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value %.1fM et/s  ms/step %.3f  e2e %.1fM  launches %d  iters/tick %.2f" % (
    d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["gpu_launches"], d["config"]["iterations_per_tick"]))
r = d["roofline"]
print("top kernel %s avg %.4f ms share %.2f achieved %s GB/s frac %s" % (
    r["kernel"], r["avg_launch_ms"], r["share_of_step"], r["achieved"], r["frac"]))
for k, v in sorted(r["per_kernel"].items(), key=lambda kv: -kv[1]["ms_per_tick"]):
    print("   %-16s %8.4f ms/tick  %5.2f launches/tick  avg %.4f ms  alg %s GB/s  frac %s" % (
        k, v["ms_per_tick"], v["launches_per_tick"], v["avg_launch_ms"], v["alg_gbs"], v["frac"]))
print("whole-tick algorithmic GB/s %.1f frac %.4f" % (r["whole_tick_algorithmic_gbs"], r["whole_tick_frac"]))
print("clocks", d["clocks"])
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"])
