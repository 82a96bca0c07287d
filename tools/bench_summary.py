"""Print the interesting parts of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value %.1fM et/s  ms/step %.3f  e2e %.1fM  launches %d  iters/tick %.2f" % (
    d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["gpu_launches"], d["config"]["iterations_per_tick"]))
r = d["roofline"]
print("top kernel %s avg %.4f ms share %.2f achieved %s GB/s frac %s" % (
    r["kernel"], r["avg_launch_ms"], r["share_of_step"], r["achieved"], r["frac"]))
for k, v in sorted(r["per_kernel_ms_per_tick"].items(), key=lambda kv: -kv[1]):
    print("   %-16s %8.4f ms/tick  %5.2f launches/tick" % (k, v, r["per_kernel_launches_per_tick"][k]))
print("clocks", d["clocks"])
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"])
