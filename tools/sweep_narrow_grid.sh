#!/bin/bash
# k_narrow grid sweep on ONE box (options narrow_blocks / narrow_threads), bench.py's 1 M world.
set -u
O=gpurun_out; mkdir -p $O; rm -f $O/sw_*
i=0
for cfg in "1184 256" "592 256" "2368 256" "4736 256" "1184 128" "2368 128" "4736 128" "9472 128" "1184 256"; do
  set -- $cfg; i=$((i+1))
  timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --opt narrow_blocks=$1 --opt narrow_threads=$2 \
    > $O/sw_$(printf %02d $i)_b$1_t$2.json 2> $O/sw_$(printf %02d $i).err || echo "bench $cfg failed"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sw_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(f, "%.4f ms" % d["ms_per_step"], "settle %.4f" % d["config"]["settle_phase"]["ms_per_step"],
              "narrow %.4f" % pk["narrow"]["ms_per_tick"], d["config"]["final_state_hashes"])
    except Exception as e:
        print(f, "unreadable", e)
PY
