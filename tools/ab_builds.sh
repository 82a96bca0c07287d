#!/bin/bash
# A/B of two builds on ONE box (boxes of the pool differ by several %): the in-tree library against
# build/libbt_<name>.so variants (nvcc ... -D<flag> -o build/libbt_<name>.so), selected with BT_LIB_SO.
# Usage (under gpurun): bash tools/ab_narrow_batch.sh name1 name2 ...   ("base" = in-tree library)
set -u
O=gpurun_out; mkdir -p $O; rm -f $O/ab_*
i=0
for v in "$@"; do
  i=$((i+1))
  if [ $v = base ]; then unset BT_LIB_SO; else export BT_LIB_SO=$PWD/build/libbt_$v.so; fi
  timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/ab_$(printf %02d $i)_$v.json 2> $O/ab_$(printf %02d $i)_$v.err || echo "bench $v failed"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(f, "%.4f ms" % d["ms_per_step"], "settle %.4f" % d["config"]["settle_phase"]["ms_per_step"],
              " ".join("%s %.4f" % (k, pk[k]["ms_per_tick"]) for k in ("narrow", "decide", "push", "post")),
              d["config"]["final_state_hashes"], "e2e %.3g" % d["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable", e)
PY
