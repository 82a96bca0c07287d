#!/bin/bash
# A/B on ONE box (boxes of the pool differ by several %): build/libbt_base.so = previous code,
# build/libbt_nb2.so = k_narrow batch 2 only, in-tree library = batch 2 + idle-resolve skip.  Selected with BT_LIB_SO.
set -u
O=gpurun_out; mkdir -p $O; rm -f $O/ab_*
i=0
run() { # name, lib ("" = in-tree), extra bench args
  i=$((i+1))
  if [ -z "$2" ]; then unset BT_LIB_SO; else export BT_LIB_SO=$PWD/build/$2; fi
  timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $3 > $O/ab_$(printf %02d $i)_$1.json 2> $O/ab_$(printf %02d $i)_$1.err || echo "bench $1 failed"
}
run base libbt_base.so ""
run nb2 libbt_nb2.so ""
run new "" ""
run base libbt_base.so ""
run new "" ""
run new_m16 "" "--opt cache_margin=16"
run new_m24 "" "--opt cache_margin=24"
run new_m48 "" "--opt cache_margin=48"
run new_m64 "" "--opt cache_margin=64"
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
unset BT_LIB_SO
timeout 500 python -m pytest tests -m gpu -x -q > $O/ab_pytest_new.log 2>&1; echo "pytest new exit $?"; tail -3 $O/ab_pytest_new.log
for v in base new base new; do
  if [ $v = base ]; then export BT_LIB_SO=$PWD/build/libbt_base.so; else unset BT_LIB_SO; fi
  timeout 120 python tools/batched_arena_bench.py --worlds 8192 --ticks 300 --check 2 > $O/ab_batched_$v.json 2> $O/ab_batched_$v.err || echo "batched $v failed"
  python -c "import json;d=json.loads(open('$O/ab_batched_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_tick'],d['value'],d['alone_equals_batched'])"
done
