#!/bin/bash
# A/B on ONE box (boxes of the pool differ by several %): in-tree library = k_narrow without the survivor queue
# (batch-2 filter, BT_NARROW_CHUNK=0), build/libbt_ch{4,8,16}.so = queue over 4 / 8 / 16 trips (nvcc ... -DBT_NARROW_CHUNK=n).  Selected with BT_LIB_SO.
set -u
O=gpurun_out; mkdir -p $O; rm -f $O/ab_*
i=0
run() { # name, lib ("" = in-tree), extra bench args
  i=$((i+1))
  if [ -z "$2" ]; then unset BT_LIB_SO; else export BT_LIB_SO=$PWD/build/$2; fi
  timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $3 > $O/ab_$(printf %02d $i)_$1.json 2> $O/ab_$(printf %02d $i)_$1.err || echo "bench $1 failed"
}
run ch0 "" ""
run ch8 libbt_ch8.so ""
run ch4 libbt_ch4.so ""
run ch16 libbt_ch16.so ""
run ch0 "" ""
run ch8 libbt_ch8.so ""
run ch8_b592 libbt_ch8.so "--opt narrow_blocks=592"
run ch8_b296 libbt_ch8.so "--opt narrow_blocks=296"
run ch16_b592 libbt_ch16.so "--opt narrow_blocks=592"
run ch8_b2368 libbt_ch8.so "--opt narrow_blocks=2368"
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
export BT_LIB_SO=$PWD/build/libbt_ch8.so
timeout 500 python -m pytest tests -m gpu -x -q > $O/ab_pytest_new.log 2>&1; echo "pytest new exit $?"; tail -3 $O/ab_pytest_new.log
for v in base new base new; do
  if [ $v = base ]; then unset BT_LIB_SO; else export BT_LIB_SO=$PWD/build/libbt_ch8.so; fi
  timeout 120 python tools/batched_arena_bench.py --worlds 8192 --ticks 300 --check 2 > $O/ab_batched_$v.json 2> $O/ab_batched_$v.err || echo "batched $v failed"
  python -c "import json;d=json.loads(open('$O/ab_batched_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_tick'],d['value'],d['alone_equals_batched'])"
done
