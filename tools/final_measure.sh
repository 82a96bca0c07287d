#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): bench both arms, batched worlds, raycast,
# then the ncu launch list and one --set full capture of a settled tick.  Outputs -> gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
python bench.py --steps 100 --warmup 5 > $O/rf_bench.json 2> $O/rf_bench.err || echo "bench failed"
python bench.py --impl reference --steps 20 --warmup 3 > $O/rf_reference_arm.json 2> $O/rf_ref.err || echo "reference arm failed"
python tools/batched_arena_bench.py --worlds 8192 --ticks 300 --check 2 > $O/rf_batched.json 2> $O/rf_batched.err || echo "batched failed"
python tools/raycast_bench.py > $O/rf_raycast.json 2> $O/rf_raycast.err || echo "raycast failed"
python tools/ncu_target.py 1000000 2 56 > $O/rf_target.log 2>&1 || { echo "target failed"; exit 1; }
# ncu_target brackets its measured ticks with cuProfilerStart/Stop: only those settled ticks are captured
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/rf_launches_steady.csv python tools/ncu_target.py 1000000 4 56 > $O/rf_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_narrow|k_push$|k_decide|k_post|k_begin_fused' -c 22 -o $O/rf_full \
    python tools/ncu_target.py 1000000 1 56 > $O/rf_ncu2.log 2>&1
ls -la $O | tail -15
