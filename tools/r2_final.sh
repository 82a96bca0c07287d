#!/bin/bash
# Round-2 final measurement on one B200 (run under gpurun): both bench arms, then ncu launch lists and --set full
# captures of the settled tick, the dynamic phase, the raycast and the one-block-per-world step.  Outputs -> gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O; P=${P:-r2k}
if [ -z "${SKIP_REF:-}" ]; then python bench.py --impl reference --steps 20 --warmup 3 > $O/${P}_reference_arm.json 2> $O/${P}_ref.err || echo "reference arm failed"; fi
python bench.py --steps 100 --warmup 5 > $O/${P}_bench.json 2> $O/${P}_bench.err || echo "bench failed"
python tools/arena_curve.py > $O/${P}_arena_curve.txt 2>&1
python tools/arena_curve.py fused_small=0 >> $O/${P}_arena_curve.txt 2>&1
for m in dynamic steady raycast; do python tools/ncu_target2.py $m > $O/${P}_target_$m.log 2>&1 || { echo "$m target failed"; tail -5 $O/${P}_target_$m.log; }; done
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/${P}_launches_dynamic.csv python tools/ncu_target2.py dynamic 3 > $O/${P}_ncu_l1.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${P}_launches_steady.csv python tools/ncu_target2.py steady 4 > $O/${P}_ncu_l2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_bounds|k_morton|k_sort|k_lbvh|k_refit|k_pairs|k_begin|k_integrate' -c 40 -o $O/${P}_full_dynamic \
    python tools/ncu_target2.py dynamic 1 > $O/${P}_ncu_f1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_narrow|k_push$|k_decide|k_post|k_begin_fused' -c 22 -o $O/${P}_full_steady \
    python tools/ncu_target2.py steady 1 > $O/${P}_ncu_f2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_raycast' -c 2 -o $O/${P}_full_raycast \
    python tools/ncu_target2.py raycast > $O/${P}_ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tick_small' -s 30 -c 2 -o $O/${P}_full_small \
    python tools/arena_curve.py only=8192 > $O/${P}_ncu_f4.log 2>&1
# the reports (with source) exceed what gpurun brings back: keep their raw pages as CSV
for r in ${P}_full_dynamic ${P}_full_steady ${P}_full_raycast ${P}_full_small; do
  if [ -f $O/$r.ncu-rep ]; then ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2> /dev/null; rm -f $O/$r.ncu-rep; fi
done
ls -la $O | grep ${P}
