#!/bin/bash
# Round-2 ncu evidence on one B200 (run under gpurun): launch lists (durations) and --set full captures of
# the dynamic phase (broadphase kernels), the steady state (resolve loop) and the raycast.  Outputs -> gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
for m in dynamic steady raycast; do python tools/ncu_target2.py $m > $O/r2_target_$m.log 2>&1 || { echo "$m target failed"; tail -5 $O/r2_target_$m.log; }; done
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/r2_launches_dynamic.csv python tools/ncu_target2.py dynamic 3 > $O/r2_ncu_l1.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/r2_launches_steady.csv python tools/ncu_target2.py steady 4 > $O/r2_ncu_l2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_bounds|k_morton|k_sort|k_lbvh|k_refit|k_pairs|k_begin|k_integrate' -c 40 -o $O/r2_full_dynamic \
    python tools/ncu_target2.py dynamic 1 > $O/r2_ncu_f1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_narrow|k_push$|k_decide|k_post|k_begin_fused' -c 22 -o $O/r2_full_steady \
    python tools/ncu_target2.py steady 1 > $O/r2_ncu_f2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_raycast|k_live_ranks' -c 4 -o $O/r2_full_raycast \
    python tools/ncu_target2.py raycast > $O/r2_ncu_f3.log 2>&1
# the reports (with source) exceed what gpurun brings back: keep their raw pages as CSV
for r in r2_full_dynamic r2_full_steady r2_full_raycast; do
  if [ -f $O/$r.ncu-rep ]; then ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2> /dev/null; rm -f $O/$r.ncu-rep; fi
done
ls -la $O | tail -12
