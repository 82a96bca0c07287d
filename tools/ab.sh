#!/bin/bash
# A/B of library builds on ONE box: tools/ab.sh name1 name2 ... [-- option=value ...]; "base" = the in-tree library,
# other names = build/libbt_<name>.so.  Two rounds, interleaved, so that drift of the box shows.
names=(); opts=()
while [ $# -gt 0 ]; do if [ "$1" = "--" ]; then shift; opts=("$@"); break; fi; names+=("$1"); shift; done
for round in 1 2; do
  for v in "${names[@]}"; do
    if [ "$v" = base ]; then unset BT_LIB_SO; else export BT_LIB_SO=$PWD/build/libbt_$v.so; fi
    timeout 300 python tools/ab_quick.py "${opts[@]}" 2>&1 | tail -1
  done
done
