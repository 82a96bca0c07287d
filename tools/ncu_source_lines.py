"""Samples and instructions of one kernel by source line, from `ncu --page source --csv` (SASS rows) and the line table of
the same object (`cuobjdump -xelf all obj; nvdisasm -gi -c x.cubin > x.sass`): by call site in the kernel's own file
(outermost inlined-at) and by innermost line.
python tools/ncu_source_lines.py source.csv x.sass kernel_file.cu [top [mangled kernel name, when x.sass holds several]]"""
import collections
import csv
import re
import sys

csv.field_size_limit(10 ** 9)
src, sass, own = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
only = sys.argv[5] if len(sys.argv) > 5 else None
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
base = int(data[0][0], 16)
col = {k: hdr.index(k) for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_barrier")}
loc, cur, fresh = {}, [], True
section = None
for line in open(sass):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        section = m.group(1)
        continue
    if only and section != only:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        if fresh:
            cur, fresh = [], False
        cur.append((m.group(1).split("/")[-1], int(m.group(2))))
        if m.group(3):
            cur.append((m.group(3).split("/")[-1], int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", line)
    if m:
        loc[int(m.group(1), 16)] = list(cur)
        fresh = True
tot_s = sum(int(r[col["# Samples"]]) for r in data)
tot_i = sum(int(r[col["Instructions Executed"]]) for r in data)
print("samples %d, warp instructions %d, SASS rows %d (%d with a line)" % (tot_s, tot_i, len(data), len(loc)))
agg = {"outer": collections.defaultdict(lambda: [0, 0, 0, 0, 0]), "inner": collections.defaultdict(lambda: [0, 0, 0, 0, 0])}
for r in data:
    chain = loc.get(int(r[0], 16) - base, [])
    mine = [x for x in chain if x[0] == own]
    keys = {"outer": mine[-1] if mine else ("?", 0), "inner": chain[0] if chain else ("?", 0)}
    v = [int(r[col[k]]) for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_barrier")]
    for which, k in keys.items():
        a = agg[which][k]
        for j in range(5):
            a[j] += v[j]
for which, title in (("outer", "by line of %s (call site)" % own), ("inner", "by innermost line")):
    print("\n| %s | samples %% | instructions %% | long scoreboard %% | barrier %% | threads / instruction |\n|---|---|---|---|---|---|" % title)
    for k, a in sorted(agg[which].items(), key=lambda kv: -kv[1][0])[:top]:
        print("| %s:%d | %.1f | %.1f | %.1f | %.1f | %.1f |" % (k[0], k[1], 100 * a[0] / tot_s, 100 * a[1] / tot_i, 100 * a[3] / tot_s,
                                                              100 * a[4] / tot_s, a[2] / max(1, a[1])))
