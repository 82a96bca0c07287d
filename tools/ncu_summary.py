"""Summarise ncu outputs into profiles/: launch list CSV -> per-kernel totals/shares;
`--set full` raw CSV -> per-launch table of the metrics that matter for an HBM-bound path.
    python tools/ncu_summary.py launches <launches.csv>
    python tools/ncu_summary.py full <report.ncu-rep>            (needs ncu on PATH)
    python tools/ncu_summary.py grep <report.ncu-rep> <kernel regex> <metric regex>   (any raw-page columns, per launch)
A report given as *.csv is read as the raw page already exported (`ncu -i x.ncu-rep --page raw --csv > x.raw.csv`, what
tools/r2_final.sh brings back from the GPU box: the reports themselves exceed the transfer limit)."""
import collections
import csv
import json
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot))
    print("\ntotal %.1f us over %d launches (cold-cache, serialised by ncu: compare shares, not absolutes)" %
          (tot / 1e3, sum(a[0] for a in agg.values())))


WANT = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active thr/inst"),
        ("smsp__inst_executed.sum", "warp inst"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-sb stall/issue"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 ld sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 ld requests")]


def raw_page(rep):
    if rep.endswith(".csv"):
        return open(rep).read()
    return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def full(rep):
    out = raw_page(rep)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(m), lab) for m, lab in WANT if m in hdr]
    ki = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join("%s" % lab for _, lab in idx) + " |")
    print("|---|" + "---|" * len(idx))
    per = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        print("| %s | " % name + " | ".join("%s %s" % (r[i][:12], units[i] if lab in ("dram rd", "dram wr", "L2 bytes") else "")
                                            for i, lab in idx) + " |")
        def val(m):
            i = hdr.index(m)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        per[name].append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    print()
    print(json.dumps({k: {"launches": len(v), "dram_bytes_per_launch": sum(v) / len(v)} for k, v in per.items()}))


def grep(rep, kernel_re, metric_re):
    """every raw-page column whose name matches metric_re, for the launches whose kernel matches kernel_re
    (what bounds a latency-looking kernel: l1tex / lts throughput %, warp stall reasons per issue)"""
    import re
    out = raw_page(rep)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [i for i, h in enumerate(hdr) if re.search(metric_re, h)]
    sel = [r for r in rows[2:] if re.search(kernel_re, r[ki])]
    print("| metric | unit | " + " | ".join("%s #%d" % (r[ki].split("(")[0], n) for n, r in enumerate(sel)) + " |")
    print("|---|---|" + "---|" * len(sel))
    for i in cols:
        print("| %s | %s | " % (hdr[i], units[i]) + " | ".join(r[i][:10] for r in sel) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full, "grep": grep}[sys.argv[1]](*sys.argv[2:])
