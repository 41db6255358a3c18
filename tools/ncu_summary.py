#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall reasons per issue, SASS opcode mix."""
import collections
import csv
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']


def run(rep, page):
    return subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(run(rep, 'raw').splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f"## {rep}")
    print("kernel:", vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:72s} {vals[i]} {units[i]}")
    print("-- stall reasons (warps per issue-active cycle)")
    st = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
            try:
                st.append((float(vals[i]), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    for v, n in sorted(st, reverse=True)[:10]:
        print(f"   {n:28s} {v:.3f}")
    out = run(rep, 'source').splitlines()
    rows = list(csv.reader(out[1:]))
    hdr, rows = rows[0], rows[1:]
    iS, iSamp, iEx = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    agg = collections.defaultdict(lambda: [0, 0])
    ts = te = 0
    for r in rows:
        if len(r) <= iEx:
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS].strip())
        op = m.group(2) if m else '?'
        base = op.split('.')[0]
        key = '.'.join(op.split('.')[:3]) if base in ('LDS', 'STS', 'LDG', 'STG') else base
        s, e = int(r[iSamp] or 0), int(r[iEx] or 0)
        agg[key][0] += s
        agg[key][1] += e
        ts += s
        te += e
    print(f"-- SASS mix: {te} warp instructions, {len(rows)} static")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {k:22s} exec {v[1]:10d} ({v[1] / te:5.1%})  stall samples {v[0] / max(ts, 1):5.1%}")


if __name__ == '__main__':
    for r in sys.argv[1:]:
        main(r)
