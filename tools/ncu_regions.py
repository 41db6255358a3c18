#!/usr/bin/env python
"""Stall samples of an .ncu-rep (source page) split into code regions delimited by block barriers (BAR.SYNC) and
by opcode class, with the per-reason breakdown for each region: where in the kernel the time goes."""
import collections
import csv
import re
import subprocess
import sys


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
    allrows = [r for r in csv.reader(out) if len(r) > 3]
    # a report with several kernels repeats the header row: one table per kernel
    starts = [i for i, r in enumerate(allrows) if 'Source' in r and '# Samples' in r]
    for n, i0 in enumerate(starts):
        i1 = starts[n + 1] if n + 1 < len(starts) else len(allrows)
        regions(f"{rep} [kernel {n}]", allrows[i0], allrows[i0 + 1:i1])


def regions(rep, hdr, rows):
    iS, iSamp, iEx = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    region, regs = 0, collections.OrderedDict()
    for r in rows:
        if len(r) <= iEx:
            continue
        src = r[iS].strip()
        d = regs.setdefault(region, {'samples': 0, 'exec': 0, 'n': 0, 'why': collections.Counter(), 'ops': collections.Counter(), 'first': src})
        s, e = int(r[iSamp] or 0), int(r[iEx] or 0)
        d['samples'] += s
        d['exec'] += e
        d['n'] += 1
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', src)
        d['ops'][m.group(2) if m else '?'] += e
        for i, h in reasons:
            d['why'][h] += int(r[i] or 0)
        if 'BAR.SYNC' in src:
            region += 1
    tot = sum(d['samples'] for d in regs.values()) or 1
    tote = sum(d['exec'] for d in regs.values()) or 1
    print(f"## {rep}: {tot} samples, {tote} warp instructions")
    for k, d in regs.items():
        if d['samples'] < tot * 0.005:
            continue
        why = ', '.join(f"{h[6:]} {v / d['samples']:.0%}" for h, v in d['why'].most_common(6) if v)
        ops = ', '.join(f"{o} {v / d['exec']:.0%}" for o, v in d['ops'].most_common(5) if v)
        print(f"region {k}: {d['n']} static, exec {d['exec'] / tote:5.1%}, samples {d['samples'] / tot:5.1%}  | {ops}")
        print(f"      stalls: {why}")


if __name__ == '__main__':
    for r in sys.argv[1:]:
        main(r)
