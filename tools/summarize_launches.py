#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (and per
grid size) launch count, total / mean duration and share of the whole list.

    python tools/summarize_launches.py gpurun_out/launches.csv [--last-step N] [--by-grid]

--last-step N keeps only the last N launches (one bench step of N launches).
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r'\(.*$', '', name)
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = re.sub(r'at::native::', '', name)
    return name[:90]


def main():
    path = sys.argv[1]
    last = None
    by_grid = '--by-grid' in sys.argv
    if '--last-step' in sys.argv:
        last = int(sys.argv[sys.argv.index('--last-step') + 1])
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = v * {'ns': 1, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6, 'nsecond': 1, 's': 1e9}.get(unit, 1)
        rows.append((short(r['Kernel Name']), r.get('Grid Size', ''), r.get('Block Size', ''), ns))
    if last:
        rows = rows[-last:]
    agg = defaultdict(lambda: [0, 0.0])
    for name, grid, block, ns in rows:
        key = (name, grid) if by_grid else (name,)
        agg[key][0] += 1
        agg[key][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f'launches {len(rows)}  total {total / 1e6:.3f} ms (serialised, cold-cache: compare shares)')
    print(f'{"kernel":90s} {"grid":>18s} {"n":>6s} {"total ms":>10s} {"mean us":>9s} {"share":>7s}')
    for key, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        grid = key[1] if by_grid else ''
        print(f'{key[0]:90s} {grid:>18s} {n:6d} {ns / 1e6:10.3f} {ns / n / 1e3:9.1f} {ns / total * 100:6.1f}%')


if __name__ == '__main__':
    main()
