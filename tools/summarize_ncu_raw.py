#!/usr/bin/env python
"""Per-kernel summary of an `ncu --csv --page raw` log: launches, mean duration, DRAM bytes and
throughput, achieved occupancy.   python tools/summarize_ncu_raw.py gpurun_out/elem_raw.csv"""
import csv
import re
import sys
from collections import defaultdict

lines = [ln for ln in open(sys.argv[1], newline='') if not ln.startswith('==')]
rd = list(csv.reader(lines))
hdr, units = rd[0], rd[1]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] == '':
        return default
    try:
        return float(r[i].replace(',', ''))
    except ValueError:
        return default


def unit_scale(name, table):
    u = units[col[name]] if name in col else ''
    return table.get(u, 1.0)


T = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6, 's': 1e6}
BY = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
agg = defaultdict(list)
for r in rd[2:]:
    name = re.sub(r'\(.*$', '', r[col['Kernel Name']])
    name = re.sub(r'^void\s+|\(anonymous namespace\)::|<unnamed>::', '', name)[:60]
    grid = r[col['Grid Size']] if 'Grid Size' in col else ''
    dur = get(r, 'gpu__time_duration.sum') * unit_scale('gpu__time_duration.sum', T)
    rd_b = get(r, 'dram__bytes_read.sum') * unit_scale('dram__bytes_read.sum', BY)
    wr_b = get(r, 'dram__bytes_write.sum') * unit_scale('dram__bytes_write.sum', BY)
    occ = get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')
    dpct = get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')
    regs = get(r, 'launch__registers_per_thread')
    agg[name].append((dur, rd_b, wr_b, occ, dpct, regs, grid))
print(f'{"kernel":60s} {"n":>4s} {"sum us":>9s} {"mean us":>8s} {"MB r":>8s} {"MB w":>8s} {"GB/s":>7s} {"dram%":>6s} {"occ%":>5s} {"regs":>5s}')
for name, v in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    n = len(v)
    tot = sum(x[0] for x in v)
    r_b = sum(x[1] for x in v)
    w_b = sum(x[2] for x in v)
    print(f'{name:60s} {n:4d} {tot:9.1f} {tot / n:8.1f} {r_b / n / 1e6:8.1f} {w_b / n / 1e6:8.1f} '
          f'{(r_b + w_b) / tot / 1e3 if tot else 0:7.0f} {sum(x[4] for x in v) / n:6.1f} {sum(x[3] for x in v) / n:5.1f} {v[0][5]:5.0f}')
