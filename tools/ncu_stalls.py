#!/usr/bin/env python
"""One line per distinct (kernel, grid) of an `ncu --set full --csv --page raw` log: duration, DRAM
and issue utilisation, occupancy, registers, cache hit rates and the per-issue stall reasons.
    python tools/ncu_stalls.py gpurun_out/head_raw.csv"""
import csv
import re
import sys

lines = [ln for ln in open(sys.argv[1], newline='') if ln.startswith('"')]
rd = list(csv.reader(lines))
hdr, units = rd[0], rd[1]
col = {h: i for i, h in enumerate(hdr)}
ST = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'
stalls = ['long_scoreboard', 'lg_throttle', 'math_pipe_throttle', 'mio_throttle', 'short_scoreboard', 'barrier',
          'wait', 'not_selected', 'dispatch_stall', 'membar', 'drain', 'tex_throttle', 'branch_resolving', 'no_instruction', 'imc_miss', 'sleeping']
seen = set()
print(f'{"kernel":46s} {"grid":>7s} {"us":>8s} {"dram%":>6s} {"issue%":>6s} {"occ%":>5s} {"regs":>4s} {"L1hit":>5s} {"L2hit":>5s} {"Minst":>7s}  stalls/issue (top 4)')
for r in rd[2:]:
    name = re.sub(r'\(.*$', '', r[col['Kernel Name']])
    name = re.sub(r'^void\s+|\(anonymous namespace\)::|<unnamed>::', '', name)[:46]
    key = (name, r[col['Grid Size']])
    if key in seen:
        continue
    seen.add(key)

    def g(n, d=0.0):
        i = col.get(n)
        try:
            return float(r[i].replace(',', '')) if i is not None and r[i] != '' else d
        except ValueError:
            return d
    tu = units[col['gpu__time_duration.sum']]
    t = g('gpu__time_duration.sum') * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(tu, 1.0)
    ss = sorted(((g(ST % s), s) for s in stalls), reverse=True)[:4]
    grid = key[1].replace(' ', '')
    print(f'{name:46s} {grid:>7s} {t:8.1f} {g("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed"):6.1f} '
          f'{g("sm__issue_active.avg.pct_of_peak_sustained_elapsed"):6.1f} {g("sm__warps_active.avg.pct_of_peak_sustained_active"):5.1f} '
          f'{g("launch__registers_per_thread"):4.0f} {g("l1tex__t_sector_hit_rate.pct"):5.1f} '
          f'{g("LTS.TriageCompute.lts__average_t_sector_hit_rate_realtime.pct"):5.1f} {g("smsp__inst_executed.sum") / 1e6:7.2f}  '
          + ' '.join(f'{s}={v:.1f}' for v, s in ss))
