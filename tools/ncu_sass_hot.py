"""Summarise an `ncu --page source --csv --print-source sass` dump: for the first kernel in the
file, print the instructions with the most executions / stall samples and totals per opcode.
Usage: python tools/ncu_sass_hot.py dump.csv [kernel_index] [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# split into kernels
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
b = blocks[ki]
hdr = b['rows'][0]
data = [dict(zip(hdr, r)) for r in b['rows'][1:] if len(r) >= len(hdr) - 2]
def num(x):
    try: return float(x)
    except Exception: return 0.0
tot_exec = sum(num(d['Instructions Executed']) for d in data)
tot_samp = sum(num(d['# Samples']) for d in data)
print(b['name'][:80], ' instrs executed', tot_exec, ' samples', tot_samp, ' sass lines', len(data))
byop = collections.Counter(); sop = collections.Counter()
for d in data:
    op = d['Source'].split()[0] if d['Source'] else '?'
    if op.startswith('@'):
        op = d['Source'].split()[1]
    op = op.split('.')[0]
    byop[op] += num(d['Instructions Executed']); sop[op] += num(d['# Samples'])
print('--- by opcode (executed share, sample share)')
for op, n in byop.most_common(22):
    print(f'{op:12s} {100*n/tot_exec:6.2f}%  {100*sop[op]/max(tot_samp,1):6.2f}%')
print('--- top stall-sample instructions')
stall_keys = [k for k in hdr if k.startswith('stall_')]
for d in sorted(data, key=lambda d: -num(d['# Samples']))[:top]:
    st = sorted(((num(d[k]), k[6:]) for k in stall_keys), reverse=True)[:2]
    print(f"{d['Address'][-5:]} {100*num(d['# Samples'])/tot_samp:5.2f}% exec {num(d['Instructions Executed']):9.0f}  {d['Source'][:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
