"""Summarise an `ncu -i X.ncu-rep --page source --csv --print-source sass` export: per-opcode executed warp-instructions
and stall samples of the first captured kernel, plus the hottest instructions.
usage: python scripts/ncu_sass_summary.py export.csv [top_n]"""
import csv, sys, collections, re
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
# first kernel only: rows[0] = kernel name, rows[1] = header
hdr = rows[1]
body = []
for r in rows[2:]:
    if len(r) < len(hdr) - 2: break      # next kernel's name row
    body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot_exec = 0; per_op = collections.Counter(); per_op_samples = collections.Counter(); stall_tot = collections.Counter()
recs = []
for r in body:
    src = r[ci['Source']]
    m = re.match(r'\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', src)
    op = m.group(1) if m else '?'
    ex = int(r[ci['Instructions Executed']] or 0)
    smp = int(r[ci['# Samples']] or 0)
    tot_exec += ex; per_op[op] += ex; per_op_samples[op] += smp
    for s in stalls:
        v = r[ci[s]]
        if v: stall_tot[s] += int(v)
    recs.append((smp, ex, r[ci['Address']], src, {s: int(r[ci[s]] or 0) for s in stalls}))
print(rows[0][1][:100])
print('warp-instructions executed:', tot_exec, ' samples:', sum(per_op_samples.values()))
print('stall samples:', ', '.join(f'{k[6:]} {v / max(1, sum(stall_tot.values())):.3f}' for k, v in stall_tot.most_common(10)))
print('per opcode (share of executed, share of samples):')
for op, ex in per_op.most_common(22):
    print(f'  {op:10s} {ex / tot_exec:6.3f} {per_op_samples[op] / max(1, sum(per_op_samples.values())):6.3f}')
print('hottest instructions by samples:')
for smp, ex, addr, src, st in sorted(recs, key=lambda t: -t[0])[:top]:
    top_st = max(st.items(), key=lambda kv: kv[1])
    print(f'  {smp:6d} ex={ex:8d} {src[:70]:70s} {top_st[0][6:]}={top_st[1]}')
