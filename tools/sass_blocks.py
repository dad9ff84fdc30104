"""Group an ncu SASS source page (csv) into basic-block-like runs with equal execution counts: share of warp instructions,
average active lanes and opcode mix per run.  usage: sass_blocks.py page.csv [min_share_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'][0]
print(rows[start][1])
hdr = rows[start + 1]; ie_i = hdr.index('Instructions Executed'); te_i = hdr.index('Thread Instructions Executed')
body = [r for r in rows[start + 2:] if len(r) > te_i]
def f(x):
    try: return float(x)
    except ValueError: return 0.0
tot = sum(f(r[ie_i]) for r in body)
print('total warp inst %.4g, thread inst per warp inst %.2f' % (tot, sum(f(r[te_i]) for r in body) / tot))
blocks = []
for i, r in enumerate(body):
    ie, te = f(r[ie_i]), f(r[te_i])
    if ie == 0: continue
    lanes = te / ie
    if blocks and blocks[-1]['ie'] == ie and abs(blocks[-1]['lanes'] - lanes) < 0.05 and i == blocks[-1]['end'] + 1:
        blocks[-1]['end'] = i; blocks[-1]['ins'].append(r[1].strip())
    else:
        blocks.append(dict(start=i, end=i, ie=ie, lanes=lanes, ins=[r[1].strip()]))
for b in blocks:
    share = b['ie'] * len(b['ins']) / tot * 100
    if share < minshare: continue
    ops = {}
    for x in b['ins']:
        t = x.split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    print(f"[{b['start']:4d}-{b['end']:4d}] n={len(b['ins']):3d} exec {b['ie'] / tot * 100:.3f}% lanes {b['lanes']:5.1f} share {share:5.1f}%  " +
          ' '.join(f"{k}{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:9]))
