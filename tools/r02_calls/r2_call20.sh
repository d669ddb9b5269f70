#!/bin/bash
# GPU box, one GPU: ncu launch list of exactly one timed S1 step of the last build (derive kernels + 3 warm-up steps of
# 61 launches come first), after the same command has run without ncu.
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree"
$B > $out/plain_r2l.log 2>&1 || { tail -5 $out/plain_r2l.log; exit 1; }
tail -c 400 $out/plain_r2l.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -s 187 -c 61 --csv --log-file $out/launches_r2l.csv $B > $out/ncu_l_r2l.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_r2l.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); mi = h.index('Metric Value')
seq = [(r[ki].split('(')[0].replace('void ', '').replace('gb::', ''), float(r[mi].replace(',', '')) * 1e-6) for r in rows[hdr + 1:] if len(r) > mi]
print('first', seq[0][0], 'last', seq[-1][0], 'n', len(seq))
agg = collections.OrderedDict()
for k, v in seq:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in seq)
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:44]:44s} {c:3d} launches {v:8.3f} ms {100 * v / tot:5.1f} %')
print('total', round(tot, 3), 'ms')
PY
