#!/bin/bash
# per-kernel ncu times of one steady-state frame for every tuning variant in build_variants/
for f in build_variants/libvrestir_*.so; do
  n=$(basename $f .so); n=${n#libvrestir_}
  VRESTIR_LIB=$PWD/$f ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 21 --csv --log-file gpurun_out/tune_$n.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - "$n" <<'PY'
import csv, sys
n = sys.argv[1]
rows = list(csv.reader(l for l in open(f'gpurun_out/tune_{n}.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
agg = {}
for r in rows[1:]:
    k = r[ki].split('(')[0].replace('void vrd::', '').replace('vrd::', '')
    agg[k] = agg.get(k, 0.0) + float(r[vi].replace(',', '')) / 1e6
print(n, ' '.join(f"{k}={v:.3f}" for k, v in sorted(agg.items())), 'total=%.3f' % sum(agg.values()))
PY
done
