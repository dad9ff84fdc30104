#!/bin/bash
# run bench.py once per tuning variant in build_variants/ and print the stage breakdown
for f in build_variants/libvrestir_*.so; do
  n=$(basename $f .so); n=${n#libvrestir_}
  VRESTIR_LIB=$PWD/$f python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$n', round(d['value'],3), 'serial', d['config'].get('pipelining',{}).get('unpipelined_ms_per_frame'), d['config']['stage_ms'])"
done
