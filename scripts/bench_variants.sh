#!/bin/bash
# A/B of prebuilt library variants (build/variants/*.so): short device-resident bench for each
cp asif_b200/libasif_b200.so /tmp/orig.so
for v in "$@"; do
  cp build/variants/$v.so asif_b200/libasif_b200.so
  echo "== $v"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g kernel_ms %.4f frac %.3f'%(d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
    else: print(l.rstrip()[:200])
"
done
cp /tmp/orig.so asif_b200/libasif_b200.so
