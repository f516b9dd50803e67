#!/bin/bash
# A/B of prebuilt library variants (build/variants/*.so) on the C5 (segway) filter kernel and the other configs
cp asif_b200/libasif_b200.so /tmp/orig.so
python scripts/profile_c5.py
for v in "$@"; do
  cp build/variants/$v.so asif_b200/libasif_b200.so
  echo "== $v"; python scripts/profile_c5.py 2>&1 | tail -2
  python scripts/bench_configs.py 2>&1 | cut -c1-200
  python -m pytest tests -m gpu -q 2>&1 | tail -15
done
cp /tmp/orig.so asif_b200/libasif_b200.so
