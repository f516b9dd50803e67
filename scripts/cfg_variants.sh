#!/bin/bash
# A/B of prebuilt library variants (build/variants/*.so) over scripts/bench_configs.py; $1 = grep pattern of the config line
pat=$1; shift
cp asif_b200/libasif_b200.so /tmp/orig.so
for v in "$@"; do
  cp build/variants/$v.so asif_b200/libasif_b200.so
  echo "== $v"; python scripts/bench_configs.py 2>&1 | grep "$pat" | cut -c1-200
done
cp /tmp/orig.so asif_b200/libasif_b200.so
