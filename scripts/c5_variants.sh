#!/bin/bash
# A/B of prebuilt library variants on the C5 (segway) filter kernel
cp asif_b200/libasif_b200.so /tmp/orig.so
python scripts/profile_c5.py
for v in "$@"; do
  cp build/variants/$v.so asif_b200/libasif_b200.so
  echo "== $v"; python scripts/profile_c5.py 2>&1 | tail -2
done
cp /tmp/orig.so asif_b200/libasif_b200.so
