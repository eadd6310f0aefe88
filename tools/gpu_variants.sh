#!/bin/bash
# time every tools/variants/lib_*.so on the A/B script
for f in tools/variants/lib_*.so; do
  echo "== $f"
  FRISK_B200_LIB=$PWD/$f python tools/direct_vs_bucket.py 2>&1 | grep -v default | tail -12
done
