#!/bin/bash
for f in learningorchestra_b200/lib/variants/libloexec_*.so; do
  echo "== $f"
  LOEXEC_LIB=$PWD/$f python scripts/kbench.py ${KB_ARGS:-} 2>&1 | grep -E "${KB_GREP:-.}"
done
