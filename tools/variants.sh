#!/bin/bash
# developer helper: per-kernel times of one config-2 batch for several builds of the library (GQ_LIB override)
for lib in "$@"; do
  echo "== $lib"
  GQ_LIB=$lib GQ_OPTIONS=overlap_classify=0 GQ_PROFILE_ITERS=4 python tools/profile_run.py 2>&1 | tail -2
done
