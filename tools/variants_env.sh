#!/bin/bash
# developer helper: per-kernel times of one config-2 batch under several environment settings ("A=1 B=2" per argument)
for envs in "$@"; do
  echo "== $envs"
  env $envs GQ_OPTIONS=overlap_classify=0 GQ_PROFILE_ITERS=4 python tools/profile_run.py 2>&1 | tail -2
done
