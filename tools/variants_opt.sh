#!/bin/bash
# developer helper: per-kernel times of one config-2 batch under several settings; each argument is
# "ENV=.. ENV=.. -- opt=val,opt=val" (GQ_OPTIONS after the --)
for spec in "$@"; do
  envs="${spec%%--*}"; opts="${spec##*--}"
  echo "== $spec"
  env $envs GQ_OPTIONS="$(echo $opts | tr -d ' ')" GQ_PROFILE_ITERS=4 python tools/profile_run.py 2>&1 | tail -2
done
