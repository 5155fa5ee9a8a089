#!/bin/bash
# developer helper: config-3 shape (1 M reads) under several settings; each argument is "ENV=.. ENV=.."
for envs in "$@"; do
  echo "== $envs"
  env $envs python tools/nested_config.py 200 5000 10 1000000 2000 2>&1 | grep -E "parity|kernels_ms" | tail -2
done
