#!/bin/bash
# tuning helper: same-box A/B of environment switches of the product build, headline bench only.
#   tools/ab_env.sh "" "PDB200_FAST_PERSIST=1" ...
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "sustained", round(d["sustained"]["ms_per_step"],4), d["config"]["kernel"], d["clocks"]["sm_mhz"], d["sustained"]["clocks"]["sm_mhz"], flush=True)'
root=$(cd "$(dirname "$0")/.." && pwd)
for rep in ${REPS:-1 2}; do
for v in "$@"; do
  env $v timeout 300 python $root/bench.py --no-cpu --no-configs --no-e2e 2>/dev/null | python -c "$show" "[$v]"
done; done
