#!/bin/bash
# tuning helper: same-box A/B of library builds (PDB200_LIB), headline bench only.  tools/ab_libs.sh name1 name2 ... ("head" = the product build)
show='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "sustained", round(d["sustained"]["ms_per_step"],4), d["config"]["kernel"], d["clocks"]["sm_mhz"], d["sustained"]["clocks"]["sm_mhz"], flush=True)'
root=$(cd "$(dirname "$0")/.." && pwd)
for rep in ${REPS:-1 2}; do
for v in "$@"; do
  lib=$root/dune-pdelab_b200/lib/variants/libpdelab_b200_$v.so
  [ "$v" == "head" ] && lib=$root/dune-pdelab_b200/lib/libpdelab_b200.so
  PDB200_LIB=$lib timeout 300 python $root/bench.py --no-cpu --no-configs --no-e2e 2>/dev/null | python -c "$show" "$v"
done; done
