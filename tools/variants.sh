#!/bin/bash
# tuning helper: time the fast kernel under different compile variants (not part of the product)
for mb in 3 2; do
  PDB200_FAST_MINB=$mb python bench.py --steps 100 --no-cpu > /tmp/v_$mb.json
  python - <<PY
import json
d=json.load(open("/tmp/v_$mb.json"))
print("minb", $mb, "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
