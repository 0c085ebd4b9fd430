#!/bin/bash
# A/B of the units-per-warp spreading on the small BASELINE configs (c2: 320 units, c3: 2816 units): DBN_LANES = units per warp
for L in 32 16 8 4 2 1 auto; do
  echo "== DBN_LANES=$L"
  if [ "$L" = auto ]; then unset DBN_LANES; else export DBN_LANES=$L; fi
  python scripts/bench_methods.py "chains" 2>&1 | grep -E "^\{" | grep -v "fp-\|4096" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-40s %6d units  %.3g unit-it/s' % (d['case'], d['units'], d['unit_iterations_per_sec']))"
done
