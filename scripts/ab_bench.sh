#!/bin/bash
# A/B the chain kernel variants built under build_variants/ (plus the default library): one short bench run each.
# usage (on the GPU box): bash scripts/ab_bench.sh [steps] [warmup]
S=${1:-50}; W=${2:-20}
for v in "" $(ls build_variants/*.so 2>/dev/null); do
  echo "== lib=${v:-default}"
  DBN_B200_LIB=${v:+$PWD/$v} python bench.py --steps $S --warmup $W --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('iters/s %.4g  ms/step %.3f  frac %.3f  e2e %.4g  acc %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('accept_rate')))"
done
