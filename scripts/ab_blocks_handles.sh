#!/bin/bash
# block shape x handles A/B of the h / nh kernel: every library under build_variants/ (python -m dynamicbayesiannetworks_b200.build --slim
# -DDBN_FAST_BLOCK=b -DDBN_FAST_MINB=m --out build_variants/fast_bxm.so) plus the default, bench.py with --handles 1 and 2
for v in "" $(ls build_variants/*.so 2>/dev/null); do for H in 1 2; do
  echo "== lib=${v:-default} handles=$H"
  DBN_B200_LIB=${v:+$PWD/$v} python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-c5 --handles $H 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('iters/s %.4g  ms/step %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done; done
