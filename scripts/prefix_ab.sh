#!/bin/bash
# prefix build A/B (device time, CUDA events) on c4 (and c5 size): tiled totals kernel vs the round-1 totals pass, scan vs in-pass sums
echo "== default (cp.async staging, tiled totals, 64-way scan)"; python scripts/prefix_tune.py 2>&1 | tail -1
echo "== DBN_PREFIX_NO_SCAN=1"; DBN_PREFIX_NO_SCAN=1 python scripts/prefix_tune.py 2>&1 | tail -1
echo "== DBN_PREFIX_OLD_TOTALS=1"; DBN_PREFIX_OLD_TOTALS=1 python scripts/prefix_tune.py 2>&1 | tail -1
echo "== c5 size"; python scripts/prefix_tune.py 500 10000 2>&1 | tail -1
