#!/bin/bash
for cfg in "PN2_TC_TMA=3 PN2_TC_FIX=0" "PN2_TC_TMA=3 PN2_TC_FIX=1" "PN2_TC_TMA=1 PN2_TC_FIX=0"; do
  echo "#### $cfg"
  env $cfg STRESS_ITERS=400 timeout 300 python scripts/stress_tc2.py 2>&1 | tail -24
done
