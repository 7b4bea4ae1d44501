#!/bin/bash
for cfg in "PN2_TC_TMA=3 PN2_TC_FIX=0" "PN2_TC_TMA=3 PN2_TC_FIX=4" "PN2_TC_TMA=3 PN2_TC_FIX=8" "PN2_TC_TMA=3 PN2_TC_FIX=0" "PN2_TC_TMA=3 PN2_TC_FIX=4"; do
  echo "#### $cfg"
  env $cfg STRESS_ITERS=2400 timeout 400 python scripts/stress_tc2.py 2>&1 | grep "BAD ITER\|Error\|error" | cut -c1-150
done
