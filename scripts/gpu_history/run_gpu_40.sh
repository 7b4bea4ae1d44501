#!/bin/bash
for cfg in "PN2_TC_TMA=3" "PN2_TC_TMA=1" "PN2_TC_TMA=2" "PN2_TC_TMA=0"; do
  echo "#### $cfg"
  env $cfg STRESS_ITERS=60 timeout 300 python scripts/stress_tc.py 2>&1 | grep -v "^   iter" | tail -8
  env $cfg STRESS_ITERS=60 timeout 300 python scripts/stress_tc.py 2>&1 | grep "^   iter" | head -6
done
