#!/bin/bash
for cfg in "PN2_TC_TMA=3 PN2_TC_FIX=0" "PN2_TC_TMA=3 PN2_TC_FIX=1" "PN2_TC_TMA=3 PN2_TC_FIX=2" "PN2_TC_TMA=3 PN2_TC_FIX=3" "PN2_TC_TMA=2 PN2_TC_FIX=0" "PN2_TC_TMA=2 PN2_TC_FIX=2"; do
  echo "#### $cfg"
  env $cfg STRESS_ITERS=120 timeout 300 python scripts/stress_tc.py > /tmp/st.log 2>&1
  grep -v "^   \|^     " /tmp/st.log | tail -5
  grep "^   \|^     " /tmp/st.log | head -12
done
