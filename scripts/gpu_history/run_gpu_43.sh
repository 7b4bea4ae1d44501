#!/bin/bash
for cfg in "PN2_TC_TMA=1 PN2_TC_FIX=4" "PN2_TC_TMA=1 PN2_TC_FIX=8" "PN2_TC_TMA=1 PN2_TC_FIX=0 PN2_TC_RAW=3" "PN2_TC_TMA=1 PN2_TC_FIX=0 PN2_TC_RAW=2" "PN2_TC_TMA=1 PN2_TC_FIX=0 PN2_TC_STREAM_B=1"; do
  echo "#### $cfg"
  env $cfg STRESS_ITERS=600 timeout 300 python scripts/stress_tc2.py 2>&1 | tail -9 | cut -c1-150
done
