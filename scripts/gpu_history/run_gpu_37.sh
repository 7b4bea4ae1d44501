#!/bin/bash
for cfg in "PN2_TC_TMA=3 PN2_POISON=1" "PN2_TC_TMA=0" "PN2_TC_TMA=1" "PN2_TC_TMA=2" "PN2_GEMM_MODE=0"; do
  echo "#### $cfg"
  env $cfg timeout 300 python scripts/debug_model.py 2>&1 | grep -v "post-act" | tail -8
done
