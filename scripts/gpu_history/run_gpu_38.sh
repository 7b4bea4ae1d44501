#!/bin/bash
timeout 300 python scripts/debug_fp4.py 2>&1 | tail -16
echo "#### poison"
PN2_POISON=1 timeout 300 python scripts/debug_fp4.py 2>&1 | tail -16
