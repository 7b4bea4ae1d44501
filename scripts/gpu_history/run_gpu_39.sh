#!/bin/bash
PN2_POISON=1 timeout 300 python scripts/debug_fp4.py 2>&1 | tail -22
