#!/bin/bash
for d in 0 1 2 4 6 7; do PN2_DBG_WGRAD=$d timeout 100 python scripts/debug_wgrad_time.py 2>&1 | tail -1; done
