#!/bin/bash
timeout 300 python scripts/debug_tc_trace.py 2>&1 | grep -v "chunks\|tiles" | head -70
