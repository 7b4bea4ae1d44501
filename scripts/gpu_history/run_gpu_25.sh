#!/bin/bash
timeout 300 python scripts/debug_tc_trace.py 2>&1 | tail -40
