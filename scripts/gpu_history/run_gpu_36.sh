#!/bin/bash
timeout 600 python scripts/debug_model.py 2>&1 | grep -v "post-act" | tail -24
