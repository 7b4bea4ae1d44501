#!/bin/bash
timeout 600 python tests/debug_model.py 2>&1 | grep -v "post-act" | tail -24
