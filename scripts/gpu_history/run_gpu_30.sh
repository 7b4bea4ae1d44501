#!/bin/bash
timeout 200 ./scripts/micro/tma_stream 524288 128 2>&1 | tail -30
