#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/debug_mask_fusion.py 2>&1 | tail -30
