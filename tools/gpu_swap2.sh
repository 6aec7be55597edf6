#!/bin/bash
timeout -k 5 200 python -m pytest tests/test_kernels_gpu.py -q -k "role_swapped" 2>&1 | grep -E "^E|assert|passed|failed|Error" | head -40
