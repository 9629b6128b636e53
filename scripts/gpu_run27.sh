#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "max_difference" 2>&1 | grep -E "assert|Error|error|^E " | head -20
