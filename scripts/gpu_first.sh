#!/bin/bash
# first GPU contact: parity tests through the C ABI
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60
