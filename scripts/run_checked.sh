#!/bin/bash
# Builds the checked variant of the library (-DRSAC_CHECKED: device-side assertions on list / ring / index arithmetic, the
# scoring ring poison-filled before every bulk copy) and runs the GPU test-suite against it.  compute-sanitizer is closed on
# this GPU pool; this is the substitute, run once per round.  Usage (GPU box): bash scripts/run_checked.sh
set -e
cd "$(dirname "$0")/.."
if [ ! -f orb-slam2-optimized_b200/libransac_b200_checked.so ] || [ "$1" = "--rebuild" ]; then
    RSAC_LIB_OUT=orb-slam2-optimized_b200/libransac_b200_checked.so RSAC_EXTRA_NVCC="-DRSAC_CHECKED" python orb-slam2-optimized_b200/build.py
fi
RSAC_LIB=$PWD/orb-slam2-optimized_b200/libransac_b200_checked.so python -m pytest tests -m gpu -x -q
