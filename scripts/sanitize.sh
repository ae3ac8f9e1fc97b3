#!/bin/bash
# usage: bash scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [kernel groups...]
TOOL=${1:-memcheck}; shift
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 3 --print-limit 20 \
   python scripts/sanitizer_smoke.py "$@" > gpurun_out/sanitize_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer smoke|Error|hazard" gpurun_out/sanitize_$TOOL.log | head -30
