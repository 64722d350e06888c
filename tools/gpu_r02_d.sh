#!/bin/bash
# Hygiene round: compute-sanitizer (memcheck / racecheck / synccheck) over the self-tests and the fused-linear cases.
mkdir -p gpurun_out
bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1
grep -E "^===|ERROR SUMMARY|rc=|passed|failed|hazard" gpurun_out/sanitize.log | head -40
