#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_r2.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python tools/sanitize_r2.py > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_racecheck.log | cut -c1-200
grep -c "RACECHECK\|Race reported" gpurun_out/r2_racecheck.log
