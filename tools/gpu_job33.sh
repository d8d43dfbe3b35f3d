#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 compute-sanitizer --tool racecheck dasr_b200/lib/selftest check > $O/r2_racecheck_selftest.log 2>&1; tail -3 $O/r2_racecheck_selftest.log; grep -c PASS $O/r2_racecheck_selftest.log
