#!/bin/bash
# compute-sanitizer over small end-to-end runs of every architecture (memcheck), racecheck on the DSD100 + stereo paths
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > $O/r2_memcheck.log 2>&1
tail -5 $O/r2_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > $O/r2_racecheck.log 2>&1
tail -5 $O/r2_racecheck.log
echo sanitize done
