#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 ) > gpurun_out/gpu_tests.log 2>&1
tail -25 gpurun_out/gpu_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" ) 2>&1 | tail -3
