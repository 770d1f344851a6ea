#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/i_pytest.log 2>&1
cat gpurun_out/i_pytest.log
