#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/quick_time.py 512 > gpurun_out/c_time.log 2>&1
timeout 300 python scripts/quick_time.py 256 >> gpurun_out/c_time.log 2>&1
timeout 300 python scripts/quick_time.py 1024 >> gpurun_out/c_time.log 2>&1
cat gpurun_out/c_time.log
