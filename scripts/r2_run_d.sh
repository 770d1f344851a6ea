#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "windowed or upload or golden_one" 2>&1 | tail -8 ) > gpurun_out/d_pytest.log
echo "== default" > gpurun_out/d_time.log
timeout 300 python scripts/quick_time.py 512 >> gpurun_out/d_time.log 2>&1
echo "== NSB200_PIPE=1" >> gpurun_out/d_time.log
NSB200_PIPE=1 timeout 300 python scripts/quick_time.py 512 >> gpurun_out/d_time.log 2>&1
cat gpurun_out/d_pytest.log gpurun_out/d_time.log
