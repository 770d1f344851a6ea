#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/mgpu_fft_check.py > gpurun_out/j_fft$N.log 2>&1
echo "rc=$?" >> gpurun_out/j_fft$N.log
grep -E "rank|rc=|Error|error" gpurun_out/j_fft$N.log | head -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/mgpu_check.py 512 > gpurun_out/j_mgpu$N.log 2>&1
echo "rc=$?" >> gpurun_out/j_mgpu$N.log
tail -12 gpurun_out/j_mgpu$N.log
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hou_li or dealias" 2>&1 | tail -3 )
