#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/mgpu_tune.py 512 > gpurun_out/l_tune$N.log 2>&1
grep "ms/step" gpurun_out/l_tune$N.log; tail -3 gpurun_out/l_tune$N.log | grep -v ms/step
