#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/h_bench$N.json 2> gpurun_out/h_bench$N.err
tail -12 gpurun_out/h_bench$N.err; head -c 5000 gpurun_out/h_bench$N.json
