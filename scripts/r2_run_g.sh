#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python bench.py --steps 20 --warmup 3 ) > gpurun_out/g_bench1.json 2> gpurun_out/g_bench1.err
tail -5 gpurun_out/g_bench1.err; head -c 6000 gpurun_out/g_bench1.json
