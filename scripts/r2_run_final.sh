#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 ) > gpurun_out/gpu_tests.log 2>&1
tail -14 gpurun_out/gpu_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" ) 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
tail -4 gpurun_out/r02_bench_1gpu.err
