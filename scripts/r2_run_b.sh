#!/bin/bash
# round-2 GPU call B: full-size parity tests, option-B z kernel timing, source-level ncu of the x inverse pass
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity_large.py -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/b_pytest.log
timeout 300 python scripts/quick_time.py 512 > gpurun_out/b_time.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fft_strided -s 2 -c 2 -o gpurun_out/b_strided python scripts/profile_target.py 512 step 1 > gpurun_out/b_ncu.log 2>&1
tail -5 gpurun_out/b_pytest.log; cat gpurun_out/b_time.log
