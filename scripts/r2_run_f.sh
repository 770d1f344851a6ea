#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fft_strided_ring -s 2 -c 2 -o gpurun_out/f_ring python scripts/profile_target.py 512 step 1 > gpurun_out/f_ncu.log 2>&1
tail -3 gpurun_out/f_ncu.log
