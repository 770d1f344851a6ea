#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity_large.py tests/test_gpu_variants.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/e_pytest.log
echo "== default (ring)" > gpurun_out/e_time.log
timeout 300 python scripts/quick_time.py 512 >> gpurun_out/e_time.log 2>&1
echo "== NSB200_RING=0" >> gpurun_out/e_time.log
NSB200_RING=0 timeout 300 python scripts/quick_time.py 512 >> gpurun_out/e_time.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fft_strided_ring -s 2 -c 2 -o gpurun_out/e_ring python scripts/profile_target.py 512 step 1 > gpurun_out/e_ncu.log 2>&1
cat gpurun_out/e_pytest.log gpurun_out/e_time.log
