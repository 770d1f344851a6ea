#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -x -q -m gpu -k "1024 or large_grid or rk4_steps or golden" 2>&1 | tail -5 ) > gpurun_out/k_pytest.log
timeout 600 python scripts/quick_time.py 1024 > gpurun_out/k_time.log 2>&1
echo "== NSB200_ZF=old" >> gpurun_out/k_time.log
NSB200_ZF=old timeout 600 python scripts/quick_time.py 1024 2>&1 | grep -E "z_c2r|z_fused|c2r\+r2c|RK4|per step" >> gpurun_out/k_time.log
timeout 300 python scripts/quick_time.py 512 2>&1 | grep -E "rk point|RK4|per step" >> gpurun_out/k_time.log
cat gpurun_out/k_pytest.log gpurun_out/k_time.log
