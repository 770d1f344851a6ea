#!/bin/bash
# round-2 GPU call A: parity tests with the warp-per-transform fused z kernel, A/B timing, ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/a_pytest.log
for v in old new pf; do
  echo "=== variant $v" >> gpurun_out/a_time.log
  if [ $v = old ]; then NSB200_ZF=old timeout 300 python scripts/quick_time.py 512 >> gpurun_out/a_time.log 2>&1
  elif [ $v = new ]; then timeout 300 python scripts/quick_time.py 512 >> gpurun_out/a_time.log 2>&1
  else NSB200_LIB=$PWD/3d_navier_stokes_b200/libnsb200_$v.so timeout 300 python scripts/quick_time.py 512 >> gpurun_out/a_time.log 2>&1; fi
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_z_fused -c 2 -o gpurun_out/a_zfw python scripts/profile_target.py 512 step 1 > gpurun_out/a_ncu.log 2>&1
tail -5 gpurun_out/a_pytest.log; cat gpurun_out/a_time.log
