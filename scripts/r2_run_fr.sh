#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/fr3.log
: > $L
for lib in libnsb200_ah0.so libnsb200_ah1.so libnsb200.so libnsb200_ah3.so; do
  echo "== $lib" >> $L
  NSB200_LIB=$PWD/3d_navier_stokes_b200/$lib timeout 300 python scripts/quick_time.py 512 2>&1 | grep -E "pass_|c2r\+r2c|RK4|per step|rror" >> $L
done
for tpc in 4 6 12 16; do
  echo "== default lib TPC=$tpc" >> $L
  NSB200_RING_TPC=$tpc timeout 300 python scripts/quick_time.py 512 2>&1 | grep -E "pass_|RK4|per step|rror" >> $L
done
echo "== parity" >> $L
( timeout 900 python -m pytest tests/test_gpu_parity_large.py -x -q -m gpu 2>&1 | tail -3 ) >> $L
for tpc in 8 3 100; do
  echo "== stress TPC=$tpc" >> $L
  ( NSB200_RING_TPC=$tpc timeout 300 python scripts/ring_stress.py 512 20 2>&1 | tail -2 ) >> $L
done
cat $L
