#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/x.log
: > $L
for zf in warp default; do
  echo "== NSB200_ZF=$zf" >> $L
  NSB200_ZF=$zf timeout 300 python scripts/quick_time.py 1024 2>&1 | grep -E "^N=|z_fused|RK4|per step|rror" >> $L
done
( timeout 900 python -m pytest tests/test_gpu_variants.py -x -q -m gpu -k "1024" 2>&1 | tail -3 ) >> $L
cat $L
