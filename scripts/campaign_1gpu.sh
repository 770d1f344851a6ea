#!/bin/bash
# measurement campaign, one GPU (gpurun -- bash scripts/campaign_1gpu.sh): bench line, ncu launch list of the same command, full captures of the top kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --no-sweep > gpurun_out/r02_ncu_launches.log 2>&1
for k in k_z_fused_w k_fft_strided_ring_fr k_rk_stage; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/r02_$k python scripts/profile_target.py 512 step 2 > gpurun_out/r02_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_z_c2r_w -c 1 -o gpurun_out/r02_k_z_c2r_w python scripts/profile_target.py 512 z 1 > gpurun_out/r02_ncu_zc2r.log 2>&1
tail -3 gpurun_out/r02_bench_1gpu.err; ls -la gpurun_out/r02_*
