#!/bin/bash
# ncu evidence for profiles/: launch list (shares) + full captures of the dominant kernels. Run via gpurun.
TAG=${1:-r1}
export S=64 NF=14
python scripts/profile_driver.py gen
# every launch of two steady-state frames (one published, one not): skip the first 10 frames' launches
ncu --metrics gpu__time_duration.sum --clock-control none -s 620 -c 140 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python scripts/profile_driver.py run > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launches_${TAG}.log
for K in lk_kernel be_feature_kernel mineig_kernel blur7_kernel be_gemm_kernel be_qr_kernel clahe_apply_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:^$K -s 8 -c 1 -o gpurun_out/prof_${K}_${TAG} -f \
      python scripts/profile_driver.py run > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_${K}_${TAG}.log
done
ls -la gpurun_out/ | grep -E "ncu-rep|launches"
