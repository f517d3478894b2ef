#!/bin/bash
# ncu evidence for profiles/: launch list (shares) + one full capture of the dominant kernels. Run via gpurun.
set -x
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 6 --warmup 4 --profile-steps 0 --cpu-frames 5 > gpurun_out/ncu_bench_${TAG}.log 2>&1
for K in lk_kernel be_feature_kernel mineig_kernel blur7_kernel be_gemm_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -o gpurun_out/prof_${K}_${TAG} -f \
      python bench.py --steps 4 --warmup 4 --profile-steps 0 --cpu-frames 5 --seqs 64 > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out/
