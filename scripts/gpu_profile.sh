#!/bin/bash
# ncu evidence for profiles/: launch list (shares) + full captures of the dominant kernels. Run via gpurun.
# Frame 66 of a 70-frame replay is a publishing frame with a full 30-pose window (QR + both EKF updates active).
TAG=${1:-r1}
export S=64 NF=70
python scripts/profile_driver.py gen
LPF=$(python - <<'PY'
# launches per frame, counted from a dry run with the library's own counter
import subprocess, re, os
out = subprocess.run(["python", "scripts/profile_driver.py", "run"], capture_output=True, text=True, env=dict(os.environ, NF="4")).stdout
m = re.search(r"(\d+) launches", out)
print((int(m.group(1)) - 3) // 4 if m else 75)
PY
)
echo "launches per frame: $LPF"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((LPF * 66 + 3)) -c $((LPF * 2)) --csv --log-file gpurun_out/launches_${TAG}.csv \
    python scripts/profile_driver.py run > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launches_${TAG}.log
# kernel : launches per frame
for KL in lk_kernel:2 be_feature_kernel:2 mineig_kernel:1 blur7_kernel:1 be_gemm_kernel:9 be_qr_kernel:3 clahe_apply_kernel:1 be_chol_kernel:3 ransac_kernel:1 orb_kernel:3; do
  K=${KL%%:*}; L=${KL##*:}
  ncu --set full --clock-control none --import-source on -k regex:^$K -s $((L * 66)) -c $L -o gpurun_out/prof_${K}_${TAG} -f \
      python scripts/profile_driver.py run > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_${K}_${TAG}.log
done
ls -la gpurun_out/ | grep -E "ncu-rep|launches"
