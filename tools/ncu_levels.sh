#!/bin/bash
# Per-level ncu evidence for the fused legs (levels 0-4 of the 8193^2 hierarchy) and the one-sweep smoother / residual
# kernels on the fine level.  Run on the GPU box:  bash tools/ncu_levels.sh   (text summaries land in gpurun_out/)
# Each command runs plain first (exit 0 required), then under ncu; the .ncu-rep files are summarised and deleted on the
# box (they are ~25 MB each, gpurun brings back at most 64 MiB).
set -e
OUT=gpurun_out/r2_ncu_levels.txt
: > $OUT
for lv in 0 1 2 3 4; do
  python tools/profile_ops.py --level $lv --ops fused_down,fused_up --reps 2 > gpurun_out/r2m_plain_l$lv.log 2>&1
done
python tools/profile_ops.py --level 0 --ops jacobi --reps 2 > gpurun_out/r2m_plain_jac.log 2>&1
python tools/profile_ops.py --level 0 --ops residual --reps 2 > gpurun_out/r2m_plain_res.log 2>&1
for lv in 0 1 2 3 4; do
  ncu --set full --clock-control none -k regex:k_jfused -s 3 -c 2 -o /tmp/r2m_fused_l$lv python tools/profile_ops.py --level $lv --ops fused_down,fused_up --reps 2 > gpurun_out/r2m_ncu_l$lv.log 2>&1 || true
  echo "===== level $lv: fused down leg (launch 1) and fused up leg (launch 2)" >> $OUT
  python tools/ncu_summary.py full /tmp/r2m_fused_l$lv.ncu-rep >> $OUT 2>&1 || true
done
for op in jacobi residual; do
  ncu --set full --clock-control none -k regex:k_stream5 -s 2 -c 1 -o /tmp/r2m_$op python tools/profile_ops.py --level 0 --ops $op --reps 2 > gpurun_out/r2m_ncu_$op.log 2>&1 || true
  echo "===== level 0: one-sweep $op kernel (k_stream5)" >> $OUT
  python tools/ncu_summary.py full /tmp/r2m_$op.ncu-rep >> $OUT 2>&1 || true
done
cat gpurun_out/r2m_plain_l*.log gpurun_out/r2m_plain_jac.log gpurun_out/r2m_plain_res.log > gpurun_out/r2_ncu_levels_plain_times.txt
wc -l $OUT
