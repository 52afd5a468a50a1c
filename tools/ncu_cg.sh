#!/bin/bash
# ncu --set full of the CG kernels at 8193^2 (k_cg_pstep: direction update + operator apply + deferred x update; k_cg_update:
# r -= a w, ||r||) and of the persistent bottom kernel; the command runs plain first.
python bench.py --workload cg --npts 8193 --steps 1 --warmup 3 --profile --no-cpu-baseline > gpurun_out/r2n_cg_plain.json 2> gpurun_out/r2n_cg_plain.err || exit 1
OUT=gpurun_out/r2_ncu_cg.txt; : > $OUT
for k in k_cg_pstep k_cg_update k_coarse_cycle; do
  ncu --set full --clock-control none -k regex:$k -s 8 -c 1 -o /tmp/r2n_$k python bench.py --workload cg --npts 8193 --steps 1 --warmup 3 --profile --no-cpu-baseline > gpurun_out/r2n_ncu_$k.log 2>&1 || true
  echo "===== $k (8193^2, MG-preconditioned CG)" >> $OUT
  python tools/ncu_summary.py full /tmp/r2n_$k.ncu-rep >> $OUT 2>&1 || true
done
wc -l $OUT
