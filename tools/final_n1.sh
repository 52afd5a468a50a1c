#!/bin/bash
# The round-end evidence run on ONE GPU: tests, smoke, both bench arms (reference first), CG lines with their reference
# runs, the ncu launch list of the bench command, the per-level ncu summaries.
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
python bench.py --impl reference --workload cg --npts 4097 > gpurun_out/r2z_cg_ref_4097.json 2>/dev/null
python bench.py --workload cg --npts 4097 --steps 5 --warmup 3 > gpurun_out/r2z_cg_4097.json 2> gpurun_out/r2z_cg_4097.err
python bench.py --impl reference --workload cg --npts 8193 > gpurun_out/r2z_cg_ref_8193.json 2>/dev/null
python bench.py --workload cg --npts 8193 --steps 5 --warmup 3 > gpurun_out/r2z_cg_8193.json 2> gpurun_out/r2z_cg_8193.err
python bench.py --workload weak --steps 20 --warmup 3 > gpurun_out/r2z_weak_n1.json 2> gpurun_out/r2z_weak_n1.err
# ncu launch list of the bench command (after it exited 0 without ncu above)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r2z_ncu_bench.log 2>&1
bash tools/ncu_levels.sh > gpurun_out/r2z_ncu_levels.log 2>&1
cat gpurun_out/r2z_tests.log gpurun_out/r2z_smoke.log
