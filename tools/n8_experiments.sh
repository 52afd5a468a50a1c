T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
python bench.py --impl reference --workload cg --npts 8193 > gpurun_out/r2t_ref_cg_8193.json 2> gpurun_out/r2t_ref_cg.err
$T 29601 bench.py --gpus 8 --workload cg --npts 8193 --steps 5 --warmup 3 > gpurun_out/r2t_cg_n8.json 2> gpurun_out/r2t_cg_n8.err
$T 29602 tools/strips_opbench.py > gpurun_out/r2t_strips_opbench_n8.txt 2> gpurun_out/r2t_strips_opbench_n8.err
$T 29603 bench.py --gpus 8 --steps 20 --warmup 3 --profile > gpurun_out/r2t_bench_n8_prof.json 2> gpurun_out/r2t_bench_n8_prof.err
MGB_AGGLOMERATE=1023 $T 29604 bench.py --gpus 8 --steps 20 --warmup 3 --profile > gpurun_out/r2t_bench_n8_a1023.json 2> gpurun_out/r2t_bench_n8_a1023.err
tail -c 600 gpurun_out/r2t_cg_n8.json; echo; cat gpurun_out/r2t_strips_opbench_n8.txt
