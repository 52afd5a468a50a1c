import importlib, time, sys
sys.path.insert(0, "/root/repo")
mgb = importlib.import_module("multigrid-petsc_b200")
for n, L in ((1025, 10), (2049, 11), (4097, 12)):
    opts = f"-npts {n} -mesh 0 -iter 1000 -grids {L} -levels {L} -cycle 0 -map 2 -v 3,3 -moreNorm 0 -mgb_csr 0 -pc_type jacobi -ksp_richardson_scale 0.8"
    for k in range(3):
        t0 = time.perf_counter(); r = mgb.run_poisson(opts, want_u=False); t1 = time.perf_counter()
        print(n, "run", k, "iters", r["num_iter"], "solve_s %.4f" % r["solve_seconds"], "total_s %.3f" % (t1 - t0), flush=True)
