import importlib, sys, os
sys.path.insert(0, "/root/repo")
mgb = importlib.import_module("multigrid-petsc_b200")
for n, L in ((8193, 13), (1025, 10), (4097, 12)):
    opts = f"-npts {n} -mesh 0 -iter 40 -grids {L} -levels {L} -cycle 0 -map 2 -v 3,3 -moreNorm 0 -mgb_csr 0 -pc_type jacobi -ksp_richardson_scale 0.8 -rtol 1e-300"
    mgb.run_poisson(opts, want_u=False)
    best = 1e9
    for k in range(3):
        r = mgb.run_poisson(opts, want_u=False)
        best = min(best, r["solve_seconds"] / r["num_iter"])
    print(os.environ.get("MGB_BOTTOM_ROWS", "63"), n, "us/cycle %.1f" % (best * 1e6), flush=True)
