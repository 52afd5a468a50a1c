"""Golden-vector cases shared by make_golden.py and the tests.

Every case is an option string in the reference's own poisson.in vocabulary.  The smoother is always
pinned explicitly (the shipped poisson.in leaves PETSc's default PC, ILU(0), in place).
"""

JAC = "-pc_type jacobi -ksp_richardson_scale 0.8"
MGJ = ("-mg_levels_ksp_type richardson -mg_levels_pc_type jacobi "
       "-mg_levels_ksp_richardson_scale 0.8 -mg_levels_ksp_max_it 3")
MGC = ("-mg_coarse_ksp_type richardson -mg_coarse_pc_type jacobi "
       "-mg_coarse_ksp_richardson_scale 0.8 -mg_coarse_ksp_max_it 3")


def base(npts, levels, cycle=0, mesh=0, it=100000, v="3,3"):
    return (f"-npts {npts} -mesh {mesh} -iter {it} -grids {levels} -levels {levels} "
            f"-cycle {cycle} -map 2 -v {v} -moreNorm 0")


# name -> (options, store_full_u)
CASES = {
    # the shipped poisson.in sizes (reference/poisson.in:1-14) with the smoother pinned
    "n17_l2_jacobi": (base(17, 2) + " " + JAC, True),
    "n17_l2_jacobi23": (base(17, 2) + " -pc_type jacobi -ksp_richardson_scale 0.6666666666666666", True),
    "n17_l2_sor": (base(17, 2) + " -pc_type sor", True),
    "n17_l2_ilu_default": (base(17, 2), True),
    "n17_l1_jacobi": (base(17, 1, it=50) + " " + JAC, True),
    # BASELINE.json configs[0]: 129x129, 4-level V-cycle
    "n129_l4_jacobi": (base(129, 4) + " " + JAC, True),
    "n129_l4_sor": (base(129, 4) + " -pc_type sor", True),
    "n129_l4_sor_forward": (base(129, 4) + " -pc_type sor -pc_sor_forward", False),
    "n129_l7_jacobi": (base(129, 7) + " " + JAC, True),
    "n129_l7_jacobi_v21": (base(129, 7, v="2,1") + " " + JAC, False),
    # non power-of-two grid: 1/h^2 inexact
    "n101_l3_jacobi": (base(101, 3) + " " + JAC, True),
    # non-uniform meshes
    "n65_l4_mesh1_jacobi": (base(65, 4, mesh=1, it=400) + " " + JAC, True),
    "n65_l4_mesh2_jacobi": (base(65, 4, mesh=2, it=400) + " " + JAC, True),
    # cycle 8: MG-preconditioned CG (BASELINE.json configs[2] at a CPU-friendly size)
    "n129_l7_cg_mg": (base(129, 7, cycle=8) + " -ksp_type cg -ksp_rtol 1e-10 " + MGJ, True),
    "n129_l4_cg_mg_jcoarse": (base(129, 4, cycle=8) + " -ksp_type cg -ksp_rtol 1e-10 " + MGJ + " " + MGC, True),
    "n129_l7_rich_mg_monitor": (base(129, 7, cycle=8, it=50) + " -ksp_monitor " + MGJ, False),
    # BASELINE.json configs[1]: 1025^2, 7 levels (u stored as SHA-256 only)
    "n1025_l7_jacobi": (base(1025, 7) + " " + JAC, False),
    "n1025_l10_jacobi": (base(1025, 10) + " " + JAC, False),
}

# red-black numbering (-map 3) does not exist in the reference: goldens for these come from the oracle
ORACLE_ONLY_CASES = {
    "n17_l2_rbsor": (base(17, 2).replace("-map 2", "-map 3") + " -pc_type sor", True),
    "n129_l4_rbsor": (base(129, 4).replace("-map 2", "-map 3") + " -pc_type sor", True),
    "n129_l7_rbsor": (base(129, 7).replace("-map 2", "-map 3") + " -pc_type sor", True),
    "n129_l7_rbsor_w12": (base(129, 7).replace("-map 2", "-map 3") + " -pc_type sor -pc_sor_omega 1.2", False),
    "n1025_l7_rbsor": (base(1025, 7).replace("-map 2", "-map 3") + " -pc_type sor", False),
}
