#!/bin/bash
# Solver walltime (the reference's own MPI_Wtime bracket) of the same poisson.in through the three front ends on one box:
# the reference's solver.c over the engine's PETSc surface (GPU, general CSR), the specialised engine (GPU), the reference
# over the CPU mini-PETSc (checker).  usage: tools/refsolver_timing.sh NPTS LEVELS
N=${1:-1025}; L=${2:-10}
D=$(mktemp -d)
printf -- "-npts $N\n-mesh 0\n-iter 1000\n-grids $L\n-levels $L\n-cycle 0\n-map 2\n-v 3,3\n-moreNorm 0\n-pc_type jacobi\n-ksp_richardson_scale 0.8\n-mgb_csr 0\n" > $D/poisson.in
for exe in multigrid-petsc_b200/lib/poisson_petsc_b200 multigrid-petsc_b200/lib/poisson_dropin oracle/_ref/poisson_ref; do
  mkdir -p $D/run; cp $D/poisson.in $D/run/; for f in uData.dat XgridData.dat YgridData.dat; do ln -sf /dev/null $D/run/$f; done
  ( cd $D/run; $OLDPWD/$exe 2>&1 | grep -E "Number of iterations|Solver walltime" | tr '\n' ' ' ); echo " <- $exe ($N^2, $L levels)"
  rm -rf $D/run
done
