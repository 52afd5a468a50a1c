/*
 * mg_oracle -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C on top of oracle/minipetsc) of the reference's hot
 * path: mesh + index maps + transfer stencils (src/mesh.c, src/matbuild.c,
 * src/problem.c), assembly of A/R/P/b (src/solver.c:185-253,489-620,1035-1209),
 * the cycle-0 V-cycle driver (src/solver.c:1414-1575), the cycle-8 PCMG driver
 * (src/solver.c:1884-1989) and post-processing (src/solver.c:1211-1380).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library -- as the CHECKER, never as the product.
 *
 * PARITY UNPINNED against real PETSc (see minipetsc/petscksp.h).  The driver
 * logic IS pinned: tests/test_oracle_vs_ref.py runs the reference's own
 * unmodified sources (compiled in place against minipetsc into oracle/_ref/)
 * and requires bit-identical rData/uData/eData.
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct MgoCtx MgoCtx;

/* options: PETSc-style string with the reference's keys (poisson.in: -npts -mesh -iter -grids
 * -levels -cycle -map -v a,b -moreNorm) plus any solver options minipetsc understands
 * (-pc_type, -ksp_richardson_scale, -pc_sor_*, -ksp_type, -ksp_rtol, -mg_levels_*, -mg_coarse_*).
 * Extensions (not in the reference): -map 3 = red-black numbering (reds (i+j even) first);
 * -rtol r = stopping tolerance of cycle 0 (reference hard-codes 1e-7, src/solver.c:1530);
 * -threads t = OpenMP threads for the Vec/Mat kernels.
 * Runs SetUpProblem .. Assemble.  Returns NULL on invalid input (message on stderr). */
MgoCtx *mgo_create(const char *options);
void    mgo_destroy(MgoCtx *c);

int mgo_levels(const MgoCtx *c);
int mgo_level_dims(const MgoCtx *c, int l, int *ni, int *nj);       /* interior rows (y) / cols (x) */
int mgo_level_h(const MgoCtx *c, int l, double h[2]);
int mgo_coords(const MgoCtx *c, int dim, double *out);              /* npts values */
int mgo_stencil(const MgoCtx *c, int which, double out[9]);         /* 0: restriction, 1: prolongation (3x3) */
int mgo_opA(const MgoCtx *c, int l, int i, int j, double As[5]);    /* coefficients (S,W,C,E,N) at grid point */
int mgo_grid_to_global(const MgoCtx *c, int l, int *out);           /* ni*nj */
int mgo_global_to_grid(const MgoCtx *c, int l, int *out);           /* 3 * ni*nj : (i,j,gridId) */

/* assembled matrices: which = 0: A[l], 1: res[l] (level l -> l+1), 2: pro[l] (level l+1 -> l) */
int mgo_csr_dims(const MgoCtx *c, int which, int l, int *m, int *n, int *nnz);
int mgo_csr_copy(const MgoCtx *c, int which, int l, int *ia, int *ja, double *va);

/* vectors in GLOBAL numbering: which = 0: b[l], 1: u[l] */
int mgo_vec_get(const MgoCtx *c, int which, int l, double *out);
int mgo_vec_set(MgoCtx *c, int which, int l, const double *in);

/* single kernels on caller vectors (global numbering) */
int mgo_matmult(const MgoCtx *c, int which, int l, const double *x, double *y);
int mgo_matmultadd(const MgoCtx *c, int which, int l, const double *x, const double *y, double *z);
int mgo_residual(const MgoCtx *c, int l, const double *b, const double *x, double *r);
/* the level-l smoother exactly as cycle 0 configures it (KSPRICHARDSON, KSP_NORM_NONE, max_it = nu,
 * options from the database): x is in/out */
int mgo_smooth(MgoCtx *c, int l, const double *b, double *x, int nu, int guess_zero);
double mgo_norm2(const double *x, int n);
double mgo_dot(const double *x, const double *y, int n);

/* Solve() : cycle 0 or 8 per -cycle.  Returns 0 on success. */
int    mgo_solve(MgoCtx *c);
int    mgo_num_iter(const MgoCtx *c);
int    mgo_rnorm(const MgoCtx *c, double *out, int nmax);           /* relative residual history, numIter+1 values */
double mgo_solve_seconds(const MgoCtx *c);                          /* wall time of the cycle loop only */
/* GetSol + GetError: u in GRID (natural row-major) order, error = {max, sum|e|, sqrt(sum e^2)} */
int    mgo_postprocess(const MgoCtx *c, double *u_grid, double error[3]);
/* writes uData.dat rData.dat eData.dat XgridData.dat YgridData.dat into dir (reference formats) */
int    mgo_write_files(const MgoCtx *c, const char *dir);

#ifdef __cplusplus
}
#endif
#endif
