/*
 * mgb200.h -- C-ABI of the B200 geometric-multigrid Poisson engine (libmgb200.so).
 *
 * This is the boundary between host C (the replacement of the reference's src/solver.c, see
 * multigrid-petsc_b200/host/solver_b200.c and INTEGRATION.md) and the sm_100a CUDA kernels.
 * Plain pointers and sizes only; every function returns 0 on success and a negative MGB_E* code on
 * failure (mgb_last_error() gives the text).  Pointers are HOST pointers unless the name ends in _dev.
 *
 * What each entry point replaces in the reference ("ref:" = /root/reference):
 *
 *   mgb_create / mgb_destroy         SetUpSolver + SetUpAssembly / DestroySolver + DestroyAssembly
 *                                    ref: src/solver.c:33-61, 63-105, 107-149
 *   mgb_set_level_operator           the per-row OpA(metrics, h) evaluation of fillJacobians
 *                                    ref: src/solver.c:231-236 ; src/problem.c:3-22 ; src/mesh.c:29-107
 *   mgb_set_transfer                 op.res[0] / op.pro[0] 3x3 stencils, ref: src/matbuild.c:398-431
 *   mgb_assemble_csr / mgb_csr_*     levelMatrixA + fillJacobians + MatAssembly (A), Res, Pro
 *                                    ref: src/solver.c:185-253, 489-510, 1035-1154
 *   mgb_set_rhs*                     levelvecb, ref: src/solver.c:558-620
 *   mgb_op_apply                     MatMult(A[l], x, y)              ref: src/solver.c:1516
 *   mgb_op_residual                  KSPBuildResidual (r = b - A u)   ref: src/solver.c:1534,1545
 *   mgb_op_smooth                    KSPSolve on KSPRICHARDSON + PCJACOBI / PCSOR, KSP_NORM_NONE
 *                                    ref: src/solver.c:1463-1510 (configuration), :1531,1536,1542 (calls)
 *   mgb_op_restrict                  MatMult(res[l], r, b[l+1])       ref: src/solver.c:1535
 *   mgb_op_prolong                   MatMult(pro[l], u[l+1], rv[l]) ; VecAXPY(u[l], 1.0, rv[l])
 *                                    ref: src/solver.c:1540-1541
 *   mgb_op_norm2 / mgb_op_dot        VecNorm(NORM_2) / VecDot         ref: src/solver.c:1512,1518,1546
 *   mgb_solve_vcycle                 MultigridVcycle (cycle 0)        ref: src/solver.c:1414-1575
 *   mgb_solve_pcmg                   MultigridPetscPCMG (cycle 8)     ref: src/solver.c:1884-1989
 *   mgb_get_solution / mgb_error_norms   GetSol / GetError            ref: src/solver.c:1211-1315
 *
 * Grid conventions (ref: src/problem.c:6-8, src/matbuild.c:64-66, 296-303): level l (0 = finest) is an
 * ni x nj array of interior unknowns, i = grid row = y, j = grid column = x, unknown (i,j) has global
 * (natural) number i*nj + j.  All host vectors crossing this boundary are dense row-major ni*nj doubles
 * in that numbering, i.e. bit-compatible with the contents of the reference's PETSc Vec.
 */
#ifndef MGB200_H
#define MGB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgb_engine mgb_engine;

/* error codes */
#define MGB_OK            0
#define MGB_EINVAL       -1   /* bad argument / unsupported configuration */
#define MGB_ECUDA        -2   /* CUDA runtime error */
#define MGB_ENOMEM       -3
#define MGB_ESTATE       -4   /* call order violated (e.g. solve before operators are set) */
#define MGB_EDIVERGED    -5

/* which vector of a level */
#define MGB_VEC_B  0          /* right-hand side b[l]            ref: Assembly.b */
#define MGB_VEC_U  1          /* solution / correction u[l]      ref: Assembly.u */
#define MGB_VEC_R  2          /* residual work vector rv[l]      ref: src/solver.c:1459 */
#define MGB_VEC_W  3          /* scratch (Jacobi ping-pong / Krylov work) */
#define MGB_VEC_P  4          /* Krylov work (CG direction) */
#define MGB_VEC_Z  5          /* Krylov work (preconditioned residual) */
#define MGB_VEC_Q  6          /* Krylov work (A p) */
#define MGB_NVEC   7

/* which assembled matrix */
#define MGB_MAT_A    0        /* A[l]                        n_l^2 x n_l^2     */
#define MGB_MAT_RES  1        /* res[l] : level l -> l+1     n_{l+1}^2 x n_l^2 */
#define MGB_MAT_PRO  2        /* pro[l] : level l+1 -> l     n_l^2 x n_{l+1}^2 */

/* smoothers: the PC of the level KSPRICHARDSON (ref: src/solver.c:1463-1510 + KSPSetFromOptions) */
#define MGB_SMOOTH_JACOBI  0  /* -pc_type jacobi : x += scale * D^{-1} (b - A x)                            */
#define MGB_SMOOTH_RBSOR   1  /* -pc_type sor on the red-black numbering (-map 3): PETSc MatSOR semantics,
                                 reds ((i+j) even) before blacks                                             */
#define MGB_SMOOTH_LEXSOR  2  /* -pc_type sor on the natural numbering (-map 0,1,2): PETSc's lexicographic MatSOR, run as
                                 anti-diagonal wavefronts -- same operations, same order, same bits (csrc/mgb_wave.cuh)       */
#define MGB_SMOOTH_ILU0    3  /* PETSc's default PC (no -pc_type): ILU(0) + triangular solves, also as wavefronts             */
#define MGB_SOR_SYMMETRIC  0  /* SOR_LOCAL_SYMMETRIC_SWEEP (PCSOR default): forward then backward           */
#define MGB_SOR_FORWARD    1
#define MGB_SOR_BACKWARD   2

typedef struct {
	int    type;        /* MGB_SMOOTH_*                                                     */
	double scale;       /* -ksp_richardson_scale (damping of Richardson; 1.0 default)       */
	double omega;       /* -pc_sor_omega (1.0 default)                                      */
	int    sor_sweep;   /* MGB_SOR_*                                                        */
	int    sor_its;     /* -pc_sor_its * -pc_sor_lits (1 default)                           */
} mgb_smoother;

typedef struct {
	int levels;         /* number of levels, one grid per level (ref: -levels == -grids)    */
	int ni, nj;         /* interior rows / columns of the finest level                      */
	int device;         /* CUDA device ordinal, -1 = current device                         */
	int red_black_numbering; /* 1: -map 3 extension (CSR assembly is then not offered)      */
	/* strip decomposition (one engine per rank; rank 0 of 1 on a single GPU) */
	int rank, nranks;
	int agglomerate_below; /* levels with ni <= this are kept whole on rank 0 (multi-GPU only; 0 = default 511) */
	int emulate;           /* 1: hold ALL nranks strips in this process on one GPU, run them in lock step (tests) */
} mgb_config;

/* ---- lifetime ------------------------------------------------------------------------------ */
int  mgb_create(const mgb_config *cfg, mgb_engine **out);
int  mgb_destroy(mgb_engine *e);
const char *mgb_last_error(void);
int  mgb_version(void);
/* ---- row strips over several GPUs (replaces the reference's per-rank row ranges, ref: src/matbuild.c:120-144,
 * and the MPI plumbing PETSc does for it) ------------------------------------------------------------------
 * One process per GPU creates an engine with (rank, nranks); mgb_ipc_export gives a 64-byte CUDA IPC handle of
 * the engine's HBM arena; the handles of all ranks, concatenated in rank order (exchanged by the caller, e.g.
 * with torch.distributed.all_gather), go to mgb_ipc_connect.  After that every call below is COLLECTIVE: all
 * ranks must make the same calls in the same order.  Host vectors stay whole-grid arrays; each rank reads /
 * writes only its own rows of them (mgb_local_rows). */
#define MGB_IPC_HANDLE_BYTES 64
int  mgb_ipc_export(mgb_engine *e, void *handle);
int  mgb_ipc_connect(mgb_engine *e, const void *handles);
int  mgb_local_rows(const mgb_engine *e, int level, int *row0, int *row1);
/* the partition itself (host arithmetic, no GPU needed): rows [row0,row1) of `level` that `rank` handles */
int  mgb_strip_rows(const mgb_config *cfg, int level, int rank, int *row0, int *row1, int *distributed);

/* dimensions of level l as the engine derived them (ref: src/matbuild.c:64-66: n_l = (N-1)/2^l - 1) */
int  mgb_level_dims(const mgb_engine *e, int level, int *ni, int *nj);

/* ---- operator definition ------------------------------------------------------------------- */
/* row_coeff: ni rows x 5 doubles (S, W, C, E, N) -- the output of prob->OpA for grid row i of this level.
 * The reference's metrics depend on y only (src/mesh.c:29-107), so one 5-tuple per grid row defines A[l]. */
int  mgb_set_level_operator(mgb_engine *e, int level, const double *row_coeff);
/* 3x3 full-weighting restriction and bilinear prolongation weights, row-major (a = row offset, b = col offset) */
int  mgb_set_transfer(mgb_engine *e, const double res3[9], const double pro3[9]);

/* ---- assembled (CSR / AIJ) operators, generated on device ----------------------------------- */
int  mgb_assemble_csr(mgb_engine *e);                     /* A[l] all levels, res[l], pro[l]        */
int  mgb_csr_dims(const mgb_engine *e, int which, int level, int *m, int *n, long long *nnz);
int  mgb_csr_get(const mgb_engine *e, int which, int level, int *rowptr, int *col, double *val);
/* With row strips each rank assembles the rows it owns (local row pointers, GLOBAL column indices -- the local rows of
 * the reference's MPIAIJ matrices, ref: src/solver.c:218,502).  rank = -1: the first strip held by this process;
 * row0 = global number of the first local row. */
int  mgb_csr_dims_rank(const mgb_engine *e, int rank, int which, int level, int *m, int *n, long long *nnz, int *row0);
int  mgb_csr_get_rank(const mgb_engine *e, int rank, int which, int level, int *rowptr, int *col, double *val);
/* y = M x with the assembled matrix (MatMult_SeqAIJ order: ascending columns from 0.0); host vectors */
int  mgb_csr_spmv(mgb_engine *e, int which, int level, const double *x, double *y);
/* same, on the engine's own level vectors (no host traffic): y_vec[lout] = M x_vec[lin] */
int  mgb_csr_spmv_vec(mgb_engine *e, int which, int level, int x_vec, int y_vec);

/* ---- vectors ------------------------------------------------------------------------------- */
int  mgb_vec_set(mgb_engine *e, int which, int level, const double *host);      /* ni*nj, natural order */
int  mgb_vec_get(mgb_engine *e, int which, int level, double *host);
int  mgb_vec_zero(mgb_engine *e, int which, int level);
int  mgb_set_rhs(mgb_engine *e, const double *b0);                               /* = mgb_vec_set(B, 0)  */
/* b0[i][j] = gx[j] * gy[i], generated on device from two host tables (the reference RHS is separable:
 * ((-2*PI*PI)*sin(PI*x_j)) * sin(PI*y_i), src/problem.c:27) -- one multiply, same rounding as the host */
int  mgb_set_rhs_separable(mgb_engine *e, const double *gx, const double *gy);
int  mgb_get_solution(mgb_engine *e, double *u0);                                /* = mgb_vec_get(U, 0)  */
/* error = { max|u-s|, sum|u-s|, sqrt(sum (u-s)^2) } against s[i][j] = sx[j]*sy[i] (ref: src/solver.c:1211-1237) */
int  mgb_error_norms_separable(mgb_engine *e, const double *sx, const double *sy, double error[3]);

/* ---- single operations on level vectors (kernel-level parity tests and bandwidth sweeps) ---- */
int  mgb_op_apply(mgb_engine *e, int level, int x_vec, int y_vec);               /* y = A x               */
int  mgb_op_residual(mgb_engine *e, int level);                                  /* R = B - A U           */
int  mgb_op_residual_norm(mgb_engine *e, int level, double *norm);               /* ||B - A U||_2, no store */
int  mgb_op_smooth(mgb_engine *e, int level, const mgb_smoother *s, int its, int guess_zero);
int  mgb_op_restrict(mgb_engine *e, int level, int fused);  /* B[l+1] = res[l] * (fused ? B[l]-A U[l] : R[l]) */
int  mgb_op_prolong(mgb_engine *e, int level, int multadd); /* U[l] += pro[l] U[l+1]; multadd: MatMultAdd order */
int  mgb_op_norm2(mgb_engine *e, int which, int level, double *out);
int  mgb_op_dot(mgb_engine *e, int xw, int yw, int level, double *out);
int  mgb_op_axpy(mgb_engine *e, int yw, double alpha, int xw, int level);        /* y = y + alpha x       */
int  mgb_op_aypx(mgb_engine *e, int yw, double beta, int xw, int level);         /* y = x + beta y        */

/* ---- solvers ------------------------------------------------------------------------------- */
typedef struct {
	mgb_smoother smoother;  /* every level (the reference's KSPs share the un-prefixed options)     */
	int    v0, v1;          /* -v a,b : sweeps on levels 0..L-2 / on the coarsest (src/solver.c:1463-1510) */
	int    max_iter;        /* -iter                                                             */
	double rtol;            /* 1e-7 in the reference (src/solver.c:1530); extension knob -rtol   */
	int    use_graph;       /* 1: replay the cycle as a CUDA graph (same kernels, same order)    */
	int    no_fuse;         /* 1: one kernel per sweep / transfer (default 0: each leg of a level in one pass,
	                           csrc/mgb_fused.cuh -- same arithmetic per value, bit-identical results)          */
	int    no_bottom;       /* 1: do not run the levels with <= 63 rows as one persistent cluster launch
	                           (csrc/mgb_coarse_cycle.cuh); only meaningful with the fused legs                 */
} mgb_vcycle_params;

/* rnorm: max_iter+1 doubles; on return rnorm[0..num_iter] are the RELATIVE residual norms
 * (src/solver.c:1554-1557).  seconds = host wall time of the call; mgb_last_solve_ms() = device time of the cycle
 * loop only, the region the reference brackets with MPI_Wtime (src/solver.c:1526-1553). */
int  mgb_solve_vcycle(mgb_engine *e, const mgb_vcycle_params *p, double *rnorm, int *num_iter, double *seconds);

/* The same solve for a stream of independent right-hand sides, pipelined: the upload of right-hand side k+1 and the
 * download of solution k-1 overlap with solve k (pinned host memory needed for real overlap).  b_hosts / u_hosts:
 * nrhs whole-grid host arrays each; num_iter / final_rnorm: nrhs entries (may be NULL); seconds: wall time of the
 * whole batch including every host<->device copy. */
int  mgb_solve_vcycle_many(mgb_engine *e, const mgb_vcycle_params *p, int nrhs, const double *const *b_hosts,
                           double *const *u_hosts, int *num_iter, double *final_rnorm, double *seconds);

#define MGB_KSP_RICHARDSON 0
#define MGB_KSP_CG         1
#define MGB_COARSE_LU          0   /* -mg_coarse_ksp_type preonly -mg_coarse_pc_type lu (PCMG default)   */
#define MGB_COARSE_RICHARDSON  1   /* -mg_coarse_ksp_type richardson + coarse smoother, coarse_its sweeps */
typedef struct {
	int    outer;            /* MGB_KSP_* : -ksp_type                                              */
	double rtol, abstol, dtol; /* -ksp_rtol (1e-7 at src/solver.c:1924) -ksp_atol (1e-50) -ksp_divtol (1e4) */
	int    max_iter;         /* -iter / -ksp_max_it                                               */
	mgb_smoother level_smoother; int level_its;     /* -mg_levels_*                               */
	int    coarse;           /* MGB_COARSE_*                                                      */
	mgb_smoother coarse_smoother; int coarse_its;   /* -mg_coarse_*                               */
	int    no_fuse, no_bottom; /* as in mgb_vcycle_params                                          */
	int    no_graph;         /* 1: launch the kernels of a CG iteration one by one (default 0: one CUDA graph per iteration) */
} mgb_pcmg_params;
/* reason: >0 converged (2 = rtol, 3 = atol, 4 = its), <0 diverged (PETSc KSPConvergedReason values) */
int  mgb_solve_pcmg(mgb_engine *e, const mgb_pcmg_params *p, double *rnorm, int *num_iter, int *reason, double *seconds);

/* ---- measurement --------------------------------------------------------------------------- */
/* number of kernels this engine has launched since creation (bench.py's gpu_launches) */
long long mgb_launch_count(const mgb_engine *e);
/* device time (CUDA events on the engine's stream) of the cycle loop of the last mgb_solve_* call, in ms:
 * the same region as the reference's MPI_Wtime bracket (ref: src/solver.c:1526-1553) */
double mgb_last_solve_ms(const mgb_engine *e);
/* time `reps` back-to-back launches of one operation with CUDA events on the engine's stream;
 * op: 0 apply, 1 residual, 2 jacobi sweep, 3 red-black full sweep (2 half sweeps), 4 fused residual+restrict,
 * 5 prolong+correct, 6 residual norm, 7 csr spmv (A), 8 nrm2, 9 dot, 10 axpy; fused legs (mgb_fused.cuh):
 * 11 down leg (3 sweeps + residual + restriction), 12 up leg (prolongation + 3 sweeps + residual norm),
 * 13 three sweeps, 14 one sweep, 15 down leg from a zero guess; 16 the persistent bottom kernel from `level`
 * down to the coarsest and back (mgb_coarse_cycle.cuh); 17 ghost-row exchange of U alone, 18 nrm2 with the all-reduce
 * over the ranks (row strips).  ms_per_launch is the average. */
int  mgb_time_op(mgb_engine *e, int op, int level, int reps, double *ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif
