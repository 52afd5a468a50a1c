/*
 * mgb200_sparse.h -- C-ABI of the general sparse / dense-vector objects of the B200 engine (lib/libmgb200.so).
 *
 * the reference's UNMODIFIED src/solver.c (all cycles, several grids per level) runs on the
 * GPU through a PETSc-subset layer (multigrid-petsc_b200/host/petsc_b200/) that owns no arithmetic of its own: every
 * Mat / Vec operation PETSc performs for the reference is one of the calls below.  Plain C, plain pointers and sizes,
 * int status (0 = ok, negative MGB_E* as in mgb200.h, text via mgb_last_error()).  No CPU fallback: without a CUDA device
 * every constructor fails with MGB_ECUDA.
 *
 * Arithmetic contract (what makes the results comparable with the reference's CPU run bit for bit): IEEE binary64,
 * round to nearest, NO fused multiply-add; every routine performs the operations of the PETSc routine it replaces in
 * PETSc's order (citations below are to the reference's call sites, "ref:" = /root/reference, and to the PETSc routine
 * names as restated in oracle/minipetsc/minipetsc.c [PETSc-upstream]).  Only the SCHEDULE is parallel.
 */
#ifndef MGB200_SPARSE_H
#define MGB200_SPARSE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgb_dvec mgb_dvec;      /* dense fp64 vector resident in HBM                       (PETSc Vec) */
typedef struct mgb_dcsr mgb_dcsr;      /* CSR matrix resident in HBM, int32 indices, ascending columns per row (PETSc SeqAIJ) */
typedef struct mgb_dindex mgb_dindex;  /* index list resident in HBM                              (PETSc IS)  */

/* ---- vectors ------------------------------------------------------------------------------------------------------- */
int mgb_dvec_create(int n, mgb_dvec **out);                               /* zero-filled; ref: MatCreateVecs src/solver.c:1172, VecDuplicate */
int mgb_dvec_destroy(mgb_dvec *v);
int mgb_dvec_size(const mgb_dvec *v);
int mgb_dvec_upload(mgb_dvec *v, const double *host);                     /* after VecSetValue / VecRestoreArray, ref: src/solver.c:597,616-617 */
int mgb_dvec_download(const mgb_dvec *v, double *host);                   /* VecGetArray, ref: src/solver.c:1255 */
int mgb_dvec_set(mgb_dvec *v, double alpha);                              /* VecSet */
int mgb_dvec_copy(mgb_dvec *dst, const mgb_dvec *src);                    /* VecCopy */
int mgb_dvec_axpy(mgb_dvec *y, double alpha, const mgb_dvec *x);          /* y = y + alpha x      ref: src/solver.c:1541 */
int mgb_dvec_aypx(mgb_dvec *y, double beta, const mgb_dvec *x);           /* y = x + beta y */
int mgb_dvec_waxpy(mgb_dvec *w, double alpha, const mgb_dvec *x, const mgb_dvec *y);   /* w = alpha x + y */
int mgb_dvec_axpbypcz(mgb_dvec *z, double alpha, double beta, double gamma, const mgb_dvec *x, const mgb_dvec *y);
                                                                          /* z = gamma z + alpha x + beta y   ref: src/solver.c:1839,1845 */
int mgb_dvec_pointwise_mult(mgb_dvec *w, const mgb_dvec *x, const mgb_dvec *y);        /* PCApply_Jacobi */
int mgb_dvec_scale(mgb_dvec *x, double alpha);
/* Reductions are deterministic: fixed 4096-element blocks, each summed left to right, then the block sums left to right. */
int mgb_dvec_dot(const mgb_dvec *x, const mgb_dvec *y, double *out);      /* VecDot / VecTDot     ref: src/solver.c:1674-1681 */
#define MGB_NORM_1 0
#define MGB_NORM_2 1
#define MGB_NORM_INF 3
int mgb_dvec_norm(const mgb_dvec *x, int type, double *out);              /* VecNorm              ref: src/solver.c:1512,1518,1546 */

/* ---- index lists and sub-vectors (research cycles: ref: src/solver.c:2210,2233-2235) -------------------------------- */
int mgb_dindex_create(int n, const int *idx, mgb_dindex **out);
int mgb_dindex_destroy(mgb_dindex *is);
int mgb_dvec_gather(mgb_dvec *sub, const mgb_dvec *x, const mgb_dindex *is);     /* sub[k] = x[idx[k]]   VecGetSubVector */
int mgb_dvec_scatter(mgb_dvec *x, const mgb_dvec *sub, const mgb_dindex *is);    /* x[idx[k]] = sub[k]   VecRestoreSubVector */

/* ---- CSR matrices -------------------------------------------------------------------------------------------------- */
/* rowptr has m+1 entries; columns ascending within a row (what MatAssemblyEnd leaves; ref: src/solver.c:508-509) */
int mgb_dcsr_create(int m, int n, const int *rowptr, const int *col, const double *val, mgb_dcsr **out);
int mgb_dcsr_destroy(mgb_dcsr *A);
int mgb_dcsr_mult(const mgb_dcsr *A, const mgb_dvec *x, mgb_dvec *y);             /* MatMult_SeqAIJ: sum from 0.0 in column order   ref: src/solver.c:1535,1540 */
int mgb_dcsr_mult_add(const mgb_dcsr *A, const mgb_dvec *x, const mgb_dvec *y, mgb_dvec *z);   /* z_i = y_i + sum (MatMultAdd; sum starts from y_i) */
int mgb_dcsr_scale(mgb_dcsr *A, double s);                                        /* MatScale             ref: src/solver.c:2104 */
int mgb_dcsr_inverse_diagonal(const mgb_dcsr *A, mgb_dvec *dinv);                 /* PCSetUp_Jacobi: 1/d, 1 where d == 0 */
/* MatSOR_SeqAIJ (flag bits as PETSc's MatSORType: 1 forward, 2 backward, 4 / 8 local forward / backward, 16 zero initial
 * guess).  The sweeps are sequential in the row number; rows whose dependencies are complete run together (level sets of
 * the triangular parts, computed once per matrix), so every row sees exactly the operands of the sequential loop.
 * Needs a structurally symmetric pattern with a full diagonal (checked). */
int mgb_dcsr_sor(mgb_dcsr *A, const mgb_dvec *b, double omega, int flag, double fshift, int its, int lits, mgb_dvec *x);
/* ILU(0) in the natural ordering (PETSc's default PC on one rank), inverted pivots, and its two triangular solves */
int mgb_dcsr_ilu0_factor(mgb_dcsr *A);
int mgb_dcsr_ilu0_solve(mgb_dcsr *A, const mgb_dvec *b, mgb_dvec *x);
/* dense LU without pivoting in the natural ordering (PCMG's coarse solve on a small grid; at most 4096 unknowns) */
int mgb_dcsr_lu_factor(mgb_dcsr *A);
int mgb_dcsr_lu_solve(mgb_dcsr *A, const mgb_dvec *b, mgb_dvec *x);
/* kernels launched by the calls above since the library was loaded (the bench / tests prove the GPU did the work) */
long long mgb_sparse_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
