/*
 * petsc_b200.h (reached as <petscksp.h>) -- the PETSc surface of the reference (SyamVangara/multigrid-petsc), served by the B200 engine.
 *
 * the reference's src/solver.c (all cycles) and src/poisson.c include <petscksp.h> and
 * call ~75 PETSc / MPI symbols (list: SURVEY.md section 8b).  This header declares exactly those, with PETSc's names
 * and signatures, and petsc_b200.c implements them on the general sparse objects of lib/libmgb200.so
 * (include/mgb200_sparse.h): a Mat is a CSR matrix in HBM, a Vec a dense vector in HBM, every MatMult / VecAXPY /
 * MatSOR / ILU(0) solve / dot product is a CUDA kernel.  Host code keeps only what PETSc keeps on the host: the options
 * database, the assembly slab behind MatSetValue, the Krylov recurrences' scalars and PCMG's recursion.  One rank
 * (MPI_Comm_size == 1): several GPUs are the strip engine's business (include/mgb200.h), not this layer's.
 *
 * Build: make -C multigrid-petsc_b200/host refsolver  ->  lib/poisson_petsc_b200 = the reference's six UNMODIFIED
 * source files compiled against this header + petsc_b200.c + lib/libmgb200.so.
 * There is no CPU path: the first object created without a CUDA device stops the program with the engine's message.
 * Not the CPU oracle's mini-PETSc (oracle/minipetsc is test infrastructure and is not linked here).
 */
#ifndef PETSC_B200_H
#define PETSC_B200_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- basic types */
typedef int    PetscInt;
typedef int    PetscMPIInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef int    PetscErrorCode;
typedef int    PetscLogStage;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;
typedef enum { PETSC_COPY_VALUES, PETSC_OWN_POINTER, PETSC_USE_POINTER } PetscCopyMode;

#define PETSC_DEFAULT   (-2)
#define PETSC_DECIDE    (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_NULL      NULL
#define PETSC_STDOUT    stdout
#define PETSC_MAX_REAL  1.7976931348623157e308

/* ---------------------------------------------------------------- 1-rank MPI */
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int dummy; } MPI_Status;
#define MPI_COMM_WORLD    0
#define PETSC_COMM_WORLD  0
#define PETSC_COMM_SELF   0
#define MPI_DOUBLE        1
#define MPI_INT           2
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
int    MPI_Comm_size(MPI_Comm comm, int *size);
int    MPI_Comm_rank(MPI_Comm comm, int *rank);
double MPI_Wtime(void);
int    MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm);
int    MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Status *st);

/* ---------------------------------------------------------------- objects */
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_IS  *IS;
typedef struct _p_KSP *KSP;
typedef struct _p_PC  *PC;
typedef void          *PetscObject;
typedef int            PetscViewer;
#define PETSC_VIEWER_STDOUT_WORLD 1
#define PETSC_VIEWER_STDOUT_SELF  1
#define PETSC_VIEWER_DRAW_WORLD   2

typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { MAT_INITIAL_MATRIX, MAT_REUSE_MATRIX } MatReuse;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
typedef enum { KSP_NORM_DEFAULT = -1, KSP_NORM_NONE = 0, KSP_NORM_PRECONDITIONED = 1,
               KSP_NORM_UNPRECONDITIONED = 2, KSP_NORM_NATURAL = 3 } KSPNormType;
typedef enum { KSP_CONVERGED_ITERATING = 0, KSP_CONVERGED_RTOL = 2, KSP_CONVERGED_ATOL = 3,
               KSP_CONVERGED_ITS = 4, KSP_DIVERGED_ITS = -3, KSP_DIVERGED_DTOL = -4,
               KSP_DIVERGED_INDEFINITE_PC = -8, KSP_DIVERGED_INDEFINITE_MAT = -10 } KSPConvergedReason;
typedef enum { PC_ASM_BASIC = 3, PC_ASM_RESTRICT = 1, PC_ASM_INTERPOLATE = 2, PC_ASM_NONE = 0 } PCASMType;

/* MatSORType bit flags, PETSc's values */
typedef enum { SOR_FORWARD_SWEEP = 1, SOR_BACKWARD_SWEEP = 2, SOR_SYMMETRIC_SWEEP = 3,
               SOR_LOCAL_FORWARD_SWEEP = 4, SOR_LOCAL_BACKWARD_SWEEP = 8, SOR_LOCAL_SYMMETRIC_SWEEP = 12,
               SOR_ZERO_INITIAL_GUESS = 16, SOR_EISENSTAT = 32, SOR_APPLY_UPPER = 64, SOR_APPLY_LOWER = 128 } MatSORType;

typedef const char *KSPType;
typedef const char *PCType;
#define KSPRICHARDSON "richardson"
#define KSPCG         "cg"
#define KSPGMRES      "gmres"
#define KSPPREONLY    "preonly"
#define PCNONE        "none"
#define PCJACOBI      "jacobi"
#define PCSOR         "sor"
#define PCILU         "ilu"
#define PCLU          "lu"
#define PCMG          "mg"
#define PCASM         "asm"
#define PCBJACOBI     "bjacobi"

/* ---------------------------------------------------------------- sys / options */
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscOptionsInsertString(void *options, const char *str);
PetscErrorCode PetscOptionsClear(void *options);
PetscErrorCode PetscOptionsGetInt(void *options, const char *pre, const char *name, PetscInt *ivalue, PetscBool *set);
PetscErrorCode PetscOptionsGetIntArray(void *options, const char *pre, const char *name, PetscInt *ivalue, PetscInt *nmax, PetscBool *set);
PetscErrorCode PetscOptionsGetReal(void *options, const char *pre, const char *name, PetscReal *dvalue, PetscBool *set);
PetscErrorCode PetscOptionsGetString(void *options, const char *pre, const char *name, char *str, size_t len, PetscBool *set);
PetscErrorCode PetscOptionsHasName(void *options, const char *pre, const char *name, PetscBool *set);
PetscErrorCode PetscPrintf(MPI_Comm comm, const char *fmt, ...);
PetscErrorCode PetscSynchronizedPrintf(MPI_Comm comm, const char *fmt, ...);
PetscErrorCode PetscSynchronizedFlush(MPI_Comm comm, FILE *fd);
PetscErrorCode PetscLogStageRegister(const char *name, PetscLogStage *stage);
PetscErrorCode PetscLogStagePush(PetscLogStage stage);
PetscErrorCode PetscLogStagePop(void);
PetscErrorCode PetscObjectSetOptionsPrefix(PetscObject obj, const char *prefix);
/* extension: kernels launched on the GPU so far (lib/libmgb200.so: mgb_sparse_launch_count) */
long long PetscB200LaunchCount(void);

/* ---------------------------------------------------------------- Vec */
PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec *v);
PetscErrorCode VecDuplicate(Vec v, Vec *newv);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecSet(Vec x, PetscScalar alpha);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecSetValue(Vec x, PetscInt row, PetscScalar value, InsertMode mode);
PetscErrorCode VecAssemblyBegin(Vec x);
PetscErrorCode VecAssemblyEnd(Vec x);
PetscErrorCode VecGetSize(Vec x, PetscInt *n);
PetscErrorCode VecGetArray(Vec x, PetscScalar **a);
PetscErrorCode VecRestoreArray(Vec x, PetscScalar **a);
PetscErrorCode VecGetOwnershipRange(Vec x, PetscInt *low, PetscInt *high);
PetscErrorCode VecGetOwnershipRanges(Vec x, const PetscInt *ranges[]);
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val);
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val);
PetscErrorCode VecTDot(Vec x, Vec y, PetscScalar *val);
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x);          /* y = y + alpha x */
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x);           /* y = x + beta y  */
PetscErrorCode VecWAXPY(Vec w, PetscScalar alpha, Vec x, Vec y);  /* w = alpha x + y */
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar alpha, PetscScalar beta, PetscScalar gamma, Vec x, Vec y);
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);
PetscErrorCode VecScale(Vec x, PetscScalar alpha);
PetscErrorCode VecGetSubVector(Vec x, IS is, Vec *y);
PetscErrorCode VecRestoreSubVector(Vec x, IS is, Vec *y);
PetscErrorCode VecView(Vec x, PetscViewer viewer);

/* ---------------------------------------------------------------- IS */
PetscErrorCode ISCreateGeneral(MPI_Comm comm, PetscInt n, const PetscInt idx[], PetscCopyMode mode, IS *is);
PetscErrorCode ISDestroy(IS *is);
PetscErrorCode ISView(IS is, PetscViewer viewer);

/* ---------------------------------------------------------------- Mat (SeqAIJ) */
PetscErrorCode MatCreateAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N,
                            PetscInt d_nz, const PetscInt d_nnz[], PetscInt o_nz, const PetscInt o_nnz[], Mat *A);
PetscErrorCode MatCreateSeqAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt nz, const PetscInt nnz[], Mat *A);
PetscErrorCode MatSetValue(Mat A, PetscInt row, PetscInt col, PetscScalar v, InsertMode mode);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType type);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType type);
PetscErrorCode MatDestroy(Mat *A);
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n);
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatMultAdd(Mat A, Vec x, Vec y, Vec z);             /* z = y + A x */
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);
PetscErrorCode MatResidual(Mat A, Vec b, Vec x, Vec r);            /* r = b - A x */
PetscErrorCode MatRestrict(Mat A, Vec x, Vec y);
PetscErrorCode MatInterpolateAdd(Mat A, Vec x, Vec y, Vec w);      /* w = y + A x */
PetscErrorCode MatScale(Mat A, PetscScalar a);
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat *C);
PetscErrorCode MatGetDiagonal(Mat A, Vec d);
PetscErrorCode MatSOR(Mat A, Vec b, PetscReal omega, MatSORType flag, PetscReal shift,
                      PetscInt its, PetscInt lits, Vec x);
PetscErrorCode MatView(Mat A, PetscViewer viewer);

/* ---------------------------------------------------------------- PC */
PetscErrorCode PCSetType(PC pc, PCType type);
PetscErrorCode PCMGSetLevels(PC pc, PetscInt levels, MPI_Comm *comms);
PetscErrorCode PCMGGetCoarseSolve(PC pc, KSP *ksp);
PetscErrorCode PCMGGetSmoother(PC pc, PetscInt l, KSP *ksp);
PetscErrorCode PCMGSetInterpolation(PC pc, PetscInt l, Mat mat);
PetscErrorCode PCMGSetRestriction(PC pc, PetscInt l, Mat mat);
PetscErrorCode PCMGSetR(PC pc, PetscInt l, Vec c);
PetscErrorCode PCMGSetRhs(PC pc, PetscInt l, Vec c);
PetscErrorCode PCMGSetX(PC pc, PetscInt l, Vec c);
PetscErrorCode PCMGSetNumberSmoothUp(PC pc, PetscInt n);
PetscErrorCode PCMGSetNumberSmoothDown(PC pc, PetscInt n);
PetscErrorCode PCASMSetType(PC pc, PCASMType type);
PetscErrorCode PCASMSetOverlap(PC pc, PetscInt ovl);
PetscErrorCode PCASMSetTotalSubdomains(PC pc, PetscInt N, IS is[], IS is_local[]);
PetscErrorCode PCApply(PC pc, Vec x, Vec y);

/* ---------------------------------------------------------------- KSP */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp);
PetscErrorCode KSPDestroy(KSP *ksp);
PetscErrorCode KSPSetType(KSP ksp, KSPType type);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P);
PetscErrorCode KSPSetNormType(KSP ksp, KSPNormType t);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPSetFromOptions(KSP ksp);
PetscErrorCode KSPSetInitialGuessNonzero(KSP ksp, PetscBool flg);
PetscErrorCode KSPRichardsonSetScale(KSP ksp, PetscReal scale);
PetscErrorCode KSPGetPC(KSP ksp, PC *pc);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPBuildResidual(KSP ksp, Vec t, Vec v, Vec *V);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its);
PetscErrorCode KSPGetConvergedReason(KSP ksp, KSPConvergedReason *reason);
PetscErrorCode KSPSetResidualHistory(KSP ksp, PetscReal a[], PetscInt na, PetscBool reset);
PetscErrorCode KSPGetResidualHistory(KSP ksp, PetscReal *a[], PetscInt *na);
PetscErrorCode KSPMonitorSet(KSP ksp, PetscErrorCode (*monitor)(KSP, PetscInt, PetscReal, void *), void *mctx,
                             PetscErrorCode (*monitordestroy)(void **));
PetscErrorCode KSPView(KSP ksp, PetscViewer viewer);

#ifdef __cplusplus
}
#endif
#endif
