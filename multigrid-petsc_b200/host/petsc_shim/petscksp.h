/*
 * petscksp.h -- PRODUCT-SIDE stand-in for the PETSc header the reference's driver files include
 * (ref: include/header.h:12, include/mesh.h:11, include/solver.h:14).
 *
 * Purpose: the true drop-in build.  The reference's UNMODIFIED src/poisson.c, src/mesh.c, src/problem.c,
 * src/matbuild.c and src/array.c are compiled in place against the reference's own headers plus this file, and
 * linked with multigrid-petsc_b200/host/solver_b200.c (built with -DPB_USE_REFERENCE_HEADERS) instead of
 * src/solver.c: only solver.c is swapped (recipe: host/Makefile target `dropin`).
 *
 * Those five files use PETSc for exactly three things (full list: SURVEY.md 8b), all provided here without any
 * linear algebra:
 *   options      PetscInitialize(&argc, &argv, "poisson.in", 0), PetscOptionsGetInt / GetIntArray   (src/poisson.c:29, 51-59)
 *                -> the pbopt_* options database of this build (file first, argv wins, '#' comments)
 *   printing     PetscPrintf, PetscSynchronizedPrintf / Flush                                        (src/poisson.c:165-214, View*)
 *   rank / size  MPI_Comm_rank / MPI_Comm_size on PETSC_COMM_WORLD: one process, rank 0 of 1
 * plus the opaque handle types (Mat, Vec, IS, KSP, PC) that appear in struct Assembly and the viewers the
 * never-called View* helpers of src/poisson.c mention (no-ops).  This is NOT the CPU oracle's mini-PETSc
 * (oracle/minipetsc, test infrastructure): nothing here computes anything.
 */
#ifndef PB200_PETSC_SHIM_H
#define PB200_PETSC_SHIM_H

#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int PetscInt;
typedef int PetscErrorCode;
typedef double PetscScalar;
typedef double PetscReal;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
typedef struct _p_Mat *Mat;
typedef struct _p_Vec *Vec;
typedef struct _p_IS *IS;
typedef struct _p_KSP *KSP;
typedef struct _p_PC *PC;
typedef struct _p_PetscViewer *PetscViewer;

#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 0
#define PETSC_STDOUT stdout
#define PETSC_NULL NULL
#define PETSC_VIEWER_STDOUT_WORLD ((PetscViewer)0)
#define PETSC_VIEWER_DRAW_WORLD ((PetscViewer)0)

/* the options database of the B200 build (host/pb_options.c) */
void pbopt_clear(void);
void pbopt_insert_file(const char *path);
void pbopt_insert_args(int argc, char **argv);
int pbopt_get_int(const char *name, int *v);
int pbopt_get_int_array(const char *name, int *v, int *n);

static inline PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help)
{
	(void)help;
	pbopt_clear();
	if (file) pbopt_insert_file(file);
	if (argc && argv) pbopt_insert_args(*argc, *argv);
	return 0;
}
static inline PetscErrorCode PetscFinalize(void) { return 0; }
static inline PetscErrorCode PetscOptionsGetInt(void *opts, const char *pre, const char *name, PetscInt *v, PetscBool *set)
{
	(void)opts; (void)pre;
	const int found = pbopt_get_int(name, v);
	if (set) *set = found ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
static inline PetscErrorCode PetscOptionsGetIntArray(void *opts, const char *pre, const char *name, PetscInt *v, PetscInt *n, PetscBool *set)
{
	(void)opts; (void)pre;
	const int found = pbopt_get_int_array(name, v, n);
	if (set) *set = found ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
static inline PetscErrorCode PetscPrintf(MPI_Comm comm, const char *fmt, ...)
{
	(void)comm;
	va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
	return 0;
}
static inline PetscErrorCode PetscSynchronizedPrintf(MPI_Comm comm, const char *fmt, ...)
{
	(void)comm;
	va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
	return 0;
}
static inline PetscErrorCode PetscSynchronizedFlush(MPI_Comm comm, FILE *f) { (void)comm; fflush(f); return 0; }

static inline int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = 0; return 0; }

/* viewers of assembled PETSc objects: the engine keeps its operators in HBM (mgb_csr_get downloads them) */
static inline PetscErrorCode MatView(Mat m, PetscViewer v) { (void)m; (void)v; return 0; }
static inline PetscErrorCode VecView(Vec x, PetscViewer v) { (void)x; (void)v; return 0; }
static inline PetscErrorCode ISView(IS s, PetscViewer v) { (void)s; (void)v; return 0; }

#endif
