/*
 * pb_api.h -- the reference's driver-facing C interface, restated for the B200 build.
 *
 * The reference's main() (ref: src/poisson.c:27-138) owns six stack structs and calls a fixed sequence of
 * functions on them; that sequence and those structs ARE the drop-in boundary.  This header
 * declares the same types with the same member names, order and meaning, and the same function names and
 * signatures, so that (a) host code written against the reference's headers compiles against this one and
 * (b) multigrid-petsc_b200/host/solver_b200.c compiles unchanged against either (define
 * PB_USE_REFERENCE_HEADERS and put the reference's include/ on the include path to build the true drop-in;
 * see INTEGRATION.md).  Where a member only made sense with PETSc (Mat/Vec/IS arrays) it is kept as an
 * opaque pointer so the struct layout is unchanged.
 *
 *   type            reference declaration
 *   Array2d/Int2d   include/array.h:31-40
 *   Problem         include/problem.h:19-24
 *   Mesh, MeshType  include/mesh.h:19-27
 *   Level, Indices  include/solver.h:17-31
 *   Operator        include/solver.h:33-37
 *   Assembly        include/solver.h:39-52
 *   Cycle, Solver   include/solver.h:54-70
 *   PostProcess     include/solver.h:72-79
 *   functions       include/solver.h:81-98, include/mesh.h:29-30, include/problem.h:26, include/array.h:62-68
 */
#ifndef PB_API_H
#define PB_API_H

#ifdef PB_USE_REFERENCE_HEADERS
#include "header.h"          /* the reference's own include/header.h (needs a petscksp.h on the path) */
#else

#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#define PI 3.14159265358979323846
#define DIMENSION 2

/* opaque stand-ins for the PETSc handles that appear in Assembly */
typedef struct _p_Mat *Mat;
typedef struct _p_Vec *Vec;
typedef struct _p_IS  *IS;

typedef struct { int ni; int nj; double *data; } Array2d;
typedef struct { int ni; int nj; int *data; } ArrayInt2d;

typedef struct {
	double (*Ffunc)(double x, double y);                       /* right-hand side f(x,y)            */
	double (*SOLfunc)(double x, double y);                     /* analytic solution                 */
	void   (*OpA)(double *A, double *metrics, double *h);      /* 5 stencil coefficients S,W,C,E,N  */
} Problem;

typedef enum { UNIFORM, NONUNIFORM1, NONUNIFORM2 } MeshType;

typedef struct {
	int    n[DIMENSION];                /* points per direction, boundary included */
	double bounds[DIMENSION * 2];
	double **coord;                     /* coord[0] = x (by column j), coord[1] = y (by row i) */
	double h;
	void   (*MetricCoefficients)(void *mesh, double x, double y, double *metrics);
} Mesh;

typedef struct {
	int        grids;                   /* grids in this level (always 1 on the accelerated path) */
	int        *gridId;
	double     (*h)[2];
	int        *ranges;
	ArrayInt2d global;                  /* global -> (i, j, gridId)  */
	ArrayInt2d *grid;                   /* grid (i, j) -> global     */
} Level;

typedef struct {
	int   levels;
	int   totalGrids;
	int   coarseningFactor;
	Level *level;
} Indices;

typedef struct {
	int     totalGrids;
	Array2d *res;
	Array2d *pro;
} Operator;

typedef struct {
	int levels;
	Mat *res;
	Mat *pro;
	Mat *A;
	Mat *A2;                            /* the B200 build parks its engine handle here (see solver_b200.c) */
	Vec *b;
	Vec *u;
	IS  *bottomIS;
	IS  *topIS;
	int moreInfo;
	IS  **gridIS;
} Assembly;

typedef enum { VCYCLE, ICYCLE, ECYCLE, D1CYCLE, D2CYCLE, D3CYCLE, D4CYCLE, D1PSCYCLE, PetscPCMG, ADDITIVE, ADDITIVE2 } Cycle;

typedef struct {
	Cycle    cycle;
	int      moreInfo;
	int      numIter;                   /* in: iteration limit (-iter) ; out: iterations done */
	int      v[2];
	int      grids;
	double   **rNormGrid;
	double   *rNormGlobal;
	double   *rnorm;                    /* numIter+1 relative residual norms */
	Assembly *assem;
} Solver;

typedef struct {
	double error[3];
	FILE   *solData;
	FILE   *errData;
	FILE   *resData;
	FILE   *XgridData;
	FILE   *YgridData;
} PostProcess;

void CreateArrayInt2d(int ni, int nj, ArrayInt2d *a);
void DeleteArrayInt2d(ArrayInt2d *a);
void CreateArray2d(int ni, int nj, Array2d *a);
void DeleteArray2d(Array2d *a);

void SetUpProblem(Problem *prob);
void SetUpMesh(Mesh *mesh, MeshType type);
void DestroyMesh(Mesh *mesh);

void SetUpIndices(Mesh *mesh, Indices *indices);
void DestroyIndices(Indices *indices);
void mapping(Indices *indices, int mappingStyleflag);
void SetUpOperator(Indices *indices, Operator *op);
void DestroyOperator(Operator *op);
void GridTransferOperators(Operator op, Indices indices);

void SetUpSolver(Indices *indices, Solver *solver, Cycle c);
void DestroySolver(Solver *solver);
void Solve(Solver *solver);
void SetUpPostProcess(PostProcess *pp);
void DestroyPostProcess(PostProcess *pp);
void Postprocessing(Problem *prob, Mesh *mesh, Indices *indices, Solver *solver, PostProcess *pp);
void Assemble(Problem *prob, Mesh *mesh, Indices *indices, Operator *op, Solver *solver);

#endif /* PB_USE_REFERENCE_HEADERS */

/* ---- B200-build additions (not in the reference) -------------------------------------------- */
/* options database: "poisson.in" first, then argv (argv wins), '#' comments -- the PetscInitialize /
 * PetscOptionsGet* behaviour the reference relies on (ref: src/poisson.c:29, 51-59) */
void pbopt_clear(void);
void pbopt_insert_file(const char *path);
void pbopt_insert_args(int argc, char **argv);
void pbopt_insert_string(const char *str);
int  pbopt_has(const char *name);                                  /* 1 if the key is present                */
int  pbopt_get_int(const char *name, int *v);                      /* 1 if present with a value (v updated)  */
int  pbopt_get_real(const char *name, double *v);
int  pbopt_get_string(const char *name, char *buf, size_t len);
int  pbopt_get_int_array(const char *name, int *v, int *n);        /* in: capacity, out: count               */
int  pbopt_get_bool(const char *name);                             /* present and not 0/false/no             */

/* the whole reference main() as a callable: options string in the poisson.in vocabulary, results out.
 * dir: where uData.dat/rData.dat/eData.dat/X/YgridData.dat are written (NULL: no files).
 * u (ni*nj doubles, may be NULL), rnorm (cap entries, may be NULL).  Returns 0 on success. */
typedef struct {
	int    num_iter;
	int    ni, nj;
	double error[3];
	double solve_seconds;
	int    levels;
	long long gpu_launches;
} pb200_result;
int pb200_run(const char *options, const char *dir, pb200_result *res, double *u, double *rnorm, int rnorm_cap);
/* the same in two steps: pb200_open = SetUpProblem .. Assemble ; pb200_solve = Solve .. PrintInfo */
typedef struct pb200_session pb200_session;
int  pb200_open(const char *options, pb200_session **out);
int  pb200_solve(pb200_session *s, const char *dir, pb200_result *res, double *u, double *rnorm, int rnorm_cap);
/* another solve on the assembled session: b (host, ni*nj) replaces the right-hand side, u (host) receives the solution */
int  pb200_solve_rhs(pb200_session *s, const double *b, double *u, pb200_result *res, double *rnorm, int rnorm_cap);
/* a stream of right-hand sides through the same Solve(): copies and solves pipelined; iters/finals: nrhs entries */
int  pb200_solve_rhs_many(pb200_session *s, int nrhs, const double *const *b, double *const *u, int *iters, double *finals, double *seconds);
void pb200_close(pb200_session *s);
struct mgb_engine *pb200_session_engine(pb200_session *s);
const char *pb200_last_error(void);
/* access to the engine behind an assembled Solver (tests: CSR download, single kernels) */
struct mgb_engine;
struct mgb_engine *pb200_engine(Solver *solver);

#endif
