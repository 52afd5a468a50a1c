/*
 * mg_oracle.c -- TEST INFRASTRUCTURE ONLY.  See mg_oracle.h.
 *
 * Restatement of the reference's hot path (one grid per level) on top of
 * oracle/minipetsc.  Every function cites the reference file:line it follows;
 * "ref:" paths are relative to /root/reference.  PARITY UNPINNED against real
 * PETSc; pinned bit-for-bit against the reference's own driver compiled over
 * the same minipetsc (tests/test_oracle_vs_ref.py).
 */
#include "mg_oracle.h"
#include "minipetsc/petscksp.h"

#define MGO_PI 3.14159265358979323846   /* ref: include/problem.h:13 */
#define MGO_ERR(...) do { fprintf(stderr, "mg_oracle: " __VA_ARGS__); fprintf(stderr, "\n"); } while (0)

typedef struct {
	int  ni, nj;              /* interior rows (y), columns (x)            ref: src/matbuild.c:64-66 */
	int  gridId;              /* = level index with one grid per level     ref: src/matbuild.c:27-47 */
	double h[2];              /* {1/(ni+1), 1/(nj+1)}                      ref: src/matbuild.c:99-104 */
	int *g2G;                 /* grid (i*nj+j) -> global                   ref: Level.grid[0]   */
	int *G2g;                 /* global -> (i,j,gridId), 3 ints per row    ref: Level.global    */
} OLevel;

struct MgoCtx {
	/* poisson.in keys, ref: src/poisson.c:51-59 */
	int npts, meshflag, maxIter, grids, levels, cycle, map, v[2], moreNorm;
	double rtol;              /* extension; 1e-7 in the reference (src/solver.c:1530) */
	/* mesh, ref: include/mesh.h:21-27 */
	int n[2]; double bounds[4]; double *coord[2]; double meshh;
	OLevel *lev;
	double res3[9], pro3[9];  /* ref: op.res[0], op.pro[0] */
	Mat *A, *res, *pro; Vec *b, *u;
	KSP *smoother;            /* lazily created for mgo_smooth */
	/* results */
	double *rnorm; int numIter; double solve_seconds; int solved;
};

/* ------------------------------------------------------------------ problem (ref: src/problem.c) */
/* ref: src/problem.c:3-22 -- (S, W, C, E, N) coefficients from metrics[5] and h[2] */
static void oracle_OpA(double *A, const double *met, const double *h)
{
	const double hx2 = h[0] * h[0];
	const double hy2 = h[1] * h[1];
	A[0] = (met[1] / hy2) - (met[3] / (2 * h[1]));
	A[1] = (met[0] / hx2) - (met[2] / (2 * h[0]));
	A[2] = -2.0 * ((met[0] / hx2) + (met[1] / hy2));
	A[3] = (met[0] / hx2) + (met[2] / (2 * h[0]));
	A[4] = (met[1] / hy2) + (met[3] / (2 * h[1]));
}
/* ref: src/problem.c:24-28 */
static double oracle_F(double x, double y) { return -2 * MGO_PI * MGO_PI * sin(MGO_PI * x) * sin(MGO_PI * y); }
/* ref: src/problem.c:30-34 */
static double oracle_SOL(double x, double y) { return sin(MGO_PI * x) * sin(MGO_PI * y); }

/* ------------------------------------------------------------------ mesh (ref: src/mesh.c) */
/* ref: src/mesh.c:29-107 -- metrics[0..4] at (x,y) for the three mesh types */
static void oracle_metrics(const MgoCtx *c, double x, double y, double *met)
{
	(void)x;
	const double *bd = c->bounds;
	if (c->meshflag == 0) {
		met[0] = 1.0; met[1] = 1.0; met[2] = 0.0; met[3] = 0.0; met[4] = 0.0;
	} else if (c->meshflag == 1) {
		const double t = ((bd[3] - bd[2]) * (bd[3] - bd[2]) - (bd[3] - y) * (bd[3] - y));
		met[0] = 1.0;
		met[1] = 4.0 / (MGO_PI * MGO_PI * t);
		met[2] = 0.0;
		met[3] = (-2.0 * (bd[3] - y)) / (MGO_PI * sqrt(t * t * t));
		met[4] = 0.0;
	} else {
		const double e = exp(2) - 1;
		const double q = (y - bd[2]) * e + (bd[3] - bd[2]);
		const double t = (e * e) / (q * q);
		met[0] = 1.0 / ((bd[1] - bd[0]) * (bd[1] - bd[0]));
		met[1] = 0.25 * t;
		met[2] = 0.0;
		met[3] = (-0.5) * t;
		met[4] = 0.0;
	}
}

/* ref: src/mesh.c:130-195 -- x uniform (accumulated), y by mesh type; mesh->h = sqrt(dx^2 + max dy^2) */
static int oracle_coords(MgoCtx *c)
{
	double d[2];
	for (int k = 0; k < 2; k++) if (c->n[k] < 2) { MGO_ERR("Need at least 2 points in each direction"); return 1; }
	double *x = c->coord[0], *y = c->coord[1];
	const int nx = c->n[0], ny = c->n[1];
	x[0] = c->bounds[0]; x[nx - 1] = c->bounds[1];
	d[0] = (x[nx - 1] - x[0]) / (nx - 1);
	for (int j = 1; j < nx - 1; j++) x[j] = x[j - 1] + d[0];

	y[0] = c->bounds[2]; y[ny - 1] = c->bounds[3];
	const double length = y[ny - 1] - y[0];
	const double step = length / (double)(ny - 1);
	d[1] = 0.0;
	for (int j = 1; j < ny - 1; j++) {
		if (c->meshflag == 1) y[j] = c->bounds[3] - length * (cos(MGO_PI * 0.5 * (j / (double)(ny - 1))));
		if (c->meshflag == 2) {
			const double eta = (j / (double)(ny - 1));
			y[j] = c->bounds[2] + length * ((exp(2 * eta) - 1) / (exp(2) - 1));
		}
		if (c->meshflag == 0) y[j] = y[j - 1] + step;
		d[1] = fmax(d[1], fabs(y[j] - y[j - 1]));
	}
	d[1] = fmax(d[1], fabs(y[ny - 2] - y[ny - 1]));
	c->meshh = sqrt(d[0] * d[0] + d[1] * d[1]);
	return 0;
}

/* ------------------------------------------------------------------ indices (ref: src/matbuild.c) */
static int ipow2(int e) { int r = 1; while (e-- > 0) r *= 2; return r; }

/* ref: src/matbuild.c:85-105 (SetUpIndices), :49-72 (sizes), :280-309 (natural numbering; the other two
 * styles collapse to it for one grid per level).  Extension: map 3 = red-black numbering. */
static int oracle_indices(MgoCtx *c)
{
	c->lev = calloc((size_t)c->levels, sizeof(OLevel));
	for (int l = 0; l < c->levels; l++) {
		OLevel *L = &c->lev[l];
		const int f = ipow2(l);
		const int n0 = (c->n[0] - 1) / f - 1;     /* x count */
		const int n1 = (c->n[1] - 1) / f - 1;     /* y count */
		if (n0 < 1 || n1 < 1) { MGO_ERR("level %d has no interior points (npts=%d)", l, c->npts); return 1; }
		L->ni = n1; L->nj = n0; L->gridId = l;
		L->h[0] = 1.0 / (L->ni + 1);
		L->h[1] = 1.0 / (L->nj + 1);
		const int N = L->ni * L->nj;
		L->g2G = malloc((size_t)N * sizeof(int));
		L->G2g = malloc((size_t)N * 3 * sizeof(int));
		if (!L->g2G || !L->G2g) { MGO_ERR("out of memory for index maps"); return 1; }
		int count = 0;
		if (c->map == 3) {
			for (int colour = 0; colour < 2; colour++)
				for (int i = 0; i < L->ni; i++)
					for (int j = 0; j < L->nj; j++) {
						if (((i + j) & 1) != colour) continue;
						L->g2G[i * L->nj + j] = count;
						L->G2g[3 * count] = i; L->G2g[3 * count + 1] = j; L->G2g[3 * count + 2] = l;
						count++;
					}
		} else {
			for (int i = 0; i < L->ni; i++)
				for (int j = 0; j < L->nj; j++) {
					L->g2G[i * L->nj + j] = count;
					L->G2g[3 * count] = i; L->G2g[3 * count + 1] = j; L->G2g[3 * count + 2] = l;
					count++;
				}
		}
	}
	return 0;
}

/* ref: src/matbuild.c:398-407 and :422-431 -- 3x3 bilinear prolongation / full-weighting restriction */
static void oracle_stencils(MgoCtx *c)
{
	for (int i = 0; i < 3; i++) {
		const double a = fabs((double)(1 - i));
		c->pro3[i * 3 + 0] = 0.5 - 0.25 * a;  c->pro3[i * 3 + 1] = 1.0 - 0.5 * a;   c->pro3[i * 3 + 2] = 0.5 - 0.25 * a;
		c->res3[i * 3 + 0] = 0.125 - 0.0625 * a; c->res3[i * 3 + 1] = 0.25 - 0.125 * a; c->res3[i * 3 + 2] = 0.125 - 0.0625 * a;
	}
}

/* ------------------------------------------------------------------ assembly (ref: src/solver.c) */
/* ref: src/solver.c:489-510 (levelMatrixA) + :185-253 (fillJacobians) */
static void oracle_level_matrix(MgoCtx *c, int l)
{
	OLevel *L = &c->lev[l];
	const int N = L->ni * L->nj;
	MatCreateAIJ(PETSC_COMM_WORLD, N, N, PETSC_DETERMINE, PETSC_DETERMINE, 6, PETSC_NULL, 6, PETSC_NULL, &c->A[l]);
	const int f = ipow2(L->gridId);
	for (int row = 0; row < N; row++) {
		const int i0 = L->G2g[3 * row], j0 = L->G2g[3 * row + 1];
		const int ifine = f * (i0 + 1) - 1, jfine = f * (j0 + 1) - 1;
		double met[5], As[5];
		oracle_metrics(c, c->coord[0][jfine + 1], c->coord[1][ifine + 1], met);
		oracle_OpA(As, met, L->h);
		if (i0 - 1 >= 0)    MatSetValue(c->A[l], row, L->g2G[(i0 - 1) * L->nj + j0], As[0], ADD_VALUES);
		if (j0 - 1 >= 0)    MatSetValue(c->A[l], row, L->g2G[i0 * L->nj + j0 - 1], As[1], ADD_VALUES);
		MatSetValue(c->A[l], row, row, As[2], ADD_VALUES);
		if (j0 + 1 < L->nj) MatSetValue(c->A[l], row, L->g2G[i0 * L->nj + j0 + 1], As[3], ADD_VALUES);
		if (i0 + 1 < L->ni) MatSetValue(c->A[l], row, L->g2G[(i0 + 1) * L->nj + j0], As[4], ADD_VALUES);
	}
	MatAssemblyBegin(c->A[l], MAT_FINAL_ASSEMBLY);
	MatAssemblyEnd(c->A[l], MAT_FINAL_ASSEMBLY);
}

/* ref: src/solver.c:558-620 (levelvecb), fine level only: b[row] = F(x_{j+1}, y_{i+1}) */
static void oracle_rhs(MgoCtx *c)
{
	OLevel *L = &c->lev[0];
	const int N = L->ni * L->nj;
	for (int row = 0; row < N; row++) {
		const int i0 = L->G2g[3 * row], j0 = L->G2g[3 * row + 1];
		VecSetValue(c->b[0], row, oracle_F(c->coord[0][j0 + 1], c->coord[1][i0 + 1]), INSERT_VALUES);
	}
	VecAssemblyBegin(c->b[0]); VecAssemblyEnd(c->b[0]);
}

/* ref: src/solver.c:1035-1094 (Res) and :1096-1154 (Pro); 3x3 stencil centred on fine (2*i1+1, 2*j1+1) */
static void oracle_transfer(MgoCtx *c, int l)
{
	OLevel *F = &c->lev[l], *C = &c->lev[l + 1];
	const int NF = F->ni * F->nj, NC = C->ni * C->nj;
	MatCreateAIJ(PETSC_COMM_WORLD, NC, NF, PETSC_DETERMINE, PETSC_DETERMINE, 9, PETSC_NULL, 9, PETSC_NULL, &c->res[l]);
	MatCreateAIJ(PETSC_COMM_WORLD, NF, NC, PETSC_DETERMINE, PETSC_DETERMINE, 4, PETSC_NULL, 4, PETSC_NULL, &c->pro[l]);
	for (int row = 0; row < NC; row++) {
		const int i1 = C->G2g[3 * row], j1 = C->G2g[3 * row + 1];
		const int i0 = 2 * (i1 + 1) - 1 - 3 / 2, j0 = 2 * (j1 + 1) - 1 - 3 / 2;
		for (int i = i0; i < i0 + 3; i++)
			for (int j = j0; j < j0 + 3; j++) {
				const double wr = c->res3[(i - i0) * 3 + (j - j0)];
				const double wp = c->pro3[(i - i0) * 3 + (j - j0)];
				if (wr != 0.0) MatSetValue(c->res[l], row, F->g2G[i * F->nj + j], wr, ADD_VALUES);
				if (wp != 0.0) MatSetValue(c->pro[l], F->g2G[i * F->nj + j], row, wp, ADD_VALUES);
			}
	}
	MatAssemblyBegin(c->res[l], MAT_FINAL_ASSEMBLY); MatAssemblyEnd(c->res[l], MAT_FINAL_ASSEMBLY);
	MatAssemblyBegin(c->pro[l], MAT_FINAL_ASSEMBLY); MatAssemblyEnd(c->pro[l], MAT_FINAL_ASSEMBLY);
}

/* ref: src/solver.c:1156-1209 (Assemble), non-delayed cycles */
static void oracle_assemble(MgoCtx *c)
{
	c->A = calloc((size_t)c->levels, sizeof(Mat));
	c->b = calloc((size_t)c->levels, sizeof(Vec));
	c->u = calloc((size_t)c->levels, sizeof(Vec));
	c->res = calloc((size_t)c->levels, sizeof(Mat));
	c->pro = calloc((size_t)c->levels, sizeof(Mat));
	for (int l = 0; l < c->levels; l++) {
		oracle_level_matrix(c, l);
		MatCreateVecs(c->A[l], &c->u[l], &c->b[l]);
	}
	oracle_rhs(c);
	for (int l = 0; l + 1 < c->levels; l++) oracle_transfer(c, l);
}

/* ------------------------------------------------------------------ create / destroy */
MgoCtx *mgo_create(const char *options)
{
	MiniPetscSetQuiet(1);
	PetscOptionsClear(NULL);
	PetscOptionsInsertString(NULL, options ? options : "");
	MgoCtx *c = calloc(1, sizeof *c);
	/* ref: src/poisson.c:51-59 -- the reference has NO defaults; the oracle insists on the keys it needs */
	PetscBool s1, s2, s3, s4; PetscInt vmax = 2;
	c->meshflag = 0; c->cycle = 0; c->map = 2; c->moreNorm = 0; c->v[0] = 3; c->v[1] = 3; c->rtol = 1e-7;
	PetscOptionsGetInt(NULL, NULL, "-npts", &c->npts, &s1);
	PetscOptionsGetInt(NULL, NULL, "-mesh", &c->meshflag, NULL);
	PetscOptionsGetInt(NULL, NULL, "-iter", &c->maxIter, &s2);
	PetscOptionsGetInt(NULL, NULL, "-grids", &c->grids, &s3);
	PetscOptionsGetInt(NULL, NULL, "-levels", &c->levels, &s4);
	PetscOptionsGetInt(NULL, NULL, "-cycle", &c->cycle, NULL);
	PetscOptionsGetInt(NULL, NULL, "-map", &c->map, NULL);
	PetscOptionsGetIntArray(NULL, NULL, "-v", c->v, &vmax, NULL);
	PetscOptionsGetInt(NULL, NULL, "-moreNorm", &c->moreNorm, NULL);
	PetscOptionsGetReal(NULL, NULL, "-rtol", &c->rtol, NULL);
	PetscInt thr = 0; PetscOptionsGetInt(NULL, NULL, "-threads", &thr, NULL); MiniPetscSetThreads(thr);
	if (!s1 || !s2 || !s4) { MGO_ERR("options -npts, -iter and -levels are required"); free(c); return NULL; }
	if (!s3) c->grids = c->levels;
	if (c->grids != c->levels) { MGO_ERR("only one grid per level is restated (-grids must equal -levels)"); free(c); return NULL; }
	if (c->cycle != 0 && c->cycle != 8) { MGO_ERR("only cycle 0 (V-cycle) and 8 (PCMG) are restated"); free(c); return NULL; }
	if (c->meshflag < 0 || c->meshflag > 2 || c->map < 0 || c->map > 3 || c->levels < 1 || c->maxIter < 0) { MGO_ERR("invalid option value"); free(c); return NULL; }
	/* ref: src/poisson.c:73-82 -- square grid on the unit square */
	c->n[0] = c->npts; c->n[1] = c->npts;
	c->bounds[0] = 0.0; c->bounds[1] = 1.0; c->bounds[2] = 0.0; c->bounds[3] = 1.0;
	c->coord[0] = malloc((size_t)c->n[0] * sizeof(double));
	c->coord[1] = malloc((size_t)c->n[1] * sizeof(double));
	if (oracle_coords(c) || oracle_indices(c)) { mgo_destroy(c); return NULL; }
	oracle_stencils(c);
	/* ref: src/solver.c:107-135 (SetUpSolver): rnorm has numIter+1 entries */
	c->rnorm = calloc((size_t)c->maxIter + 2, sizeof(double));
	oracle_assemble(c);
	return c;
}

void mgo_destroy(MgoCtx *c)
{
	if (!c) return;
	if (c->smoother) { for (int l = 0; l < c->levels; l++) KSPDestroy(&c->smoother[l]); free(c->smoother); }
	for (int l = 0; l < c->levels; l++) {
		if (c->A) MatDestroy(&c->A[l]);
		if (c->res) MatDestroy(&c->res[l]);
		if (c->pro) MatDestroy(&c->pro[l]);
		if (c->b) VecDestroy(&c->b[l]);
		if (c->u) VecDestroy(&c->u[l]);
		if (c->lev) { free(c->lev[l].g2G); free(c->lev[l].G2g); }
	}
	free(c->A); free(c->res); free(c->pro); free(c->b); free(c->u); free(c->lev);
	free(c->coord[0]); free(c->coord[1]); free(c->rnorm);
	free(c);
}

/* ------------------------------------------------------------------ accessors */
int mgo_levels(const MgoCtx *c) { return c->levels; }
int mgo_level_dims(const MgoCtx *c, int l, int *ni, int *nj)
{ if (l < 0 || l >= c->levels) return 1; *ni = c->lev[l].ni; *nj = c->lev[l].nj; return 0; }
int mgo_level_h(const MgoCtx *c, int l, double h[2])
{ if (l < 0 || l >= c->levels) return 1; h[0] = c->lev[l].h[0]; h[1] = c->lev[l].h[1]; return 0; }
int mgo_coords(const MgoCtx *c, int dim, double *out)
{ if (dim < 0 || dim > 1) return 1; memcpy(out, c->coord[dim], (size_t)c->n[dim] * sizeof(double)); return 0; }
int mgo_stencil(const MgoCtx *c, int which, double out[9])
{ memcpy(out, which == 0 ? c->res3 : c->pro3, 9 * sizeof(double)); return 0; }
int mgo_opA(const MgoCtx *c, int l, int i, int j, double As[5])
{
	if (l < 0 || l >= c->levels) return 1;
	const int f = ipow2(c->lev[l].gridId);
	double met[5];
	oracle_metrics(c, c->coord[0][f * (j + 1) - 1 + 1], c->coord[1][f * (i + 1) - 1 + 1], met);
	oracle_OpA(As, met, c->lev[l].h);
	return 0;
}
int mgo_grid_to_global(const MgoCtx *c, int l, int *out)
{ if (l < 0 || l >= c->levels) return 1; memcpy(out, c->lev[l].g2G, (size_t)c->lev[l].ni * c->lev[l].nj * sizeof(int)); return 0; }
int mgo_global_to_grid(const MgoCtx *c, int l, int *out)
{ if (l < 0 || l >= c->levels) return 1; memcpy(out, c->lev[l].G2g, (size_t)c->lev[l].ni * c->lev[l].nj * 3 * sizeof(int)); return 0; }

static Mat pick_mat(const MgoCtx *c, int which, int l)
{
	if (l < 0 || l >= c->levels) return NULL;
	if (which == 0) return c->A[l];
	if (l >= c->levels - 1) return NULL;
	return which == 1 ? c->res[l] : (which == 2 ? c->pro[l] : NULL);
}
int mgo_csr_dims(const MgoCtx *c, int which, int l, int *m, int *n, int *nnz)
{
	Mat M = pick_mat(c, which, l); if (!M) return 1;
	const PetscInt *ia; MatSeqAIJGetCSR(M, m, n, &ia, NULL, NULL); *nnz = ia[*m];
	return 0;
}
int mgo_csr_copy(const MgoCtx *c, int which, int l, int *ia_out, int *ja_out, double *va_out)
{
	Mat M = pick_mat(c, which, l); if (!M) return 1;
	PetscInt m, n; const PetscInt *ia, *ja; const PetscScalar *va;
	MatSeqAIJGetCSR(M, &m, &n, &ia, &ja, &va);
	memcpy(ia_out, ia, ((size_t)m + 1) * sizeof(int));
	memcpy(ja_out, ja, (size_t)ia[m] * sizeof(int));
	memcpy(va_out, va, (size_t)ia[m] * sizeof(double));
	return 0;
}
int mgo_vec_get(const MgoCtx *c, int which, int l, double *out)
{
	if (l < 0 || l >= c->levels) return 1;
	Vec v = which == 0 ? c->b[l] : c->u[l]; PetscScalar *a; PetscInt n;
	VecGetArray(v, &a); VecGetSize(v, &n); memcpy(out, a, (size_t)n * sizeof(double));
	return 0;
}
int mgo_vec_set(MgoCtx *c, int which, int l, const double *in)
{
	if (l < 0 || l >= c->levels) return 1;
	Vec v = which == 0 ? c->b[l] : c->u[l]; PetscScalar *a; PetscInt n;
	VecGetArray(v, &a); VecGetSize(v, &n); memcpy(a, in, (size_t)n * sizeof(double));
	return 0;
}

/* wrap caller memory in temporary Vecs */
static Vec vec_from(const double *p, int n) { Vec v; VecCreateSeq(0, n, &v); PetscScalar *a; VecGetArray(v, &a); if (p) memcpy(a, p, (size_t)n * sizeof(double)); return v; }
static void vec_to(Vec v, double *p) { PetscScalar *a; PetscInt n; VecGetArray(v, &a); VecGetSize(v, &n); memcpy(p, a, (size_t)n * sizeof(double)); }

int mgo_matmult(const MgoCtx *c, int which, int l, const double *x, double *y)
{
	Mat M = pick_mat(c, which, l); if (!M) return 1;
	PetscInt m, n; MatGetSize(M, &m, &n);
	Vec vx = vec_from(x, n), vy = vec_from(NULL, m);
	MatMult(M, vx, vy); vec_to(vy, y);
	VecDestroy(&vx); VecDestroy(&vy); return 0;
}
int mgo_matmultadd(const MgoCtx *c, int which, int l, const double *x, const double *y, double *z)
{
	Mat M = pick_mat(c, which, l); if (!M) return 1;
	PetscInt m, n; MatGetSize(M, &m, &n);
	Vec vx = vec_from(x, n), vy = vec_from(y, m), vz = vec_from(NULL, m);
	MatMultAdd(M, vx, vy, vz); vec_to(vz, z);
	VecDestroy(&vx); VecDestroy(&vy); VecDestroy(&vz); return 0;
}
int mgo_residual(const MgoCtx *c, int l, const double *b, const double *x, double *r)
{
	Mat M = pick_mat(c, 0, l); if (!M) return 1;
	PetscInt m, n; MatGetSize(M, &m, &n);
	Vec vb = vec_from(b, m), vx = vec_from(x, n), vr = vec_from(NULL, m);
	MatResidual(M, vb, vx, vr); vec_to(vr, r);
	VecDestroy(&vb); VecDestroy(&vx); VecDestroy(&vr); return 0;
}
double mgo_norm2(const double *x, int n) { Vec v = vec_from(x, n); PetscReal r; VecNorm(v, NORM_2, &r); VecDestroy(&v); return r; }
double mgo_dot(const double *x, const double *y, int n)
{ Vec a = vec_from(x, n), b = vec_from(y, n); PetscScalar r; VecDot(a, b, &r); VecDestroy(&a); VecDestroy(&b); return r; }

/* the level smoother as configured at ref: src/solver.c:1463-1510 */
static KSP oracle_make_smoother(MgoCtx *c, int l, int max_it)
{
	KSP k;
	KSPCreate(PETSC_COMM_WORLD, &k);
	KSPSetType(k, KSPRICHARDSON);
	KSPSetOperators(k, c->A[l], c->A[l]);
	KSPSetNormType(k, KSP_NORM_NONE);
	KSPSetTolerances(k, 1.e-7, PETSC_DEFAULT, PETSC_DEFAULT, max_it);
	KSPSetFromOptions(k);
	return k;
}
int mgo_smooth(MgoCtx *c, int l, const double *b, double *x, int nu, int guess_zero)
{
	if (l < 0 || l >= c->levels) return 1;
	if (!c->smoother) c->smoother = calloc((size_t)c->levels, sizeof(KSP));
	if (!c->smoother[l]) c->smoother[l] = oracle_make_smoother(c, l, nu);
	KSP k = c->smoother[l];
	KSPSetTolerances(k, PETSC_DEFAULT, PETSC_DEFAULT, PETSC_DEFAULT, nu);
	KSPSetInitialGuessNonzero(k, guess_zero ? PETSC_FALSE : PETSC_TRUE);
	const int N = c->lev[l].ni * c->lev[l].nj;
	Vec vb = vec_from(b, N), vx = vec_from(x, N);
	KSPSolve(k, vb, vx); vec_to(vx, x);
	VecDestroy(&vb); VecDestroy(&vx); return 0;
}

/* ------------------------------------------------------------------ cycle 0 (ref: src/solver.c:1414-1575) */
static void oracle_vcycle(MgoCtx *c)
{
	const int levels = c->levels, maxIter = c->maxIter, *v = c->v;
	Mat *A = c->A, *res = c->res, *pro = c->pro; Vec *b = c->b, *u = c->u;
	KSP *ksp = calloc((size_t)levels, sizeof(KSP));
	Vec *r = calloc((size_t)levels, sizeof(Vec)), *rv = calloc((size_t)levels, sizeof(Vec));
	double *rnorm = c->rnorm, rnormchk, bnorm;
	int iter;

	for (int i = 0; i < levels; i++) VecDuplicate(b[i], &rv[i]);
	/* :1463-1510 -- level 0 .. L-2 get v[0] sweeps, the coarsest v[1] (level 0 keeps v[0] when L == 1) */
	for (int i = 0; i < levels; i++) ksp[i] = oracle_make_smoother(c, i, (i == levels - 1 && levels > 1) ? v[1] : v[0]);

	VecNorm(b[0], NORM_2, &bnorm);                                       /* :1512 */
	VecSet(u[0], 0.0);                                                   /* :1514 */
	MatMult(A[0], u[0], rv[0]);                                          /* :1516 */
	VecAXPY(rv[0], -1.0, b[0]);                                          /* :1517 */
	VecNorm(rv[0], NORM_2, &rnormchk);                                   /* :1518 */
	rnorm[0] = rnormchk;
	iter = 0;
	const double t0 = MPI_Wtime();
	while (iter < maxIter && 100000000 * bnorm > rnormchk && rnormchk > (c->rtol) * bnorm) {   /* :1530 */
		KSPSolve(ksp[0], b[0], u[0]);
		if (iter == 0) KSPSetInitialGuessNonzero(ksp[0], PETSC_TRUE);
		for (int l = 1; l < levels; l++) {
			KSPBuildResidual(ksp[l - 1], NULL, rv[l - 1], &r[l - 1]);
			MatMult(res[l - 1], r[l - 1], b[l]);
			KSPSolve(ksp[l], b[l], u[l]);
			if (l != levels - 1) KSPSetInitialGuessNonzero(ksp[l], PETSC_TRUE);
		}
		for (int l = levels - 2; l >= 0; l = l - 1) {
			MatMult(pro[l], u[l + 1], rv[l]);
			VecAXPY(u[l], 1.0, rv[l]);
			KSPSolve(ksp[l], b[l], u[l]);
			if (l != 0) KSPSetInitialGuessNonzero(ksp[l], PETSC_FALSE);
		}
		KSPBuildResidual(ksp[0], NULL, rv[0], &r[0]);
		VecNorm(r[0], NORM_2, &rnormchk);
		iter = iter + 1;
		rnorm[iter] = rnormchk;
	}
	c->solve_seconds = MPI_Wtime() - t0;
	/* :1554-1558 -- normalise by rnorm[0]; the reference also divides the uninitialised tail, we stop at iter */
	rnormchk = rnorm[0];
	for (int i = 0; i <= iter; i++) rnorm[i] = rnorm[i] / rnormchk;
	c->numIter = iter;
	for (int i = 0; i < levels; i++) { VecDestroy(&rv[i]); KSPDestroy(&ksp[i]); }
	free(ksp); free(r); free(rv);
}

/* ------------------------------------------------------------------ cycle 8 (ref: src/solver.c:1884-1989) */
static void oracle_pcmg(MgoCtx *c)
{
	const int levels = c->levels;
	Mat *A = c->A, *res = c->res, *pro = c->pro; Vec *b = c->b, *u = c->u;
	KSP ksp, kt; PC pc;
	KSPCreate(PETSC_COMM_WORLD, &ksp);
	KSPSetType(ksp, KSPRICHARDSON);
	KSPSetOperators(ksp, A[0], A[0]);
	KSPSetNormType(ksp, KSP_NORM_UNPRECONDITIONED);
	KSPSetResidualHistory(ksp, c->rnorm, c->maxIter, PETSC_FALSE);
	KSPSetTolerances(ksp, 1.e-7, PETSC_DEFAULT, PETSC_DEFAULT, c->maxIter);
	KSPGetPC(ksp, &pc);
	PCSetType(pc, PCMG);
	PCMGSetLevels(pc, levels, NULL);
	PCMGGetCoarseSolve(pc, &kt);
	KSPSetOperators(kt, A[levels - 1], A[levels - 1]);
	for (int i = 1; i < levels; i++) {                      /* PETSc numbers levels coarse -> fine */
		PCMGGetSmoother(pc, i, &kt);
		KSPSetOperators(kt, A[levels - i - 1], A[levels - i - 1]);
		PCMGSetInterpolation(pc, i, pro[levels - i - 1]);
		PCMGSetRestriction(pc, i, res[levels - i - 1]);
	}
	Vec *r = calloc((size_t)levels, sizeof(Vec));
	for (int i = 0; i < levels; i++) VecDuplicate(b[i], &r[i]);
	PCMGSetR(pc, levels - 1, r[0]);
	for (int i = 1; i < levels - 1; i++) {
		PCMGSetRhs(pc, i, b[levels - i - 1]);
		PCMGSetX(pc, i, u[levels - i - 1]);
		PCMGSetR(pc, i, r[levels - i - 1]);
	}
	PCMGSetRhs(pc, 0, b[levels - 1]);
	PCMGSetX(pc, 0, u[levels - 1]);
	KSPSetFromOptions(ksp);

	for (int i = 0; i <= c->maxIter; i++) c->rnorm[i] = NAN;  /* the reference leaves these uninitialised */
	const double t0 = MPI_Wtime();
	KSPSolve(ksp, b[0], u[0]);
	c->solve_seconds = MPI_Wtime() - t0;
	KSPGetIterationNumber(ksp, &c->numIter);
	const double rnorm0 = c->rnorm[0];
	for (int i = 0; i < c->numIter + 1 && i <= c->maxIter; i++) c->rnorm[i] = c->rnorm[i] / rnorm0;
	for (int i = 0; i < levels; i++) VecDestroy(&r[i]);
	free(r);
	KSPDestroy(&ksp);
}

/* ref: src/solver.c:2617-2630 */
int mgo_solve(MgoCtx *c)
{
	if (c->cycle == 0) oracle_vcycle(c);
	else if (c->cycle == 8) {
		if (c->levels < 2) { MGO_ERR("cycle 8 needs at least two levels"); return 1; }
		oracle_pcmg(c);
	} else return 1;
	c->solved = 1;
	return 0;
}
int mgo_num_iter(const MgoCtx *c) { return c->numIter; }
int mgo_rnorm(const MgoCtx *c, double *out, int nmax)
{
	int n = c->numIter + 1; if (n > nmax) n = nmax;
	memcpy(out, c->rnorm, (size_t)n * sizeof(double));
	return n;
}
double mgo_solve_seconds(const MgoCtx *c) { return c->solve_seconds; }

/* ------------------------------------------------------------------ post-processing */
/* ref: src/solver.c:1239-1315 (GetSol, one rank) + :1211-1237 (GetError) */
int mgo_postprocess(const MgoCtx *c, double *ug, double error[3])
{
	const OLevel *L = &c->lev[0];
	const int N = L->ni * L->nj;
	PetscScalar *px; VecGetArray(c->u[0], &px);
	for (int row = 0; row < N; row++) ug[L->G2g[3 * row] * L->nj + L->G2g[3 * row + 1]] = px[row];
	error[0] = 0.0; error[1] = 0.0; error[2] = 0.0;
	for (int i = 0; i < L->ni; i++)
		for (int j = 0; j < L->nj; j++) {
			const double sol = oracle_SOL(c->coord[0][j + 1], c->coord[1][i + 1]);
			const double diff = fabs(ug[i * L->nj + j] - sol);
			error[0] = fmax(diff, error[0]);
			error[1] = error[1] + diff;
			error[2] = error[2] + diff * diff;
		}
	error[2] = sqrt(error[2]);
	return 0;
}

/* ref: src/solver.c:1317-1380 (Postprocessing file formats; note X/YgridData print coord[.][j], not [j+1]) */
int mgo_write_files(const MgoCtx *c, const char *dir)
{
	const OLevel *L = &c->lev[0];
	double *ug = malloc((size_t)L->ni * L->nj * sizeof(double)), err[3];
	mgo_postprocess(c, ug, err);
	char path[1024]; FILE *fu, *fr, *fe, *fx, *fy;
	snprintf(path, sizeof path, "%s/uData.dat", dir); fu = fopen(path, "w");
	snprintf(path, sizeof path, "%s/rData.dat", dir); fr = fopen(path, "w");
	snprintf(path, sizeof path, "%s/eData.dat", dir); fe = fopen(path, "w");
	snprintf(path, sizeof path, "%s/XgridData.dat", dir); fx = fopen(path, "w");
	snprintf(path, sizeof path, "%s/YgridData.dat", dir); fy = fopen(path, "w");
	if (!fu || !fr || !fe || !fx || !fy) { free(ug); return 1; }
	for (int i = 0; i < 3; i++) fprintf(fe, "%.16e\n", err[i]);
	for (int i = 0; i < L->ni; i++) {
		for (int j = 0; j < L->nj; j++) {
			fprintf(fx, "%lf    ", c->coord[0][j]);
			fprintf(fy, "%lf    ", c->coord[1][i]);
			fprintf(fu, "%.16e    ", ug[i * L->nj + j]);
		}
		fprintf(fx, "\n"); fprintf(fy, "\n"); fprintf(fu, "\n");
	}
	for (int i = 0; i < c->numIter + 1; i++) fprintf(fr, "%.16e ", c->rnorm[i]);
	fprintf(fr, "\n");
	fclose(fu); fclose(fr); fclose(fe); fclose(fx); fclose(fy);
	free(ug);
	return 0;
}
