/*
 * pb_setup.c -- problem, mesh, index and transfer-stencil setup of the standalone B200 driver.
 *
 * Host-side, runs once, cheap.  Everything here feeds the engine tables that must be BIT-IDENTICAL to what the
 * reference computes (coordinates, metrics, stencil coefficients, transfer weights), so each expression keeps
 * the reference's order of floating-point operations (compile with -ffp-contract=off):
 *   problem    ref: src/problem.c:3-46
 *   mesh       ref: src/mesh.c:29-107 (metrics), :130-195 (coordinates), :236-256
 *   indices    ref: src/matbuild.c:27-118 (sizes, h) and :280-323 (numbering; with one grid per level all three
 *              styles are the natural row-major numbering)
 *   stencils   ref: src/matbuild.c:326-442
 * In a true drop-in build the reference's own problem.c / mesh.c / matbuild.c / array.c are used instead of
 * this file.
 *
 * Difference by design: the reference materialises an n^2 grid->global map and a 3n^2 global->grid map on every
 * rank (1.07 GB at 8193^2) because its assembly loop walks them.  The engine generates the operator in closed
 * form, so the maps are only built on request (option -pb_index_maps 1); their dimensions are always set
 * because PrintInfo reads them (ref: src/poisson.c:185-193).
 */
#include "pb_api.h"
#include <string.h>

/* ---------------------------------------------------------------- containers (ref: src/array.c:18-43) */
void CreateArrayInt2d(int ni, int nj, ArrayInt2d *a) { a->ni = ni; a->nj = nj; a->data = malloc((size_t)ni * nj * sizeof(int)); }
void DeleteArrayInt2d(ArrayInt2d *a) { free(a->data); a->data = NULL; }
void CreateArray2d(int ni, int nj, Array2d *a) { a->ni = ni; a->nj = nj; a->data = malloc((size_t)ni * nj * sizeof(double)); }
void DeleteArray2d(Array2d *a) { free(a->data); a->data = NULL; }

/* ---------------------------------------------------------------- problem */
/* coefficients (S, W, C, E, N) of the transformed Laplacian at one point; ref: src/problem.c:15-21 */
static void poisson_stencil(double *A, double *metrics, double *h)
{
	const double hx2 = h[0] * h[0];
	const double hy2 = h[1] * h[1];
	A[0] = (metrics[1] / hy2) - (metrics[3] / (2 * h[1]));
	A[1] = (metrics[0] / hx2) - (metrics[2] / (2 * h[0]));
	A[2] = -2.0 * ((metrics[0] / hx2) + (metrics[1] / hy2));
	A[3] = (metrics[0] / hx2) + (metrics[2] / (2 * h[0]));
	A[4] = (metrics[1] / hy2) + (metrics[3] / (2 * h[1]));
}
static double poisson_rhs(double x, double y) { return -2 * PI * PI * sin(PI * x) * sin(PI * y); }   /* :27 */
static double poisson_exact(double x, double y) { return sin(PI * x) * sin(PI * y); }                /* :33 */

void SetUpProblem(Problem *prob)
{
	prob->Ffunc = &poisson_rhs;
	prob->SOLfunc = &poisson_exact;
	prob->OpA = &poisson_stencil;
}

/* ---------------------------------------------------------------- mesh */
/* metrics[0..4] = xi_x^2+xi_y^2, eta_x^2+eta_y^2, lap(xi), lap(eta), cross term; all three depend on y only */
static void metrics_uniform(void *mesh, double x, double y, double *m)
{
	(void)mesh; (void)x; (void)y;
	m[0] = 1.0; m[1] = 1.0; m[2] = 0.0; m[3] = 0.0; m[4] = 0.0;
}
/* cosine-clustered y; ref: src/mesh.c:45-74 */
static void metrics_cosine(void *mesh, double x, double y, double *m)
{
	(void)x;
	const double *bd = ((Mesh *)mesh)->bounds;
	const double t = ((bd[3] - bd[2]) * (bd[3] - bd[2]) - (bd[3] - y) * (bd[3] - y));
	m[0] = 1.0;
	m[1] = 4.0 / (PI * PI * t);
	m[2] = 0.0;
	m[3] = (-2.0 * (bd[3] - y)) / (PI * sqrt(t * t * t));
	m[4] = 0.0;
}
/* exponentially stretched y; ref: src/mesh.c:76-107 */
static void metrics_exponential(void *mesh, double x, double y, double *m)
{
	(void)x;
	const double *bd = ((Mesh *)mesh)->bounds;
	const double t = ((exp(2) - 1) * (exp(2) - 1)) /
	                 (((y - bd[2]) * (exp(2) - 1) + (bd[3] - bd[2])) * ((y - bd[2]) * (exp(2) - 1) + (bd[3] - bd[2])));
	m[0] = 1.0 / ((bd[1] - bd[0]) * (bd[1] - bd[0]));
	m[1] = 0.25 * t;
	m[2] = 0.0;
	m[3] = (-0.5) * t;
	m[4] = 0.0;
}

void SetUpMesh(Mesh *mesh, MeshType type)
{
	const int nx = mesh->n[0], ny = mesh->n[1];
	mesh->coord = NULL;
	if (nx < 2 || ny < 2) { printf("ERROR: Need at least 2 points in each direction\n"); return; }
	/* one contiguous block, coord[0] = x, coord[1] = y (same ownership as the reference: free(coord[0]); free(coord)) */
	mesh->coord = malloc(2 * sizeof(double *));
	mesh->coord[0] = malloc((size_t)(nx + ny) * sizeof(double));
	mesh->coord[1] = mesh->coord[0] + nx;
	double *x = mesh->coord[0], *y = mesh->coord[1];

	/* x: uniform, built by repeated addition of the spacing (ref: src/mesh.c:146-152) */
	x[0] = mesh->bounds[0]; x[nx - 1] = mesh->bounds[1];
	const double dx = (x[nx - 1] - x[0]) / (nx - 1);
	for (int j = 1; j < nx - 1; j++) x[j] = x[j - 1] + dx;

	/* y: by mesh type (ref: src/mesh.c:154-174) */
	y[0] = mesh->bounds[2]; y[ny - 1] = mesh->bounds[3];
	const double len = y[ny - 1] - y[0];
	const double step = len / (double)(ny - 1);
	double dymax = 0.0;
	for (int j = 1; j < ny - 1; j++) {
		const double eta = j / (double)(ny - 1);
		if (type == NONUNIFORM1) y[j] = mesh->bounds[3] - len * (cos(PI * 0.5 * eta));
		else if (type == NONUNIFORM2) y[j] = mesh->bounds[2] + len * ((exp(2 * eta) - 1) / (exp(2) - 1));
		else y[j] = y[j - 1] + step;
		dymax = fmax(dymax, fabs(y[j] - y[j - 1]));
	}
	dymax = fmax(dymax, fabs(y[ny - 2] - y[ny - 1]));
	mesh->h = sqrt(dx * dx + dymax * dymax);

	mesh->MetricCoefficients = (type == UNIFORM) ? &metrics_uniform : (type == NONUNIFORM1) ? &metrics_cosine : &metrics_exponential;
}

void DestroyMesh(Mesh *mesh)
{
	if (mesh->coord) { free(mesh->coord[0]); free(mesh->coord); mesh->coord = NULL; }
}

/* ---------------------------------------------------------------- indices */
static int int_pow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

void SetUpIndices(Mesh *mesh, Indices *indices)
{
	const int L = indices->levels, f = indices->coarseningFactor;
	indices->level = calloc((size_t)L, sizeof(Level));
	int next_id = 0;
	for (int l = 0; l < L; l++) {
		Level *lv = &indices->level[l];
		/* one grid per level; surplus grids all land on the last level (ref: src/matbuild.c:33-38) */
		lv->grids = 1 + ((l == L - 1) ? indices->totalGrids - L : 0);
		if (lv->grids < 1) lv->grids = 1;
		lv->gridId = malloc((size_t)lv->grids * sizeof(int));
		lv->grid = calloc((size_t)lv->grids, sizeof(ArrayInt2d));
		lv->h = malloc((size_t)lv->grids * sizeof(double[2]));
		lv->ranges = malloc(2 * sizeof(int));
		int total = 0;
		for (int g = 0; g < lv->grids; g++) {
			lv->gridId[g] = next_id++;
			const int c = int_pow(f, lv->gridId[g]);
			const int ncols = (mesh->n[0] - 1) / c - 1;       /* x count */
			const int nrows = (mesh->n[1] - 1) / c - 1;       /* y count */
			lv->grid[g].ni = nrows; lv->grid[g].nj = ncols; lv->grid[g].data = NULL;
			lv->h[g][0] = 1.0 / (lv->grid[g].ni + 1);         /* ref: src/matbuild.c:101-102 */
			lv->h[g][1] = 1.0 / (lv->grid[g].nj + 1);
			total += nrows * ncols;
		}
		lv->global.ni = total; lv->global.nj = 3; lv->global.data = NULL;
		lv->ranges[0] = 0; lv->ranges[1] = total;             /* one host process */
	}
}

void DestroyIndices(Indices *indices)
{
	for (int l = 0; l < indices->levels; l++) {
		Level *lv = &indices->level[l];
		for (int g = 0; g < lv->grids; g++) free(lv->grid[g].data);
		free(lv->global.data); free(lv->grid); free(lv->h); free(lv->gridId); free(lv->ranges);
	}
	free(indices->level);
	indices->level = NULL;
}

/* Numbering styles 0,1,2 of the reference coincide (natural row-major) when each level holds one grid
 * (ref: src/matbuild.c:146-323); style 3 is this build's red-black extension (reds, (i+j) even, first).
 * The maps are only materialised when -pb_index_maps 1 is given. */
void mapping(Indices *indices, int style)
{
	int want = 0;
	pbopt_get_int("-pb_index_maps", &want);
	if (!want) return;
	for (int l = 0; l < indices->levels; l++) {
		Level *lv = &indices->level[l];
		if (lv->grids != 1) continue;
		ArrayInt2d *g = &lv->grid[0];
		if (!g->data) g->data = malloc((size_t)g->ni * g->nj * sizeof(int));
		if (!lv->global.data) lv->global.data = malloc((size_t)lv->global.ni * 3 * sizeof(int));
		int count = 0;
		for (int pass = 0; pass < (style == 3 ? 2 : 1); pass++)
			for (int i = 0; i < g->ni; i++)
				for (int j = 0; j < g->nj; j++) {
					if (style == 3 && ((i + j) & 1) != pass) continue;
					g->data[i * g->nj + j] = count;
					lv->global.data[3 * count] = i; lv->global.data[3 * count + 1] = j; lv->global.data[3 * count + 2] = lv->gridId[0];
					count++;
				}
	}
}

/* ---------------------------------------------------------------- transfer stencils */
void SetUpOperator(Indices *indices, Operator *op)
{
	op->totalGrids = indices->totalGrids;
	const int n = op->totalGrids - 1 > 0 ? op->totalGrids - 1 : 0;
	op->res = malloc((size_t)(n ? n : 1) * sizeof(Array2d));
	op->pro = malloc((size_t)(n ? n : 1) * sizeof(Array2d));
	int size = 1;
	for (int k = 0; k < n; k++) {
		size = (size + 1) * indices->coarseningFactor - 1;         /* 3, 7, 15, ... (ref: src/matbuild.c:337) */
		CreateArray2d(size, size, &op->res[k]);
		CreateArray2d(size, size, &op->pro[k]);
		memset(op->res[k].data, 0, (size_t)size * size * sizeof(double));
		memset(op->pro[k].data, 0, (size_t)size * size * sizeof(double));
	}
}

void DestroyOperator(Operator *op)
{
	for (int k = 0; k < op->totalGrids - 1; k++) { DeleteArray2d(&op->res[k]); DeleteArray2d(&op->pro[k]); }
	free(op->res); free(op->pro);
}

/* 3x3 bilinear prolongation and full-weighting restriction (ref: src/matbuild.c:398-431).  Only the 3x3 pair
 * op.res[0] / op.pro[0] is used when every level holds one grid; the composite (7x7, ...) stencils of the
 * reference serve its several-grids-per-level research cycles and are left zero here. */
void GridTransferOperators(Operator op, Indices indices)
{
	(void)indices;
	if (op.totalGrids < 2) return;
	for (int a = 0; a < 3; a++) {
		const double d = fabs((double)(1 - a));
		op.pro[0].data[a * 3 + 0] = 0.5 - 0.25 * d;
		op.pro[0].data[a * 3 + 1] = 1.0 - 0.5 * d;
		op.pro[0].data[a * 3 + 2] = 0.5 - 0.25 * d;
		op.res[0].data[a * 3 + 0] = 0.125 - 0.0625 * d;
		op.res[0].data[a * 3 + 1] = 0.25 - 0.125 * d;
		op.res[0].data[a * 3 + 2] = 0.125 - 0.0625 * d;
	}
}
