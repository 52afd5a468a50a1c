/*
 * minipetsc.c -- TEST INFRASTRUCTURE ONLY (part of oracle/).  See petscksp.h.
 *
 * Single-rank CPU restatement of the PETSc subset the reference calls.  All
 * statements about what PETSc does are [PETSc-upstream]: recalled from the
 * published PETSc sources (src/mat/impls/aij/seq/aij.c, src/ksp/ksp/impls/
 * {rich,cg}, src/ksp/pc/impls/{jacobi,sor,mg,factor}), version unpinned,
 * not verifiable in this image.  PARITY UNPINNED against real PETSc.
 *
 * Arithmetic convention: IEEE-754 binary64, round-to-nearest-even, no FMA
 * contraction (compile with -ffp-contract=off), PETSc's operation order.
 * Reductions (VecNorm/VecDot) are blocked: fixed 4096-element blocks summed
 * left to right, then the block sums summed left to right -- deterministic
 * and independent of the thread count.
 */
#include "petscksp.h"
#include <stdarg.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MP_ERR(...) do { fprintf(stderr, "minipetsc error: " __VA_ARGS__); fprintf(stderr, "\n"); exit(77); } while (0)

static int g_quiet   = 0;
static int g_threads = 0;
void MiniPetscSetQuiet(int q) { g_quiet = q; }
void MiniPetscSetThreads(int n) { g_threads = n; }
int  MiniPetscGetThreads(void)
{
#ifdef _OPENMP
	return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
	return 1;
#endif
}
#ifdef _OPENMP
#define MP_PARFOR _Pragma("omp parallel for schedule(static) num_threads(MiniPetscGetThreads()) if (n_par > 32768)")
#else
#define MP_PARFOR
#endif

/* =========================================================================
 * 1-rank MPI
 * ========================================================================= */
int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
double MPI_Wtime(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; MP_ERR("MPI_Send on a 1-rank communicator"); return 1; }
int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; MP_ERR("MPI_Recv on a 1-rank communicator"); return 1; }

/* =========================================================================
 * Options database: "-key [value]" pairs; file first, then argv (argv wins
 * because later insertions overwrite).  '#' starts a comment in files.
 * ========================================================================= */
typedef struct { char *key; char *val; } OptEntry;
static OptEntry *g_opt = NULL;
static int g_nopt = 0, g_capopt = 0;

static void opt_set(const char *key, const char *val)
{
	for (int i = 0; i < g_nopt; i++) {
		if (strcmp(g_opt[i].key, key) == 0) {
			free(g_opt[i].val);
			g_opt[i].val = val ? strdup(val) : NULL;
			return;
		}
	}
	if (g_nopt == g_capopt) {
		g_capopt = g_capopt ? 2 * g_capopt : 32;
		g_opt = realloc(g_opt, (size_t)g_capopt * sizeof(OptEntry));
	}
	g_opt[g_nopt].key = strdup(key);
	g_opt[g_nopt].val = val ? strdup(val) : NULL;
	g_nopt++;
}

static int is_key(const char *tok)
{
	/* "-name" is a key; "-1.5" / "-3" are values */
	return tok[0] == '-' && tok[1] != '\0' && !((tok[1] >= '0' && tok[1] <= '9') || tok[1] == '.');
}

static void opt_insert_tokens(char **tok, int n)
{
	for (int i = 0; i < n; i++) {
		if (!is_key(tok[i])) continue;
		if (i + 1 < n && !is_key(tok[i + 1])) { opt_set(tok[i] + 1, tok[i + 1]); i++; }
		else opt_set(tok[i] + 1, NULL);
	}
}

PetscErrorCode PetscOptionsInsertString(void *options, const char *str)
{
	(void)options;
	char *copy = strdup(str);
	char **tok = NULL; int n = 0, cap = 0;
	for (char *p = strtok(copy, " \t\r\n"); p; p = strtok(NULL, " \t\r\n")) {
		if (n == cap) { cap = cap ? 2 * cap : 16; tok = realloc(tok, (size_t)cap * sizeof(char *)); }
		tok[n++] = p;
	}
	opt_insert_tokens(tok, n);
	free(tok); free(copy);
	return 0;
}

static void opt_insert_file(const char *file)
{
	FILE *f = fopen(file, "r");
	if (!f) return; /* PETSc silently ignores a missing default options file */
	char line[4096];
	while (fgets(line, sizeof line, f)) {
		char *hash = strchr(line, '#');
		if (hash) *hash = '\0';
		PetscOptionsInsertString(NULL, line);
	}
	fclose(f);
}

PetscErrorCode PetscOptionsClear(void *options)
{
	(void)options;
	for (int i = 0; i < g_nopt; i++) { free(g_opt[i].key); free(g_opt[i].val); }
	g_nopt = 0;
	return 0;
}

PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help)
{
	(void)help;
	if (file) opt_insert_file(file);
	if (argc && argv && *argc > 1) opt_insert_tokens(*argv + 1, *argc - 1);
	return 0;
}
PetscErrorCode PetscFinalize(void) { PetscOptionsClear(NULL); return 0; }

static const char *opt_find(const char *pre, const char *name, int *found)
{
	char key[512];
	snprintf(key, sizeof key, "%s%s", pre ? pre : "", name[0] == '-' ? name + 1 : name);
	for (int i = 0; i < g_nopt; i++)
		if (strcmp(g_opt[i].key, key) == 0) { *found = 1; return g_opt[i].val; }
	*found = 0;
	return NULL;
}

PetscErrorCode PetscOptionsHasName(void *o, const char *pre, const char *name, PetscBool *set)
{ (void)o; int f; opt_find(pre, name, &f); *set = f ? PETSC_TRUE : PETSC_FALSE; return 0; }

PetscErrorCode PetscOptionsGetInt(void *o, const char *pre, const char *name, PetscInt *iv, PetscBool *set)
{
	(void)o; int f; const char *v = opt_find(pre, name, &f);
	if (f && v) *iv = (PetscInt)strtol(v, NULL, 10);
	if (set) *set = (f && v) ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
PetscErrorCode PetscOptionsGetReal(void *o, const char *pre, const char *name, PetscReal *dv, PetscBool *set)
{
	(void)o; int f; const char *v = opt_find(pre, name, &f);
	if (f && v) *dv = strtod(v, NULL);
	if (set) *set = (f && v) ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
PetscErrorCode PetscOptionsGetString(void *o, const char *pre, const char *name, char *str, size_t len, PetscBool *set)
{
	(void)o; int f; const char *v = opt_find(pre, name, &f);
	if (f && v) { strncpy(str, v, len - 1); str[len - 1] = '\0'; }
	if (set) *set = (f && v) ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
PetscErrorCode PetscOptionsGetIntArray(void *o, const char *pre, const char *name, PetscInt *iv, PetscInt *nmax, PetscBool *set)
{
	(void)o; int f; const char *v = opt_find(pre, name, &f);
	if (!(f && v)) { *nmax = 0; if (set) *set = PETSC_FALSE; return 0; }
	char *copy = strdup(v); int n = 0;
	for (char *p = strtok(copy, ","); p && n < *nmax; p = strtok(NULL, ",")) iv[n++] = (PetscInt)strtol(p, NULL, 10);
	free(copy);
	*nmax = n; if (set) *set = PETSC_TRUE;
	return 0;
}
static int opt_bool(const char *pre, const char *name)
{
	int f; const char *v = opt_find(pre, name, &f);
	if (!f) return 0;
	if (!v) return 1;
	return !(strcmp(v, "0") == 0 || strcmp(v, "false") == 0 || strcmp(v, "no") == 0);
}

PetscErrorCode PetscPrintf(MPI_Comm c, const char *fmt, ...)
{ (void)c; if (g_quiet) return 0; va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); return 0; }
PetscErrorCode PetscSynchronizedPrintf(MPI_Comm c, const char *fmt, ...)
{ (void)c; if (g_quiet) return 0; va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); return 0; }
PetscErrorCode PetscSynchronizedFlush(MPI_Comm c, FILE *fd) { (void)c; if (fd) fflush(fd); return 0; }
PetscErrorCode PetscLogStageRegister(const char *n, PetscLogStage *s) { (void)n; *s = 0; return 0; }
PetscErrorCode PetscLogStagePush(PetscLogStage s) { (void)s; return 0; }
PetscErrorCode PetscLogStagePop(void) { return 0; }

/* =========================================================================
 * Vec
 * ========================================================================= */
struct _p_Vec {
	PetscInt n;
	PetscScalar *a;
	PetscInt ranges[2];
	/* sub-vector bookkeeping (VecGetSubVector returns a copy that is scattered back on restore) */
	int is_sub;
};

PetscErrorCode VecCreateSeq(MPI_Comm c, PetscInt n, Vec *v)
{
	(void)c;
	Vec x = calloc(1, sizeof *x);
	x->n = n; x->a = calloc((size_t)(n > 0 ? n : 1), sizeof(PetscScalar));
	if (!x->a) MP_ERR("VecCreateSeq: out of memory (n=%d)", n);
	x->ranges[0] = 0; x->ranges[1] = n;
	*v = x; return 0;
}
PetscErrorCode VecDuplicate(Vec v, Vec *nv) { return VecCreateSeq(0, v->n, nv); }
PetscErrorCode VecDestroy(Vec *v) { if (v && *v) { free((*v)->a); free(*v); *v = NULL; } return 0; }
PetscErrorCode VecGetSize(Vec x, PetscInt *n) { *n = x->n; return 0; }
PetscErrorCode VecGetArray(Vec x, PetscScalar **a) { *a = x->a; return 0; }
PetscErrorCode VecRestoreArray(Vec x, PetscScalar **a) { (void)x; if (a) *a = NULL; return 0; }
PetscErrorCode VecGetOwnershipRange(Vec x, PetscInt *lo, PetscInt *hi) { if (lo) *lo = 0; if (hi) *hi = x->n; return 0; }
PetscErrorCode VecGetOwnershipRanges(Vec x, const PetscInt *r[]) { *r = x->ranges; return 0; }
PetscErrorCode VecAssemblyBegin(Vec x) { (void)x; return 0; }
PetscErrorCode VecAssemblyEnd(Vec x) { (void)x; return 0; }
PetscErrorCode VecSetValue(Vec x, PetscInt row, PetscScalar v, InsertMode m)
{
	if (row < 0 || row >= x->n) MP_ERR("VecSetValue: row %d out of range [0,%d)", row, x->n);
	if (m == ADD_VALUES) x->a[row] += v; else x->a[row] = v;
	return 0;
}
PetscErrorCode VecSet(Vec x, PetscScalar alpha)
{
	const PetscInt n_par = x->n; PetscScalar *a = x->a;
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) a[i] = alpha;
	return 0;
}
PetscErrorCode VecCopy(Vec x, Vec y)
{
	if (x->n != y->n) MP_ERR("VecCopy: size mismatch");
	if (x != y) memcpy(y->a, x->a, (size_t)x->n * sizeof(PetscScalar));
	return 0;
}

#define RED_BLK 4096
/* blocked, deterministic dot product: sum_b ( sum_{i in b} x_i*y_i ) */
static PetscScalar blocked_dot(PetscInt n, const PetscScalar *x, const PetscScalar *y)
{
	const PetscInt nb = (n + RED_BLK - 1) / RED_BLK;
	if (nb <= 1) {
		PetscScalar s = 0.0;
		for (PetscInt i = 0; i < n; i++) s += x[i] * y[i];
		return s;
	}
	PetscScalar *part = malloc((size_t)nb * sizeof(PetscScalar));
	const PetscInt n_par = n;
	MP_PARFOR
	for (PetscInt b = 0; b < nb; b++) {
		const PetscInt lo = b * RED_BLK, hi = (lo + RED_BLK < n) ? lo + RED_BLK : n;
		PetscScalar s = 0.0;
		for (PetscInt i = lo; i < hi; i++) s += x[i] * y[i];
		part[b] = s;
	}
	(void)n_par;
	PetscScalar tot = 0.0;
	for (PetscInt b = 0; b < nb; b++) tot += part[b];
	free(part);
	return tot;
}
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val) { *val = blocked_dot(x->n, x->a, y->a); return 0; }
PetscErrorCode VecTDot(Vec x, Vec y, PetscScalar *val) { return VecDot(x, y, val); }
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val)
{
	if (type == NORM_2 || type == NORM_FROBENIUS) { *val = sqrt(blocked_dot(x->n, x->a, x->a)); return 0; }
	PetscReal s = 0.0;
	for (PetscInt i = 0; i < x->n; i++) {
		const PetscReal t = fabs(x->a[i]);
		if (type == NORM_1) s += t; else if (t > s) s = t;
	}
	*val = s; return 0;
}
/* y = y + alpha*x  (one multiply, one add, separately rounded) */
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x)
{
	const PetscInt n_par = y->n; PetscScalar *ya = y->a; const PetscScalar *xa = x->a;
	if (x->n != y->n) MP_ERR("VecAXPY: size mismatch");
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) ya[i] = ya[i] + alpha * xa[i];
	return 0;
}
/* y = x + beta*y */
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x)
{
	const PetscInt n_par = y->n; PetscScalar *ya = y->a; const PetscScalar *xa = x->a;
	if (x->n != y->n) MP_ERR("VecAYPX: size mismatch");
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) ya[i] = xa[i] + beta * ya[i];
	return 0;
}
PetscErrorCode VecWAXPY(Vec w, PetscScalar alpha, Vec x, Vec y)
{
	const PetscInt n_par = w->n; PetscScalar *wa = w->a; const PetscScalar *xa = x->a, *ya = y->a;
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) wa[i] = alpha * xa[i] + ya[i];
	return 0;
}
/* z = alpha x + beta y + gamma z */
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar alpha, PetscScalar beta, PetscScalar gamma, Vec x, Vec y)
{
	const PetscInt n_par = z->n; PetscScalar *za = z->a; const PetscScalar *xa = x->a, *ya = y->a;
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) za[i] = gamma * za[i] + alpha * xa[i] + beta * ya[i];
	return 0;
}
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y)
{
	const PetscInt n_par = w->n; PetscScalar *wa = w->a; const PetscScalar *xa = x->a, *ya = y->a;
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) wa[i] = xa[i] * ya[i];
	return 0;
}
PetscErrorCode VecScale(Vec x, PetscScalar alpha)
{
	const PetscInt n_par = x->n; PetscScalar *a = x->a;
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) a[i] = alpha * a[i];
	return 0;
}
PetscErrorCode VecView(Vec x, PetscViewer v)
{
	(void)v; if (g_quiet) return 0;
	for (PetscInt i = 0; i < x->n; i++) printf("%g\n", x->a[i]);
	return 0;
}

/* =========================================================================
 * IS + sub-vectors (only the research cycles use them; gather/scatter copies)
 * ========================================================================= */
struct _p_IS { PetscInt n; PetscInt *idx; };
PetscErrorCode ISCreateGeneral(MPI_Comm c, PetscInt n, const PetscInt idx[], PetscCopyMode mode, IS *is)
{
	(void)c; (void)mode;
	IS s = calloc(1, sizeof *s);
	s->n = n; s->idx = malloc((size_t)(n > 0 ? n : 1) * sizeof(PetscInt));
	memcpy(s->idx, idx, (size_t)n * sizeof(PetscInt));
	*is = s; return 0;
}
PetscErrorCode ISDestroy(IS *is) { if (is && *is) { free((*is)->idx); free(*is); *is = NULL; } return 0; }
PetscErrorCode ISView(IS is, PetscViewer v)
{ (void)v; if (g_quiet) return 0; for (PetscInt i = 0; i < is->n; i++) printf("%d %d\n", i, is->idx[i]); return 0; }
PetscErrorCode VecGetSubVector(Vec x, IS is, Vec *y)
{
	VecCreateSeq(0, is->n, y);
	for (PetscInt i = 0; i < is->n; i++) (*y)->a[i] = x->a[is->idx[i]];
	(*y)->is_sub = 1;
	return 0;
}
PetscErrorCode VecRestoreSubVector(Vec x, IS is, Vec *y)
{
	if (!y || !*y) return 0;
	for (PetscInt i = 0; i < is->n; i++) x->a[is->idx[i]] = (*y)->a[i];
	return VecDestroy(y);
}

/* =========================================================================
 * Mat: SeqAIJ.  Build phase = fixed-stride slab (stride = preallocated nz per
 * row, regrown if a row overflows); MatAssemblyEnd compresses to CSR with
 * ascending column indices per row, as PETSc's AIJ stores them.
 * ========================================================================= */
struct _p_Mat {
	PetscInt m, n;
	/* build phase */
	PetscInt stride;
	PetscInt *bcol; PetscScalar *bval; PetscInt *blen;
	/* assembled CSR */
	int assembled;
	PetscInt *ia, *ja; PetscScalar *va;
	PetscInt *diag;               /* position of the diagonal entry in each row (or -1) */
	/* SOR cache (MatInvertDiagonal_SeqAIJ) */
	PetscScalar *idiag, *mdiag, *ssor_work;
	PetscReal sor_omega, sor_fshift; int idiagvalid;
};

PetscErrorCode MatCreateSeqAIJ(MPI_Comm c, PetscInt m, PetscInt n, PetscInt nz, const PetscInt nnz[], Mat *A)
{
	(void)c;
	if (nnz) { nz = 0; for (PetscInt i = 0; i < m; i++) if (nnz[i] > nz) nz = nnz[i]; }
	if (nz == PETSC_DEFAULT || nz < 1) nz = 5;
	Mat a = calloc(1, sizeof *a);
	a->m = m; a->n = n; a->stride = nz;
	a->bcol = malloc((size_t)m * (size_t)nz * sizeof(PetscInt) + 8);
	a->bval = malloc((size_t)m * (size_t)nz * sizeof(PetscScalar) + 8);
	a->blen = calloc((size_t)(m > 0 ? m : 1), sizeof(PetscInt));
	if (!a->bcol || !a->bval || !a->blen) MP_ERR("MatCreateSeqAIJ: out of memory (m=%d nz=%d)", m, nz);
	*A = a; return 0;
}
PetscErrorCode MatCreateAIJ(MPI_Comm c, PetscInt m, PetscInt n, PetscInt M, PetscInt N,
                            PetscInt d_nz, const PetscInt d_nnz[], PetscInt o_nz, const PetscInt o_nnz[], Mat *A)
{
	(void)o_nz; (void)o_nnz;
	if (m < 0) m = M;
	if (n < 0) n = N;
	return MatCreateSeqAIJ(c, m, n, d_nz, d_nnz, A);
}

static void mat_regrow(Mat a)
{
	const PetscInt ns = 2 * a->stride;
	PetscInt *nc = malloc((size_t)a->m * (size_t)ns * sizeof(PetscInt) + 8);
	PetscScalar *nv = malloc((size_t)a->m * (size_t)ns * sizeof(PetscScalar) + 8);
	if (!nc || !nv) MP_ERR("MatSetValue: out of memory while growing rows");
	for (PetscInt i = 0; i < a->m; i++) {
		memcpy(nc + (size_t)i * ns, a->bcol + (size_t)i * a->stride, (size_t)a->blen[i] * sizeof(PetscInt));
		memcpy(nv + (size_t)i * ns, a->bval + (size_t)i * a->stride, (size_t)a->blen[i] * sizeof(PetscScalar));
	}
	free(a->bcol); free(a->bval);
	a->bcol = nc; a->bval = nv; a->stride = ns;
}

PetscErrorCode MatSetValue(Mat a, PetscInt row, PetscInt col, PetscScalar v, InsertMode mode)
{
	if (a->assembled) MP_ERR("MatSetValue after MatAssemblyEnd is not restated");
	if (row < 0 || row >= a->m || col < 0 || col >= a->n) MP_ERR("MatSetValue: (%d,%d) outside %dx%d", row, col, a->m, a->n);
	PetscInt *c = a->bcol + (size_t)row * a->stride;
	PetscScalar *w = a->bval + (size_t)row * a->stride;
	PetscInt len = a->blen[row], k = len;
	while (k > 0 && c[k - 1] >= col) k--;          /* first position with c[k] >= col */
	if (k < len && c[k] == col) { if (mode == ADD_VALUES) w[k] += v; else w[k] = v; return 0; }
	if (len == a->stride) { mat_regrow(a); c = a->bcol + (size_t)row * a->stride; w = a->bval + (size_t)row * a->stride; }
	for (PetscInt t = len; t > k; t--) { c[t] = c[t - 1]; w[t] = w[t - 1]; }
	c[k] = col; w[k] = v; a->blen[row] = len + 1;
	return 0;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat a, MatAssemblyType t)
{
	if (t != MAT_FINAL_ASSEMBLY || a->assembled) return 0;
	a->ia = malloc(((size_t)a->m + 1) * sizeof(PetscInt));
	long long nnz = 0;
	for (PetscInt i = 0; i < a->m; i++) { a->ia[i] = (PetscInt)nnz; nnz += a->blen[i]; }
	if (nnz > 2147483647LL) MP_ERR("MatAssemblyEnd: nnz overflows 32-bit PetscInt");
	a->ia[a->m] = (PetscInt)nnz;
	a->ja = malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(PetscInt));
	a->va = malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(PetscScalar));
	a->diag = malloc((size_t)(a->m > 0 ? a->m : 1) * sizeof(PetscInt));
	if (!a->ja || !a->va || !a->diag) MP_ERR("MatAssemblyEnd: out of memory");
	for (PetscInt i = 0; i < a->m; i++) {
		const PetscInt len = a->blen[i];
		memcpy(a->ja + a->ia[i], a->bcol + (size_t)i * a->stride, (size_t)len * sizeof(PetscInt));
		memcpy(a->va + a->ia[i], a->bval + (size_t)i * a->stride, (size_t)len * sizeof(PetscScalar));
		a->diag[i] = -1;
		for (PetscInt k = 0; k < len; k++) if (a->ja[a->ia[i] + k] == i) a->diag[i] = a->ia[i] + k;
	}
	free(a->bcol); free(a->bval); free(a->blen);
	a->bcol = NULL; a->bval = NULL; a->blen = NULL;
	a->assembled = 1;
	return 0;
}
PetscErrorCode MatDestroy(Mat *A)
{
	if (!A || !*A) return 0;
	Mat a = *A;
	free(a->bcol); free(a->bval); free(a->blen);
	free(a->ia); free(a->ja); free(a->va); free(a->diag);
	free(a->idiag); free(a->mdiag); free(a->ssor_work);
	free(a); *A = NULL; return 0;
}
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n) { if (m) *m = A->m; if (n) *n = A->n; return 0; }
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left)
{
	if (right) VecCreateSeq(0, A->n, right);
	if (left) VecCreateSeq(0, A->m, left);
	return 0;
}
PetscErrorCode MatSeqAIJGetCSR(Mat A, PetscInt *m, PetscInt *n, const PetscInt **ia, const PetscInt **ja, const PetscScalar **va)
{
	if (!A->assembled) MP_ERR("MatSeqAIJGetCSR: matrix not assembled");
	if (m) *m = A->m;
	if (n) *n = A->n;
	if (ia) *ia = A->ia;
	if (ja) *ja = A->ja;
	if (va) *va = A->va;
	return 0;
}

/* y_i = sum_k a_ik x_k, accumulated from 0.0 in ascending column order (MatMult_SeqAIJ) */
PetscErrorCode MatMult(Mat A, Vec x, Vec y)
{
	if (!A->assembled) MP_ERR("MatMult: matrix not assembled");
	if (x->n != A->n || y->n != A->m) MP_ERR("MatMult: size mismatch (A %dx%d, x %d, y %d)", A->m, A->n, x->n, y->n);
	if (x == y) MP_ERR("MatMult: x and y must differ");
	const PetscInt n_par = A->m; const PetscInt *ia = A->ia, *ja = A->ja; const PetscScalar *va = A->va, *xa = x->a; PetscScalar *ya = y->a;
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) {
		PetscScalar sum = 0.0;
		for (PetscInt k = ia[i]; k < ia[i + 1]; k++) sum += va[k] * xa[ja[k]];
		ya[i] = sum;
	}
	return 0;
}
/* z_i = y_i + sum_k a_ik x_k, accumulation STARTS from y_i (MatMultAdd_SeqAIJ) */
PetscErrorCode MatMultAdd(Mat A, Vec x, Vec y, Vec z)
{
	if (!A->assembled) MP_ERR("MatMultAdd: matrix not assembled");
	if (x->n != A->n || y->n != A->m || z->n != A->m) MP_ERR("MatMultAdd: size mismatch");
	const PetscInt n_par = A->m; const PetscInt *ia = A->ia, *ja = A->ja; const PetscScalar *va = A->va, *xa = x->a, *ya = y->a; PetscScalar *za = z->a;
	MP_PARFOR
	for (PetscInt i = 0; i < n_par; i++) {
		PetscScalar sum = ya[i];
		for (PetscInt k = ia[i]; k < ia[i + 1]; k++) sum += va[k] * xa[ja[k]];
		za[i] = sum;
	}
	return 0;
}
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y)
{
	if (x->n != A->m || y->n != A->n) MP_ERR("MatMultTranspose: size mismatch");
	VecSet(y, 0.0);
	for (PetscInt i = 0; i < A->m; i++)
		for (PetscInt k = A->ia[i]; k < A->ia[i + 1]; k++) y->a[A->ja[k]] += A->va[k] * x->a[i];
	return 0;
}
/* r = b - A x : MatMult then VecAYPX(r,-1,b) */
PetscErrorCode MatResidual(Mat A, Vec b, Vec x, Vec r) { MatMult(A, x, r); return VecAYPX(r, -1.0, b); }
PetscErrorCode MatRestrict(Mat A, Vec x, Vec y)
{
	if (A->n == x->n) return MatMult(A, x, y);
	return MatMultTranspose(A, x, y);
}
PetscErrorCode MatInterpolateAdd(Mat A, Vec x, Vec y, Vec w)
{
	if (A->n == x->n) return MatMultAdd(A, x, y, w);
	MP_ERR("MatInterpolateAdd with a transposed interpolation is not restated");
	return 1;
}
PetscErrorCode MatScale(Mat A, PetscScalar s)
{
	for (PetscInt k = 0; k < A->ia[A->m]; k++) A->va[k] *= s;
	A->idiagvalid = 0; return 0;
}
PetscErrorCode MatGetDiagonal(Mat A, Vec d)
{
	for (PetscInt i = 0; i < A->m; i++) d->a[i] = A->diag[i] >= 0 ? A->va[A->diag[i]] : 0.0;
	return 0;
}
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat *C)
{
	(void)scall; (void)fill;
	if (A->n != B->m) MP_ERR("MatMatMult: size mismatch");
	Mat c; MatCreateSeqAIJ(0, A->m, B->n, 16, NULL, &c);
	for (PetscInt i = 0; i < A->m; i++)
		for (PetscInt k = A->ia[i]; k < A->ia[i + 1]; k++) {
			const PetscInt r = A->ja[k];
			for (PetscInt t = B->ia[r]; t < B->ia[r + 1]; t++) MatSetValue(c, i, B->ja[t], A->va[k] * B->va[t], ADD_VALUES);
		}
	MatAssemblyEnd(c, MAT_FINAL_ASSEMBLY);
	*C = c; return 0;
}
PetscErrorCode MatView(Mat A, PetscViewer v)
{
	if (g_quiet || v != PETSC_VIEWER_STDOUT_WORLD) return 0;
	for (PetscInt i = 0; i < A->m; i++) {
		printf("row %d:", i);
		for (PetscInt k = A->ia[i]; k < A->ia[i + 1]; k++) printf(" (%d, %g) ", A->ja[k], A->va[k]);
		printf("\n");
	}
	return 0;
}

/* MatInvertDiagonal_SeqAIJ: idiag = omega/(fshift + d), mdiag = d */
static void mat_invert_diagonal(Mat a, PetscReal omega, PetscReal fshift)
{
	if (!a->idiag) {
		a->idiag = malloc((size_t)a->m * sizeof(PetscScalar));
		a->mdiag = malloc((size_t)a->m * sizeof(PetscScalar));
		a->ssor_work = malloc((size_t)a->m * sizeof(PetscScalar));
	}
	for (PetscInt i = 0; i < a->m; i++) {
		if (a->diag[i] < 0) MP_ERR("MatSOR: missing diagonal in row %d", i);
		const PetscScalar d = a->va[a->diag[i]];
		a->mdiag[i] = d;
		if (omega == 1.0 && fshift == 0.0) {
			if (d == 0.0) MP_ERR("MatSOR: zero diagonal in row %d", i);
			a->idiag[i] = 1.0 / d;
		} else {
			a->idiag[i] = omega / (fshift + d);
		}
	}
	a->sor_omega = omega; a->sor_fshift = fshift; a->idiagvalid = 1;
}

/*
 * MatSOR_SeqAIJ.  sum -= v[k]*x[idx[k]] is applied entry by entry in ascending
 * column order (PetscSparseDenseMinusDot).  Local and global sweeps coincide on
 * one rank.  The "t" array checkpoints b - L x from the forward sweep so that
 * the backward sweep only applies the strictly upper part.
 */
PetscErrorCode MatSOR(Mat a, Vec bb, PetscReal omega, MatSORType flag, PetscReal fshift, PetscInt its, PetscInt lits, Vec xx)
{
	if (!a->assembled) MP_ERR("MatSOR: matrix not assembled");
	if (flag & (SOR_EISENSTAT | SOR_APPLY_UPPER | SOR_APPLY_LOWER)) MP_ERR("MatSOR: Eisenstat/apply-upper/lower not restated");
	if (its <= 0 || lits <= 0) MP_ERR("MatSOR: its=%d lits=%d must be positive", its, lits);
	its = its * lits;
	if (!a->idiagvalid || fshift != a->sor_fshift || omega != a->sor_omega) mat_invert_diagonal(a, omega, fshift);
	const PetscInt m = a->m, *ai = a->ia, *aj = a->ja, *diag = a->diag;
	const PetscScalar *aa = a->va, *idiag = a->idiag, *mdiag = a->mdiag, *b = bb->a, *xb;
	PetscScalar *x = xx->a, *t = a->ssor_work, sum;
	const int fwd = (flag & SOR_FORWARD_SWEEP) || (flag & SOR_LOCAL_FORWARD_SWEEP);
	const int bwd = (flag & SOR_BACKWARD_SWEEP) || (flag & SOR_LOCAL_BACKWARD_SWEEP);

	if (flag & SOR_ZERO_INITIAL_GUESS) {
		if (fwd) {
			for (PetscInt i = 0; i < m; i++) {
				sum = b[i];
				for (PetscInt k = ai[i]; k < diag[i]; k++) sum -= aa[k] * x[aj[k]];
				t[i] = sum;
				x[i] = sum * idiag[i];
			}
			xb = t;
		} else xb = b;
		if (bwd) {
			for (PetscInt i = m - 1; i >= 0; i--) {
				sum = xb[i];
				for (PetscInt k = diag[i] + 1; k < ai[i + 1]; k++) sum -= aa[k] * x[aj[k]];
				if (xb == b) x[i] = sum * idiag[i];
				else         x[i] = (1 - omega) * x[i] + sum * idiag[i];
			}
		}
		its--;
	}
	while (its--) {
		if (fwd) {
			for (PetscInt i = 0; i < m; i++) {
				sum = b[i];
				for (PetscInt k = ai[i]; k < diag[i]; k++) sum -= aa[k] * x[aj[k]];
				t[i] = sum;
				for (PetscInt k = diag[i] + 1; k < ai[i + 1]; k++) sum -= aa[k] * x[aj[k]];
				x[i] = (1. - omega) * x[i] + sum * idiag[i];
			}
			xb = t;
		} else xb = b;
		if (bwd) {
			for (PetscInt i = m - 1; i >= 0; i--) {
				sum = xb[i];
				if (xb == b) {
					for (PetscInt k = ai[i]; k < ai[i + 1]; k++) sum -= aa[k] * x[aj[k]];
					x[i] = (1. - omega) * x[i] + (sum + mdiag[i] * x[i]) * idiag[i];
				} else {
					for (PetscInt k = diag[i] + 1; k < ai[i + 1]; k++) sum -= aa[k] * x[aj[k]];
					x[i] = (1. - omega) * x[i] + sum * idiag[i];
				}
			}
		}
	}
	return 0;
}

/* =========================================================================
 * PC
 * ========================================================================= */
typedef struct {
	KSP  smooth;          /* smoothd == smoothu (PETSc default) ; level 0: coarse solver */
	Mat  restrct, interpolate;
	Vec  b, x, r;
	int  own_b, own_x, own_r;
} MGLevel;

struct _p_PC {
	char type[32]; int type_set;
	char prefix[128];
	Mat  A; int setup;
	/* jacobi */
	Vec  dinv;
	/* sor */
	PetscReal omega, fshift; PetscInt its, lits; MatSORType sym;
	/* ilu(0) / dense lu */
	PetscScalar *fac; PetscInt *fdiag; PetscScalar *dense; PetscScalar *tmp;
	/* mg */
	PetscInt nlevels; MGLevel *lev; PetscInt cycles;
	PetscReal mg_ttol;
};

struct _p_KSP {
	char type[32];
	char prefix[128];
	Mat A; PC pc;
	KSPNormType normtype; int normtype_set;
	PetscReal rtol, abstol, dtol; PetscInt max_it;
	PetscBool guess_nonzero;
	PetscReal scale;
	Vec vec_sol, vec_rhs;
	Vec work[4]; PetscInt worksize;
	PetscInt its; KSPConvergedReason reason; PetscReal rnorm, rnorm0, ttol;
	PetscReal *res_hist; PetscInt res_hist_len, res_hist_max;
	PetscErrorCode (*monitor[4])(KSP, PetscInt, PetscReal, void *); void *mctx[4]; PetscInt nmon;
	int print_monitor;
};

static PC pc_create(void)
{
	PC pc = calloc(1, sizeof *pc);
	pc->omega = 1.0; pc->fshift = 0.0; pc->its = 1; pc->lits = 1; pc->sym = SOR_LOCAL_SYMMETRIC_SWEEP;
	pc->cycles = 1;
	return pc;
}
PetscErrorCode PCSetType(PC pc, PCType type)
{
	if (pc->type_set && strcmp(pc->type, type) == 0) return 0;
	strncpy(pc->type, type, sizeof pc->type - 1); pc->type_set = 1; pc->setup = 0;
	return 0;
}

static void pc_free_data(PC pc)
{
	VecDestroy(&pc->dinv);
	free(pc->fac); free(pc->fdiag); free(pc->dense); free(pc->tmp);
	pc->fac = NULL; pc->fdiag = NULL; pc->dense = NULL; pc->tmp = NULL;
}
static void pc_destroy(PC *ppc)
{
	if (!ppc || !*ppc) return;
	PC pc = *ppc;
	pc_free_data(pc);
	if (pc->lev) {
		for (PetscInt l = 0; l < pc->nlevels; l++) {
			KSPDestroy(&pc->lev[l].smooth);
			if (pc->lev[l].own_b) VecDestroy(&pc->lev[l].b);
			if (pc->lev[l].own_x) VecDestroy(&pc->lev[l].x);
			if (pc->lev[l].own_r) VecDestroy(&pc->lev[l].r);
		}
		free(pc->lev);
	}
	free(pc); *ppc = NULL;
}

/* ILU(0) on the CSR pattern, natural ordering; stores L (unit, strictly lower),
 * U (strictly upper) in place and the INVERTED pivot on the diagonal
 * (MatLUFactorNumeric_SeqAIJ convention: multiplier = a_ik * (1/pivot_k)). */
static void pc_setup_ilu0(PC pc)
{
	Mat a = pc->A; const PetscInt m = a->m;
	const PetscInt nnz = a->ia[m];
	pc->fac = malloc((size_t)nnz * sizeof(PetscScalar));
	pc->tmp = calloc((size_t)m, sizeof(PetscScalar));
	memcpy(pc->fac, a->va, (size_t)nnz * sizeof(PetscScalar));
	PetscInt *pos = malloc((size_t)a->n * sizeof(PetscInt));
	for (PetscInt j = 0; j < a->n; j++) pos[j] = -1;
	for (PetscInt i = 0; i < m; i++) {
		if (a->diag[i] < 0) MP_ERR("ILU(0): missing diagonal in row %d", i);
		for (PetscInt k = a->ia[i]; k < a->ia[i + 1]; k++) pos[a->ja[k]] = k;
		for (PetscInt k = a->ia[i]; k < a->diag[i]; k++) {
			const PetscInt row = a->ja[k];
			if (pc->fac[k] != 0.0) {
				const PetscScalar mult = pc->fac[k] * pc->fac[a->diag[row]];
				pc->fac[k] = mult;
				for (PetscInt t = a->diag[row] + 1; t < a->ia[row + 1]; t++) {
					const PetscInt p = pos[a->ja[t]];
					if (p >= 0) pc->fac[p] -= mult * pc->fac[t];
				}
			}
		}
		if (pc->fac[a->diag[i]] == 0.0) MP_ERR("ILU(0): zero pivot in row %d", i);
		pc->fac[a->diag[i]] = 1.0 / pc->fac[a->diag[i]];
		for (PetscInt k = a->ia[i]; k < a->ia[i + 1]; k++) pos[a->ja[k]] = -1;
	}
	free(pos);
}
static void pc_apply_ilu0(PC pc, Vec bvec, Vec xvec)
{
	Mat a = pc->A; const PetscInt m = a->m; PetscScalar *tmp = pc->tmp; const PetscScalar *b = bvec->a; PetscScalar *x = xvec->a;
	for (PetscInt i = 0; i < m; i++) {
		PetscScalar sum = b[i];
		for (PetscInt k = a->ia[i]; k < a->diag[i]; k++) sum -= pc->fac[k] * tmp[a->ja[k]];
		tmp[i] = sum;
	}
	for (PetscInt i = m - 1; i >= 0; i--) {
		PetscScalar sum = tmp[i];
		for (PetscInt k = a->diag[i] + 1; k < a->ia[i + 1]; k++) sum -= pc->fac[k] * tmp[a->ja[k]];
		x[i] = tmp[i] = sum * pc->fac[a->diag[i]];
	}
}
/* dense LU, natural order, no pivoting, inverted pivots (tiny coarse problems only).
 * NOT bit-comparable with PETSc's sparse LU under its default nested-dissection ordering
 * except for the 1x1 case, where both give x = b * (1/a). */
static void pc_setup_lu(PC pc)
{
	Mat a = pc->A; const PetscInt m = a->m;
	if (m != a->n) MP_ERR("LU: matrix not square");
	if (m > 4096) MP_ERR("LU: the dense restatement only handles coarse problems up to 4096 unknowns (got %d)", m);
	pc->dense = calloc((size_t)m * (size_t)m, sizeof(PetscScalar));
	pc->tmp = calloc((size_t)m, sizeof(PetscScalar));
	PetscScalar *d = pc->dense;
	for (PetscInt i = 0; i < m; i++) for (PetscInt k = a->ia[i]; k < a->ia[i + 1]; k++) d[(size_t)i * m + a->ja[k]] = a->va[k];
	for (PetscInt i = 0; i < m; i++) {
		for (PetscInt k = 0; k < i; k++) {
			if (d[(size_t)i * m + k] != 0.0) {
				const PetscScalar mult = d[(size_t)i * m + k] * d[(size_t)k * m + k];
				d[(size_t)i * m + k] = mult;
				for (PetscInt j = k + 1; j < m; j++) d[(size_t)i * m + j] -= mult * d[(size_t)k * m + j];
			}
		}
		if (d[(size_t)i * m + i] == 0.0) MP_ERR("LU: zero pivot in row %d", i);
		d[(size_t)i * m + i] = 1.0 / d[(size_t)i * m + i];
	}
}
static void pc_apply_lu(PC pc, Vec bvec, Vec xvec)
{
	const PetscInt m = pc->A->m; const PetscScalar *d = pc->dense, *b = bvec->a; PetscScalar *tmp = pc->tmp, *x = xvec->a;
	for (PetscInt i = 0; i < m; i++) {
		PetscScalar sum = b[i];
		for (PetscInt k = 0; k < i; k++) sum -= d[(size_t)i * m + k] * tmp[k];
		tmp[i] = sum;
	}
	for (PetscInt i = m - 1; i >= 0; i--) {
		PetscScalar sum = tmp[i];
		for (PetscInt k = i + 1; k < m; k++) sum -= d[(size_t)i * m + k] * tmp[k];
		x[i] = tmp[i] = sum * d[(size_t)i * m + i];
	}
}

static void pc_setup_mg(PC pc);
static void pc_apply_mg(PC pc, Vec b, Vec x);

static void pc_setup(PC pc)
{
	if (pc->setup) return;
	if (!pc->A) MP_ERR("PCSetUp: no operator set");
	if (!pc->type_set) { strcpy(pc->type, PCILU); pc->type_set = 1; }  /* PETSc's 1-rank default for AIJ */
	pc_free_data(pc);
	if (strcmp(pc->type, PCJACOBI) == 0) {
		/* PCSetUp_Jacobi: stores the reciprocal of the diagonal; zero diagonals become 1 */
		VecCreateSeq(0, pc->A->m, &pc->dinv);
		for (PetscInt i = 0; i < pc->A->m; i++) {
			const PetscScalar d = pc->A->diag[i] >= 0 ? pc->A->va[pc->A->diag[i]] : 0.0;
			pc->dinv->a[i] = (d != 0.0) ? 1.0 / d : 1.0;
		}
	} else if (strcmp(pc->type, PCILU) == 0) pc_setup_ilu0(pc);
	else if (strcmp(pc->type, PCLU) == 0) pc_setup_lu(pc);
	else if (strcmp(pc->type, PCMG) == 0) pc_setup_mg(pc);
	else if (strcmp(pc->type, PCSOR) == 0 || strcmp(pc->type, PCNONE) == 0) { /* nothing */ }
	else MP_ERR("PC type '%s' is not restated in minipetsc (have: none jacobi sor ilu lu mg)", pc->type);
	pc->setup = 1;
}

PetscErrorCode PCApply(PC pc, Vec x, Vec y)
{
	pc_setup(pc);
	if (strcmp(pc->type, PCNONE) == 0) return VecCopy(x, y);
	if (strcmp(pc->type, PCJACOBI) == 0) return VecPointwiseMult(y, x, pc->dinv);
	if (strcmp(pc->type, PCSOR) == 0)   /* PCApply_SOR */
		return MatSOR(pc->A, x, pc->omega, (MatSORType)(pc->sym | SOR_ZERO_INITIAL_GUESS), pc->fshift, pc->its, pc->lits, y);
	if (strcmp(pc->type, PCILU) == 0) { pc_apply_ilu0(pc, x, y); return 0; }
	if (strcmp(pc->type, PCLU) == 0) { pc_apply_lu(pc, x, y); return 0; }
	if (strcmp(pc->type, PCMG) == 0) { pc_apply_mg(pc, x, y); return 0; }
	MP_ERR("PCApply: unknown type %s", pc->type);
	return 1;
}

static void pc_set_from_options(PC pc, const char *prefix)
{
	char buf[64]; PetscBool set;
	strncpy(pc->prefix, prefix, sizeof pc->prefix - 1);
	PetscOptionsGetString(NULL, prefix, "-pc_type", buf, sizeof buf, &set);
	if (set) PCSetType(pc, buf);
	PetscOptionsGetReal(NULL, prefix, "-pc_sor_omega", &pc->omega, NULL);
	PetscOptionsGetInt(NULL, prefix, "-pc_sor_its", &pc->its, NULL);
	PetscOptionsGetInt(NULL, prefix, "-pc_sor_lits", &pc->lits, NULL);
	if (opt_bool(prefix, "-pc_sor_symmetric")) pc->sym = SOR_SYMMETRIC_SWEEP;
	if (opt_bool(prefix, "-pc_sor_backward")) pc->sym = SOR_BACKWARD_SWEEP;
	if (opt_bool(prefix, "-pc_sor_forward")) pc->sym = SOR_FORWARD_SWEEP;
	if (opt_bool(prefix, "-pc_sor_local_symmetric")) pc->sym = SOR_LOCAL_SYMMETRIC_SWEEP;
	if (opt_bool(prefix, "-pc_sor_local_backward")) pc->sym = SOR_LOCAL_BACKWARD_SWEEP;
	if (opt_bool(prefix, "-pc_sor_local_forward")) pc->sym = SOR_LOCAL_FORWARD_SWEEP;
}

/* ---- PCMG (levels numbered coarse = 0 .. fine = nlevels-1, as in PETSc) ---- */
PetscErrorCode PCMGSetLevels(PC pc, PetscInt levels, MPI_Comm *comms)
{
	(void)comms;
	if (pc->lev) MP_ERR("PCMGSetLevels called twice");
	pc->nlevels = levels;
	pc->lev = calloc((size_t)levels, sizeof(MGLevel));
	for (PetscInt l = 0; l < levels; l++) {
		KSP k; KSPCreate(0, &k);
		pc->lev[l].smooth = k;
		KSPSetNormType(k, KSP_NORM_NONE);
		if (l == 0) {
			/* coarse solver: preonly + LU, prefix mg_coarse_ */
			KSPSetType(k, KSPPREONLY); PCSetType(k->pc, PCLU);
			KSPSetTolerances(k, PETSC_DEFAULT, PETSC_DEFAULT, PETSC_DEFAULT, 1);
		} else {
			/* level smoother: PETSc default is Chebyshev(2)+SOR, prefix mg_levels_ */
			KSPSetType(k, "chebyshev"); PCSetType(k->pc, PCSOR);
			KSPSetTolerances(k, PETSC_DEFAULT, PETSC_DEFAULT, PETSC_DEFAULT, 2);
		}
	}
	return 0;
}
PetscErrorCode PCMGGetCoarseSolve(PC pc, KSP *ksp) { *ksp = pc->lev[0].smooth; return 0; }
PetscErrorCode PCMGGetSmoother(PC pc, PetscInt l, KSP *ksp) { *ksp = pc->lev[l].smooth; return 0; }
PetscErrorCode PCMGSetInterpolation(PC pc, PetscInt l, Mat m) { if (l <= 0) MP_ERR("PCMGSetInterpolation on level 0"); pc->lev[l].interpolate = m; return 0; }
PetscErrorCode PCMGSetRestriction(PC pc, PetscInt l, Mat m) { if (l <= 0) MP_ERR("PCMGSetRestriction on level 0"); pc->lev[l].restrct = m; return 0; }
PetscErrorCode PCMGSetR(PC pc, PetscInt l, Vec c) { pc->lev[l].r = c; return 0; }
PetscErrorCode PCMGSetRhs(PC pc, PetscInt l, Vec c) { pc->lev[l].b = c; return 0; }
PetscErrorCode PCMGSetX(PC pc, PetscInt l, Vec c) { pc->lev[l].x = c; return 0; }
PetscErrorCode PCMGSetNumberSmoothUp(PC pc, PetscInt n)
{ for (PetscInt l = 1; l < pc->nlevels; l++) pc->lev[l].smooth->max_it = n; return 0; }
PetscErrorCode PCMGSetNumberSmoothDown(PC pc, PetscInt n) { return PCMGSetNumberSmoothUp(pc, n); }
PetscErrorCode PCASMSetType(PC pc, PCASMType t) { (void)pc; (void)t; MP_ERR("PCASM is not restated"); return 1; }
PetscErrorCode PCASMSetOverlap(PC pc, PetscInt o) { (void)pc; (void)o; MP_ERR("PCASM is not restated"); return 1; }
PetscErrorCode PCASMSetTotalSubdomains(PC pc, PetscInt N, IS a[], IS b[]) { (void)pc; (void)N; (void)a; (void)b; MP_ERR("PCASM is not restated"); return 1; }

static void ksp_set_from_options_prefixed(KSP ksp, const char *prefix);

static void pc_setup_mg(PC pc)
{
	const PetscInt n = pc->nlevels;
	if (!pc->lev) MP_ERR("PCMG: PCMGSetLevels was not called");
	for (PetscInt l = 0; l < n; l++) {
		MGLevel *L = &pc->lev[l];
		char pre[160];
		if (l == 0) snprintf(pre, sizeof pre, "%smg_coarse_", pc->prefix);
		else        snprintf(pre, sizeof pre, "%smg_levels_", pc->prefix);
		ksp_set_from_options_prefixed(L->smooth, pre);
		if (l > 0) {
			snprintf(pre, sizeof pre, "%smg_levels_%d_", pc->prefix, l);
			ksp_set_from_options_prefixed(L->smooth, pre);
		}
		if (!L->smooth->A) {
			if (l == n - 1) KSPSetOperators(L->smooth, pc->A, pc->A);
			else MP_ERR("PCMG: no operator on level %d (Galerkin coarsening is not restated)", l);
		}
		if (l > 0 && (!L->restrct || !L->interpolate)) MP_ERR("PCMG: missing restriction/interpolation on level %d", l);
		/* work vectors the user did not supply */
		const PetscInt m = L->smooth->A->m;
		if (l < n - 1) {
			if (!L->b) { VecCreateSeq(0, m, &L->b); L->own_b = 1; }
			if (!L->x) { VecCreateSeq(0, m, &L->x); L->own_x = 1; }
		}
		if (l > 0 && !L->r) { VecCreateSeq(0, m, &L->r); L->own_r = 1; }
		/* PCSetUp_MG: the (shared) level smoother runs with a nonzero initial guess */
		if (l > 0) KSPSetInitialGuessNonzero(L->smooth, PETSC_TRUE);
	}
}

/* PCMGMCycle_Private (V: cycles = 1) */
static void pc_mg_cycle(PC pc, PetscInt l, int *converged)
{
	MGLevel *L = &pc->lev[l];
	KSPSolve(L->smooth, L->b, L->x);                                 /* pre-smooth (coarse: the solve) */
	if (l > 0) {
		MatResidual(L->smooth->A, L->b, L->x, L->r);                 /* r = b - A x */
		if (l == pc->nlevels - 1 && pc->mg_ttol > 0.0 && converged) {
			PetscReal rn; VecNorm(L->r, NORM_2, &rn);
			if (rn <= pc->mg_ttol) { *converged = 1; return; }
		}
		MGLevel *C = &pc->lev[l - 1];
		MatRestrict(L->restrct, L->r, C->b);
		VecSet(C->x, 0.0);
		PetscInt cyc = (l == 1) ? 1 : pc->cycles;
		while (cyc--) pc_mg_cycle(pc, l - 1, converged);
		MatInterpolateAdd(L->interpolate, C->x, L->x, L->x);
		KSPSolve(L->smooth, L->b, L->x);                             /* post-smooth */
	}
}
/* PCApply_MG, multiplicative: x = 0, one cycle */
static void pc_apply_mg(PC pc, Vec b, Vec x)
{
	MGLevel *F = &pc->lev[pc->nlevels - 1];
	F->b = b; F->x = x;
	VecSet(x, 0.0);
	pc->mg_ttol = 0.0;
	pc_mg_cycle(pc, pc->nlevels - 1, NULL);
}

/* =========================================================================
 * KSP
 * ========================================================================= */
PetscErrorCode KSPCreate(MPI_Comm c, KSP *out)
{
	(void)c;
	KSP k = calloc(1, sizeof *k);
	strcpy(k->type, KSPGMRES);                /* PETSc's default type; the reference always overrides it */
	k->pc = pc_create();
	k->normtype = KSP_NORM_PRECONDITIONED;
	k->rtol = 1e-5; k->abstol = 1e-50; k->dtol = 1e4; k->max_it = 10000;
	k->scale = 1.0;
	*out = k; return 0;
}
PetscErrorCode KSPDestroy(KSP *pk)
{
	if (!pk || !*pk) return 0;
	KSP k = *pk;
	for (int i = 0; i < 4; i++) VecDestroy(&k->work[i]);
	pc_destroy(&k->pc);
	free(k); *pk = NULL; return 0;
}
PetscErrorCode PetscObjectSetOptionsPrefix(PetscObject obj, const char *prefix)
{
	/* only ever called on KSP objects by the reference (src/solver.c:1624-1643) */
	KSP k = (KSP)obj;
	strncpy(k->prefix, prefix ? prefix : "", sizeof k->prefix - 1);
	return 0;
}
PetscErrorCode KSPSetType(KSP k, KSPType t) { strncpy(k->type, t, sizeof k->type - 1); return 0; }
PetscErrorCode KSPSetOperators(KSP k, Mat A, Mat P)
{
	(void)P; k->A = A;
	if (k->pc->A != A) { k->pc->A = A; k->pc->setup = 0; }
	return 0;
}
PetscErrorCode KSPSetNormType(KSP k, KSPNormType t) { k->normtype = t; k->normtype_set = 1; return 0; }
PetscErrorCode KSPSetTolerances(KSP k, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits)
{
	if (rtol != (PetscReal)PETSC_DEFAULT) k->rtol = rtol;
	if (abstol != (PetscReal)PETSC_DEFAULT) k->abstol = abstol;
	if (dtol != (PetscReal)PETSC_DEFAULT) k->dtol = dtol;
	if (maxits != PETSC_DEFAULT) k->max_it = maxits;
	return 0;
}
PetscErrorCode KSPSetInitialGuessNonzero(KSP k, PetscBool f) { k->guess_nonzero = f; return 0; }
PetscErrorCode KSPRichardsonSetScale(KSP k, PetscReal s) { k->scale = s; return 0; }
PetscErrorCode KSPGetPC(KSP k, PC *pc) { *pc = k->pc; return 0; }
PetscErrorCode KSPGetIterationNumber(KSP k, PetscInt *its) { *its = k->its; return 0; }
PetscErrorCode KSPGetConvergedReason(KSP k, KSPConvergedReason *r) { *r = k->reason; return 0; }
PetscErrorCode KSPSetResidualHistory(KSP k, PetscReal a[], PetscInt na, PetscBool reset)
{ (void)reset; k->res_hist = a; k->res_hist_max = na; k->res_hist_len = 0; return 0; }
PetscErrorCode KSPGetResidualHistory(KSP k, PetscReal *a[], PetscInt *na)
{ if (a) *a = k->res_hist; if (na) *na = k->res_hist_len; return 0; }
PetscErrorCode KSPMonitorSet(KSP k, PetscErrorCode (*mon)(KSP, PetscInt, PetscReal, void *), void *ctx, PetscErrorCode (*d)(void **))
{
	(void)d;
	if (k->nmon >= 4) MP_ERR("KSPMonitorSet: too many monitors");
	k->monitor[k->nmon] = mon; k->mctx[k->nmon] = ctx; k->nmon++;
	return 0;
}

static void ksp_set_from_options_prefixed(KSP k, const char *prefix)
{
	char buf[64]; PetscBool set;
	PetscOptionsGetString(NULL, prefix, "-ksp_type", buf, sizeof buf, &set);
	if (set) KSPSetType(k, buf);
	PetscOptionsGetInt(NULL, prefix, "-ksp_max_it", &k->max_it, NULL);
	PetscOptionsGetReal(NULL, prefix, "-ksp_rtol", &k->rtol, NULL);
	PetscOptionsGetReal(NULL, prefix, "-ksp_atol", &k->abstol, NULL);
	PetscOptionsGetReal(NULL, prefix, "-ksp_divtol", &k->dtol, NULL);
	PetscOptionsGetReal(NULL, prefix, "-ksp_richardson_scale", &k->scale, NULL);
	PetscOptionsGetString(NULL, prefix, "-ksp_norm_type", buf, sizeof buf, &set);
	if (set) {
		if (!strcmp(buf, "none")) KSPSetNormType(k, KSP_NORM_NONE);
		else if (!strcmp(buf, "preconditioned")) KSPSetNormType(k, KSP_NORM_PRECONDITIONED);
		else if (!strcmp(buf, "unpreconditioned")) KSPSetNormType(k, KSP_NORM_UNPRECONDITIONED);
		else if (!strcmp(buf, "natural")) KSPSetNormType(k, KSP_NORM_NATURAL);
		else MP_ERR("unknown -ksp_norm_type %s", buf);
	}
	int f; opt_find(prefix, "-ksp_initial_guess_nonzero", &f);
	if (f) k->guess_nonzero = opt_bool(prefix, "-ksp_initial_guess_nonzero") ? PETSC_TRUE : PETSC_FALSE;
	if (opt_bool(prefix, "-ksp_monitor")) k->print_monitor = 1;
	pc_set_from_options(k->pc, prefix);
}
PetscErrorCode KSPSetFromOptions(KSP k) { ksp_set_from_options_prefixed(k, k->prefix); return 0; }

static void ksp_get_work(KSP k, PetscInt nw, PetscInt n)
{
	for (PetscInt i = 0; i < nw; i++) {
		if (k->work[i] && k->work[i]->n != n) VecDestroy(&k->work[i]);
		if (!k->work[i]) VecCreateSeq(0, n, &k->work[i]);
	}
}
static void ksp_log(KSP k, PetscReal rn)
{ if (k->res_hist && k->res_hist_max > k->res_hist_len) k->res_hist[k->res_hist_len++] = rn; }
static void ksp_monitor(KSP k, PetscInt it, PetscReal rn)
{
	for (PetscInt i = 0; i < k->nmon; i++) k->monitor[i](k, it, rn, k->mctx[i]);
	if (k->print_monitor && !g_quiet) printf("%3d KSP Residual norm %14.12e\n", it, rn);
}
/* KSPConvergedDefault (zero initial guess or default UIRNorm off: reference norm is ||b||-based rnorm at it 0) */
static KSPConvergedReason ksp_converged(KSP k, PetscInt it, PetscReal rn)
{
	if (k->normtype == KSP_NORM_NONE) return KSP_CONVERGED_ITERATING;   /* KSPConvergedSkip */
	if (it == 0) { k->rnorm0 = rn; k->ttol = fmax(k->rtol * rn, k->abstol); }
	if (rn != rn) return KSP_DIVERGED_DTOL;
	if (rn <= k->ttol) return (rn < k->abstol) ? KSP_CONVERGED_ATOL : KSP_CONVERGED_RTOL;
	if (rn >= k->dtol * k->rnorm0) return KSP_DIVERGED_DTOL;
	return KSP_CONVERGED_ITERATING;
}

/* PCApplyRichardson_SOR / _MG exist; Jacobi, ILU, LU, none do not */
static int pc_apply_richardson_exists(PC pc)
{ return strcmp(pc->type, PCSOR) == 0 || strcmp(pc->type, PCMG) == 0; }

static void pc_apply_richardson(PC pc, Vec b, Vec x, Vec w, PetscReal rtol, PetscReal abstol, PetscInt its, int guesszero,
                                PetscInt *outits, KSPConvergedReason *reason)
{
	if (strcmp(pc->type, PCSOR) == 0) {
		MatSORType st = pc->sym;
		if (guesszero) st = (MatSORType)(st | SOR_ZERO_INITIAL_GUESS);
		MatSOR(pc->A, b, pc->omega, st, pc->fshift, its * pc->its, pc->lits, x);
		*outits = its; *reason = KSP_CONVERGED_ITS;
		return;
	}
	/* PCApplyRichardson_MG */
	MGLevel *F = &pc->lev[pc->nlevels - 1];
	F->b = b; F->x = x;
	if (rtol) {
		PetscReal rn;
		if (guesszero) VecNorm(b, NORM_2, &rn);
		else { MatResidual(F->smooth->A, b, x, w); VecNorm(w, NORM_2, &rn); }
		pc->mg_ttol = fmax(rtol * rn, abstol);
	} else pc->mg_ttol = abstol;
	int conv = 0; PetscInt i;
	for (i = 0; i < its; i++) { pc_mg_cycle(pc, pc->nlevels - 1, &conv); if (conv) break; }
	*reason = conv ? KSP_CONVERGED_RTOL : KSP_CONVERGED_ITS;
	*outits = i;
}

/* KSPSolve_Richardson */
static void ksp_solve_richardson(KSP k)
{
	Vec x = k->vec_sol, b = k->vec_rhs;
	const PetscInt maxit = k->max_it;
	ksp_get_work(k, 2, b->n);
	Vec r = k->work[0], z = k->work[1];
	pc_setup(k->pc);
	if (pc_apply_richardson_exists(k->pc) && maxit > 0 && k->scale == 1.0 && k->nmon == 0 && !k->print_monitor) {
		pc_apply_richardson(k->pc, b, x, r, k->rtol, k->abstol, maxit, !k->guess_nonzero, &k->its, &k->reason);
		return;
	}
	if (k->guess_nonzero) { MatMult(k->A, x, r); VecAYPX(r, -1.0, b); }     /* r <- b - A x */
	else VecCopy(b, r);
	k->its = 0;
	PetscReal rnorm = 0.0;
	for (PetscInt i = 0; i < maxit; i++) {
		if (k->normtype == KSP_NORM_UNPRECONDITIONED) {
			VecNorm(r, NORM_2, &rnorm);
			ksp_monitor(k, i, rnorm); k->rnorm = rnorm; ksp_log(k, rnorm);
			k->reason = ksp_converged(k, i, rnorm);
			if (k->reason) break;
		}
		PCApply(k->pc, r, z);                                               /* z <- B r */
		if (k->normtype == KSP_NORM_PRECONDITIONED) {
			VecNorm(z, NORM_2, &rnorm);
			ksp_monitor(k, i, rnorm); k->rnorm = rnorm; ksp_log(k, rnorm);
			k->reason = ksp_converged(k, i, rnorm);
			if (k->reason) break;
		}
		VecAXPY(x, k->scale, z);                                            /* x <- x + scale z */
		k->its++;
		if (i + 1 < maxit || k->normtype != KSP_NORM_NONE) {
			MatMult(k->A, x, r); VecAYPX(r, -1.0, b);                       /* r <- b - A x */
		}
	}
	if (!k->reason) {
		if (k->normtype != KSP_NORM_NONE) {
			if (k->normtype == KSP_NORM_UNPRECONDITIONED) VecNorm(r, NORM_2, &rnorm);
			else { PCApply(k->pc, r, z); VecNorm(z, NORM_2, &rnorm); }
			k->rnorm = rnorm; ksp_log(k, rnorm); ksp_monitor(k, k->its, rnorm);
		}
		if (k->its >= k->max_it) {
			if (k->normtype != KSP_NORM_NONE) {
				k->reason = ksp_converged(k, k->its, rnorm);
				if (!k->reason) k->reason = KSP_DIVERGED_ITS;
			} else k->reason = KSP_CONVERGED_ITS;
		}
	}
}

/* KSPSolve_CG, left preconditioning; supports UNPRECONDITIONED (the reference's choice,
 * src/solver.c:1922), PRECONDITIONED and NONE norm types. */
static void ksp_solve_cg(KSP k)
{
	Vec X = k->vec_sol, B = k->vec_rhs;
	ksp_get_work(k, 4, B->n);
	Vec R = k->work[0], Z = k->work[1], P = k->work[2], W = k->work[3];
	PetscScalar a = 1.0, beta = 0.0, betaold = 1.0, b = 0.0, dpi = 0.0, dpiold;
	PetscReal dp = 0.0;
	pc_setup(k->pc);
	k->its = 0;
	if (k->guess_nonzero) { MatMult(k->A, X, R); VecAYPX(R, -1.0, B); } else VecCopy(B, R);
	switch (k->normtype) {
	case KSP_NORM_PRECONDITIONED: PCApply(k->pc, R, Z); VecNorm(Z, NORM_2, &dp); break;
	case KSP_NORM_UNPRECONDITIONED: VecNorm(R, NORM_2, &dp); break;
	case KSP_NORM_NONE: dp = 0.0; break;
	default: MP_ERR("KSPCG: norm type %d not restated", (int)k->normtype);
	}
	ksp_log(k, dp); ksp_monitor(k, 0, dp); k->rnorm = dp;
	k->reason = ksp_converged(k, 0, dp);
	if (k->reason) return;
	if (k->normtype != KSP_NORM_PRECONDITIONED) PCApply(k->pc, R, Z);       /* z <- B r */
	VecDot(Z, R, &beta);                                                    /* beta <- z'r */
	PetscInt i = 0;
	do {
		k->its = i + 1;
		if (beta == 0.0) { k->reason = KSP_CONVERGED_ATOL; break; }
		else if (i > 0 && beta * betaold < 0.0) { k->reason = KSP_DIVERGED_INDEFINITE_PC; break; }
		if (!i) { VecCopy(Z, P); b = 0.0; }                                 /* p <- z */
		else { b = beta / betaold; VecAYPX(P, b, Z); }                      /* p <- z + b p */
		dpiold = dpi;
		MatMult(k->A, P, W);                                                /* w <- A p */
		VecDot(P, W, &dpi);                                                 /* dpi <- p'w */
		betaold = beta;
		if (dpi == 0.0 || (i > 0 && dpi * dpiold <= 0.0)) { k->reason = KSP_DIVERGED_INDEFINITE_MAT; break; }
		a = beta / dpi;
		VecAXPY(X, a, P);                                                   /* x <- x + a p */
		VecAXPY(R, -a, W);                                                  /* r <- r - a w */
		if (k->normtype == KSP_NORM_PRECONDITIONED) { PCApply(k->pc, R, Z); VecNorm(Z, NORM_2, &dp); }
		else if (k->normtype == KSP_NORM_UNPRECONDITIONED) VecNorm(R, NORM_2, &dp);
		else dp = 0.0;
		k->rnorm = dp; ksp_log(k, dp); ksp_monitor(k, i + 1, dp);
		k->reason = ksp_converged(k, i + 1, dp);
		if (k->reason) break;
		if (k->normtype != KSP_NORM_PRECONDITIONED) PCApply(k->pc, R, Z);   /* z <- B r */
		VecDot(Z, R, &beta);
		i++;
	} while (i < k->max_it);
	if (i >= k->max_it && !k->reason) k->reason = (k->normtype == KSP_NORM_NONE) ? KSP_CONVERGED_ITS : KSP_DIVERGED_ITS;
}

static void ksp_solve_preonly(KSP k)
{
	if (k->guess_nonzero) MP_ERR("KSPPREONLY with a nonzero initial guess is not allowed");
	PCApply(k->pc, k->vec_rhs, k->vec_sol);
	k->its = 1; k->reason = KSP_CONVERGED_ITS;
}

PetscErrorCode KSPSolve(KSP k, Vec b, Vec x)
{
	if (!k->A) MP_ERR("KSPSolve: no operator");
	if (b == x) MP_ERR("KSPSolve: b and x must differ");
	k->vec_rhs = b; k->vec_sol = x;
	k->reason = KSP_CONVERGED_ITERATING; k->its = 0; k->res_hist_len = 0;
	if (!k->guess_nonzero) VecSet(x, 0.0);
	if (!strcmp(k->type, KSPRICHARDSON)) ksp_solve_richardson(k);
	else if (!strcmp(k->type, KSPCG)) ksp_solve_cg(k);
	else if (!strcmp(k->type, KSPPREONLY)) ksp_solve_preonly(k);
	else if (!strcmp(k->type, "chebyshev"))
		MP_ERR("KSP type 'chebyshev' (PCMG's default level smoother, with GMRES eigenvalue estimates on a random "
		       "right-hand side) is not reproducible and not restated: pass -mg_levels_ksp_type richardson "
		       "-mg_levels_pc_type {jacobi,sor} [-mg_levels_ksp_richardson_scale w] -mg_levels_ksp_max_it nu");
	else MP_ERR("KSP type '%s' is not restated in minipetsc (have: richardson cg preonly)", k->type);
	return 0;
}

/* KSPBuildResidualDefault: v = b - A x with the KSP's last rhs/solution */
PetscErrorCode KSPBuildResidual(KSP k, Vec t, Vec v, Vec *V)
{
	(void)t;
	if (!k->vec_sol || !k->vec_rhs) MP_ERR("KSPBuildResidual before KSPSolve");
	if (!v) MP_ERR("KSPBuildResidual: a result vector must be supplied");
	MatMult(k->A, k->vec_sol, v);
	VecAYPX(v, -1.0, k->vec_rhs);
	if (V) *V = v;
	return 0;
}

PetscErrorCode KSPView(KSP k, PetscViewer viewer)
{
	(void)viewer;
	if (g_quiet) return 0;
	PC pc = k->pc;
	printf("KSP Object: (%s) type: %s\n", k->prefix, k->type);
	if (!strcmp(k->type, KSPRICHARDSON)) printf("  Richardson: damping factor=%g\n", k->scale);
	printf("  maximum iterations=%d, %s initial guess\n", k->max_it, k->guess_nonzero ? "nonzero" : "zero");
	printf("  tolerances:  relative=%g, absolute=%g, divergence=%g\n", k->rtol, k->abstol, k->dtol);
	printf("  norm type: %d\n", (int)k->normtype);
	printf("PC Object: type: %s%s\n", pc->type_set ? pc->type : "ilu", pc->type_set ? "" : " (default)");
	if (!strcmp(pc->type, PCSOR)) printf("  SOR: type = %d, iterations = %d, local iterations = %d, omega = %g\n", (int)pc->sym, pc->its, pc->lits, pc->omega);
	if (!strcmp(pc->type, PCMG) && pc->lev) {
		printf("  MG: type is MULTIPLICATIVE, levels=%d cycles=v\n", pc->nlevels);
		for (PetscInt l = 0; l < pc->nlevels; l++) { printf("  -- level %d --\n", l); KSPView(pc->lev[l].smooth, viewer); }
	}
	if (k->A) printf("  linear system matrix: rows=%d, cols=%d\n", k->A->m, k->A->n);
	return 0;
}
