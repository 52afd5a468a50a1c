/*
 * petsc_b200.c -- the PETSc surface of the reference on the B200 engine.  See petsc_b200.h.
 *
 * Host C over the C-ABI of lib/libmgb200.so (include/mgb200_sparse.h).  A Vec is a dense vector in HBM with a lazily
 * synchronised host mirror (VecSetValue / VecGetArray work on the mirror; the next device operation uploads it), a Mat is
 * a host assembly slab until MatAssemblyEnd and a CSR matrix in HBM afterwards.  All linear algebra runs on the GPU; what
 * stays here is control flow, in PETSc's order [PETSc-upstream, version unpinned -- the same published algorithms the CPU
 * checker restates]:
 *   KSPSolve_Richardson (incl. the PCApplyRichardson short cut for SOR and MG), KSPSolve_CG, KSPSolve_PREONLY,
 *   PCApply for none / Jacobi / SOR / ILU(0) / LU / MG (PCMGMCycle_Private, multiplicative V), KSPConvergedDefault,
 *   KSPBuildResidual, the options database ("-key value", file first, argv wins), 1-rank MPI.
 * Reference call sites served: src/solver.c:185-253,489-620,1035-1209 (assembly), :1414-1575 (cycle 0), :1577-1882
 * (additive cycles), :1884-1989 (PCMG), :1991-2615 (I / E / D1 / D2 / D1PS cycles), src/poisson.c:29-214.
 */
#include "petsc_b200.h"
#include "../../../include/mgb200.h"
#include "../../../include/mgb200_sparse.h"
#include <stdarg.h>
#include <time.h>

#define PB_ERR(...) do { fprintf(stderr, "petsc_b200 error: " __VA_ARGS__); fprintf(stderr, "\n"); exit(78); } while (0)
#define GPU(call) do { if ((call) != MGB_OK) PB_ERR("%s: %s", #call, mgb_last_error()); } while (0)

long long PetscB200LaunchCount(void) { return mgb_sparse_launch_count(); }

/* ------------------------------------------------------------------------------------------------ 1-rank MPI */
int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
double MPI_Wtime(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; PB_ERR("MPI_Send: this layer runs on one rank"); return 1; }
int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; PB_ERR("MPI_Recv: this layer runs on one rank"); return 1; }

/* ------------------------------------------------------------------------------------------------ options database */
typedef struct { char *key, *val; } Opt;
static Opt *g_opt; static int g_nopt, g_capopt;

static void opt_put(const char *key, const char *val)
{
	for (int i = 0; i < g_nopt; i++)
		if (!strcmp(g_opt[i].key, key)) { free(g_opt[i].val); g_opt[i].val = val ? strdup(val) : NULL; return; }
	if (g_nopt == g_capopt) { g_capopt = g_capopt ? 2 * g_capopt : 32; g_opt = realloc(g_opt, (size_t)g_capopt * sizeof(Opt)); }
	g_opt[g_nopt].key = strdup(key); g_opt[g_nopt].val = val ? strdup(val) : NULL; g_nopt++;
}
static int looks_like_key(const char *t) { return t[0] == '-' && t[1] && !((t[1] >= '0' && t[1] <= '9') || t[1] == '.'); }
static void opt_tokens(char **tok, int n)
{
	for (int i = 0; i < n; i++) {
		if (!looks_like_key(tok[i])) continue;
		if (i + 1 < n && !looks_like_key(tok[i + 1])) { opt_put(tok[i] + 1, tok[i + 1]); i++; }
		else opt_put(tok[i] + 1, NULL);
	}
}
PetscErrorCode PetscOptionsInsertString(void *o, const char *str)
{
	(void)o;
	char *copy = strdup(str), **tok = NULL; int n = 0, cap = 0;
	for (char *p = strtok(copy, " \t\r\n"); p; p = strtok(NULL, " \t\r\n")) {
		if (n == cap) { cap = cap ? 2 * cap : 16; tok = realloc(tok, (size_t)cap * sizeof(char *)); }
		tok[n++] = p;
	}
	opt_tokens(tok, n);
	free(tok); free(copy);
	return 0;
}
PetscErrorCode PetscOptionsClear(void *o)
{ (void)o; for (int i = 0; i < g_nopt; i++) { free(g_opt[i].key); free(g_opt[i].val); } g_nopt = 0; return 0; }
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help)
{
	(void)help;
	if (file) {
		FILE *f = fopen(file, "r");                          /* a missing options file is not an error in PETSc */
		if (f) {
			char line[4096];
			while (fgets(line, sizeof line, f)) { char *h = strchr(line, '#'); if (h) *h = 0; PetscOptionsInsertString(NULL, line); }
			fclose(f);
		}
	}
	if (argc && argv && *argc > 1) opt_tokens(*argv + 1, *argc - 1);
	return 0;
}
PetscErrorCode PetscFinalize(void) { return PetscOptionsClear(NULL); }
static const char *opt_get(const char *pre, const char *name, int *found)
{
	char key[512];
	snprintf(key, sizeof key, "%s%s", pre ? pre : "", name[0] == '-' ? name + 1 : name);
	for (int i = 0; i < g_nopt; i++) if (!strcmp(g_opt[i].key, key)) { *found = 1; return g_opt[i].val; }
	*found = 0; return NULL;
}
PetscErrorCode PetscOptionsHasName(void *o, const char *pre, const char *name, PetscBool *set)
{ (void)o; int f; opt_get(pre, name, &f); *set = f ? PETSC_TRUE : PETSC_FALSE; return 0; }
PetscErrorCode PetscOptionsGetInt(void *o, const char *pre, const char *name, PetscInt *iv, PetscBool *set)
{
	(void)o; int f; const char *v = opt_get(pre, name, &f);
	if (f && v) *iv = (PetscInt)strtol(v, NULL, 10);
	if (set) *set = (f && v) ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
PetscErrorCode PetscOptionsGetReal(void *o, const char *pre, const char *name, PetscReal *dv, PetscBool *set)
{
	(void)o; int f; const char *v = opt_get(pre, name, &f);
	if (f && v) *dv = strtod(v, NULL);
	if (set) *set = (f && v) ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
PetscErrorCode PetscOptionsGetString(void *o, const char *pre, const char *name, char *str, size_t len, PetscBool *set)
{
	(void)o; int f; const char *v = opt_get(pre, name, &f);
	if (f && v) { strncpy(str, v, len - 1); str[len - 1] = 0; }
	if (set) *set = (f && v) ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
PetscErrorCode PetscOptionsGetIntArray(void *o, const char *pre, const char *name, PetscInt *iv, PetscInt *nmax, PetscBool *set)
{
	(void)o; int f; const char *v = opt_get(pre, name, &f);
	if (!(f && v)) { *nmax = 0; if (set) *set = PETSC_FALSE; return 0; }
	char *copy = strdup(v); int n = 0;
	for (char *p = strtok(copy, ","); p && n < *nmax; p = strtok(NULL, ",")) iv[n++] = (PetscInt)strtol(p, NULL, 10);
	free(copy);
	*nmax = n; if (set) *set = PETSC_TRUE;
	return 0;
}
static int opt_flag(const char *pre, const char *name)
{
	int f; const char *v = opt_get(pre, name, &f);
	if (!f) return 0;
	if (!v) return 1;
	return !(!strcmp(v, "0") || !strcmp(v, "false") || !strcmp(v, "no"));
}
PetscErrorCode PetscPrintf(MPI_Comm c, const char *fmt, ...) { (void)c; va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); return 0; }
PetscErrorCode PetscSynchronizedPrintf(MPI_Comm c, const char *fmt, ...) { (void)c; va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); return 0; }
PetscErrorCode PetscSynchronizedFlush(MPI_Comm c, FILE *fd) { (void)c; if (fd) fflush(fd); return 0; }
PetscErrorCode PetscLogStageRegister(const char *n, PetscLogStage *s) { (void)n; *s = 0; return 0; }
PetscErrorCode PetscLogStagePush(PetscLogStage s) { (void)s; return 0; }
PetscErrorCode PetscLogStagePop(void) { return 0; }

/* ------------------------------------------------------------------------------------------------ Vec */
struct _p_Vec {
	PetscInt n;
	mgb_dvec *d;                 /* the vector, in HBM */
	PetscScalar *h;              /* host mirror, allocated on first host access */
	int dev_ok, host_ok;         /* which copy is current (both may be) */
	PetscInt ranges[2];
};
static void v_dev(Vec x) { if (!x->dev_ok) { GPU(mgb_dvec_upload(x->d, x->h)); x->dev_ok = 1; } }
static void v_host(Vec x)
{
	if (!x->h) { x->h = calloc((size_t)(x->n > 0 ? x->n : 1), sizeof(PetscScalar)); if (!x->h) PB_ERR("out of host memory"); x->host_ok = 0; }
	if (!x->host_ok) { GPU(mgb_dvec_download(x->d, x->h)); x->host_ok = 1; }
}
static void v_wrote_dev(Vec x) { x->dev_ok = 1; x->host_ok = 0; }
static void v_wrote_host(Vec x) { x->host_ok = 1; x->dev_ok = 0; }

PetscErrorCode VecCreateSeq(MPI_Comm c, PetscInt n, Vec *v)
{
	(void)c;
	Vec x = calloc(1, sizeof *x);
	x->n = n; GPU(mgb_dvec_create(n, &x->d));
	x->dev_ok = 1; x->host_ok = 0; x->ranges[0] = 0; x->ranges[1] = n;
	*v = x; return 0;
}
PetscErrorCode VecDuplicate(Vec v, Vec *nv) { return VecCreateSeq(0, v->n, nv); }
PetscErrorCode VecDestroy(Vec *v) { if (v && *v) { mgb_dvec_destroy((*v)->d); free((*v)->h); free(*v); *v = NULL; } return 0; }
PetscErrorCode VecGetSize(Vec x, PetscInt *n) { *n = x->n; return 0; }
PetscErrorCode VecGetArray(Vec x, PetscScalar **a) { v_host(x); v_wrote_host(x); *a = x->h; return 0; }   /* the caller may write */
PetscErrorCode VecRestoreArray(Vec x, PetscScalar **a) { (void)x; if (a) *a = NULL; return 0; }
PetscErrorCode VecGetOwnershipRange(Vec x, PetscInt *lo, PetscInt *hi) { if (lo) *lo = 0; if (hi) *hi = x->n; return 0; }
PetscErrorCode VecGetOwnershipRanges(Vec x, const PetscInt *r[]) { *r = x->ranges; return 0; }
PetscErrorCode VecAssemblyBegin(Vec x) { (void)x; return 0; }
PetscErrorCode VecAssemblyEnd(Vec x) { (void)x; return 0; }
PetscErrorCode VecSetValue(Vec x, PetscInt row, PetscScalar v, InsertMode m)
{
	if (row < 0 || row >= x->n) PB_ERR("VecSetValue: row %d out of range [0,%d)", row, x->n);
	v_host(x);
	if (m == ADD_VALUES) x->h[row] += v; else x->h[row] = v;
	v_wrote_host(x);
	return 0;
}
PetscErrorCode VecSet(Vec x, PetscScalar a) { GPU(mgb_dvec_set(x->d, a)); v_wrote_dev(x); return 0; }
PetscErrorCode VecCopy(Vec x, Vec y) { if (x == y) return 0; v_dev(x); GPU(mgb_dvec_copy(y->d, x->d)); v_wrote_dev(y); return 0; }
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val) { v_dev(x); v_dev(y); GPU(mgb_dvec_dot(x->d, y->d, val)); return 0; }
PetscErrorCode VecTDot(Vec x, Vec y, PetscScalar *val) { return VecDot(x, y, val); }
PetscErrorCode VecNorm(Vec x, NormType t, PetscReal *val)
{
	v_dev(x);
	GPU(mgb_dvec_norm(x->d, t == NORM_1 ? MGB_NORM_1 : (t == NORM_INFINITY ? MGB_NORM_INF : MGB_NORM_2), val));
	return 0;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x) { v_dev(x); v_dev(y); GPU(mgb_dvec_axpy(y->d, a, x->d)); v_wrote_dev(y); return 0; }
PetscErrorCode VecAYPX(Vec y, PetscScalar b, Vec x) { v_dev(x); v_dev(y); GPU(mgb_dvec_aypx(y->d, b, x->d)); v_wrote_dev(y); return 0; }
PetscErrorCode VecWAXPY(Vec w, PetscScalar a, Vec x, Vec y) { v_dev(x); v_dev(y); GPU(mgb_dvec_waxpy(w->d, a, x->d, y->d)); v_wrote_dev(w); return 0; }
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar a, PetscScalar b, PetscScalar g, Vec x, Vec y)
{ v_dev(x); v_dev(y); v_dev(z); GPU(mgb_dvec_axpbypcz(z->d, a, b, g, x->d, y->d)); v_wrote_dev(z); return 0; }
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y) { v_dev(x); v_dev(y); GPU(mgb_dvec_pointwise_mult(w->d, x->d, y->d)); v_wrote_dev(w); return 0; }
PetscErrorCode VecScale(Vec x, PetscScalar a) { v_dev(x); GPU(mgb_dvec_scale(x->d, a)); v_wrote_dev(x); return 0; }
PetscErrorCode VecView(Vec x, PetscViewer v) { (void)v; v_host(x); for (PetscInt i = 0; i < x->n; i++) printf("%g\n", x->h[i]); return 0; }

/* ------------------------------------------------------------------------------------------------ IS, sub-vectors */
struct _p_IS { PetscInt n; PetscInt *idx; mgb_dindex *d; };
PetscErrorCode ISCreateGeneral(MPI_Comm c, PetscInt n, const PetscInt idx[], PetscCopyMode mode, IS *is)
{
	(void)c; (void)mode;
	IS s = calloc(1, sizeof *s);
	s->n = n; s->idx = malloc((size_t)(n > 0 ? n : 1) * sizeof(PetscInt));
	memcpy(s->idx, idx, (size_t)n * sizeof(PetscInt));
	GPU(mgb_dindex_create(n, s->idx, &s->d));
	*is = s; return 0;
}
PetscErrorCode ISDestroy(IS *is) { if (is && *is) { mgb_dindex_destroy((*is)->d); free((*is)->idx); free(*is); *is = NULL; } return 0; }
PetscErrorCode ISView(IS is, PetscViewer v) { (void)v; for (PetscInt i = 0; i < is->n; i++) printf("%d %d\n", i, is->idx[i]); return 0; }
/* a sub-vector is a gathered copy that VecRestoreSubVector scatters back */
PetscErrorCode VecGetSubVector(Vec x, IS is, Vec *y)
{
	for (PetscInt i = 0; i < is->n; i++) if (is->idx[i] < 0 || is->idx[i] >= x->n) PB_ERR("VecGetSubVector: index %d out of range", is->idx[i]);
	VecCreateSeq(0, is->n, y);
	v_dev(x); GPU(mgb_dvec_gather((*y)->d, x->d, is->d)); v_wrote_dev(*y);
	return 0;
}
PetscErrorCode VecRestoreSubVector(Vec x, IS is, Vec *y)
{
	if (!y || !*y) return 0;
	v_dev(x); v_dev(*y); GPU(mgb_dvec_scatter(x->d, (*y)->d, is->d)); v_wrote_dev(x);
	return VecDestroy(y);
}

/* ------------------------------------------------------------------------------------------------ Mat (SeqAIJ) */
struct _p_Mat {
	PetscInt m, n;
	PetscInt stride, *bcol, *blen; PetscScalar *bval;      /* assembly slab: `stride` sorted entries per row, regrown on overflow */
	int assembled;
	PetscInt *ia, *ja; PetscScalar *va;                    /* host CSR (pattern + values as assembled; MatMatMult, MatView) */
	mgb_dcsr *d;                                           /* the matrix, in HBM */
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
	if (!a->bcol || !a->bval || !a->blen) PB_ERR("MatCreateSeqAIJ: out of memory (m=%d nz=%d)", m, nz);
	*A = a; return 0;
}
PetscErrorCode MatCreateAIJ(MPI_Comm c, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscInt d_nz, const PetscInt d_nnz[],
                            PetscInt o_nz, const PetscInt o_nnz[], Mat *A)
{ (void)o_nz; (void)o_nnz; if (m < 0) m = M; if (n < 0) n = N; return MatCreateSeqAIJ(c, m, n, d_nz, d_nnz, A); }
PetscErrorCode MatSetValue(Mat a, PetscInt row, PetscInt col, PetscScalar v, InsertMode mode)
{
	if (a->assembled) PB_ERR("MatSetValue after MatAssemblyEnd is not offered");
	if (row < 0 || row >= a->m || col < 0 || col >= a->n) PB_ERR("MatSetValue: (%d,%d) outside %dx%d", row, col, a->m, a->n);
	PetscInt *c = a->bcol + (size_t)row * a->stride; PetscScalar *w = a->bval + (size_t)row * a->stride;
	PetscInt len = a->blen[row], k = len;
	while (k > 0 && c[k - 1] >= col) k--;
	if (k < len && c[k] == col) { if (mode == ADD_VALUES) w[k] += v; else w[k] = v; return 0; }
	if (len == a->stride) {
		const PetscInt ns = 2 * a->stride;
		PetscInt *nc = malloc((size_t)a->m * (size_t)ns * sizeof(PetscInt) + 8);
		PetscScalar *nv = malloc((size_t)a->m * (size_t)ns * sizeof(PetscScalar) + 8);
		if (!nc || !nv) PB_ERR("MatSetValue: out of memory");
		for (PetscInt i = 0; i < a->m; i++) {
			memcpy(nc + (size_t)i * ns, a->bcol + (size_t)i * a->stride, (size_t)a->blen[i] * sizeof(PetscInt));
			memcpy(nv + (size_t)i * ns, a->bval + (size_t)i * a->stride, (size_t)a->blen[i] * sizeof(PetscScalar));
		}
		free(a->bcol); free(a->bval); a->bcol = nc; a->bval = nv; a->stride = ns;
		c = a->bcol + (size_t)row * a->stride; w = a->bval + (size_t)row * a->stride;
	}
	for (PetscInt t = len; t > k; t--) { c[t] = c[t - 1]; w[t] = w[t - 1]; }
	c[k] = col; w[k] = v; a->blen[row] = len + 1;
	return 0;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
static void mat_finish(Mat a)       /* host CSR complete: put the matrix into HBM */
{
	GPU(mgb_dcsr_create(a->m, a->n, a->ia, a->ja, a->va, &a->d));
	a->assembled = 1;
}
PetscErrorCode MatAssemblyEnd(Mat a, MatAssemblyType t)
{
	if (t != MAT_FINAL_ASSEMBLY || a->assembled) return 0;
	long long nnz = 0;
	a->ia = malloc(((size_t)a->m + 1) * sizeof(PetscInt));
	for (PetscInt i = 0; i < a->m; i++) { a->ia[i] = (PetscInt)nnz; nnz += a->blen[i]; }
	if (nnz > 2147483647LL) PB_ERR("MatAssemblyEnd: nnz overflows 32-bit PetscInt");
	a->ia[a->m] = (PetscInt)nnz;
	a->ja = malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(PetscInt));
	a->va = malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(PetscScalar));
	if (!a->ja || !a->va) PB_ERR("MatAssemblyEnd: out of memory");
	for (PetscInt i = 0; i < a->m; i++) {
		memcpy(a->ja + a->ia[i], a->bcol + (size_t)i * a->stride, (size_t)a->blen[i] * sizeof(PetscInt));
		memcpy(a->va + a->ia[i], a->bval + (size_t)i * a->stride, (size_t)a->blen[i] * sizeof(PetscScalar));
	}
	free(a->bcol); free(a->bval); free(a->blen); a->bcol = NULL; a->bval = NULL; a->blen = NULL;
	mat_finish(a);
	return 0;
}
PetscErrorCode MatDestroy(Mat *A)
{
	if (!A || !*A) return 0;
	Mat a = *A;
	free(a->bcol); free(a->bval); free(a->blen); free(a->ia); free(a->ja); free(a->va);
	mgb_dcsr_destroy(a->d);
	free(a); *A = NULL; return 0;
}
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n) { if (m) *m = A->m; if (n) *n = A->n; return 0; }
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left) { if (right) VecCreateSeq(0, A->n, right); if (left) VecCreateSeq(0, A->m, left); return 0; }
static void need_assembled(Mat A, const char *who) { if (!A->assembled) PB_ERR("%s: matrix not assembled", who); }
PetscErrorCode MatMult(Mat A, Vec x, Vec y)
{
	need_assembled(A, "MatMult");
	if (x->n != A->n || y->n != A->m) PB_ERR("MatMult: size mismatch (A %dx%d, x %d, y %d)", A->m, A->n, x->n, y->n);
	v_dev(x); GPU(mgb_dcsr_mult(A->d, x->d, y->d)); v_wrote_dev(y);
	return 0;
}
PetscErrorCode MatMultAdd(Mat A, Vec x, Vec y, Vec z)
{
	need_assembled(A, "MatMultAdd");
	v_dev(x); v_dev(y); GPU(mgb_dcsr_mult_add(A->d, x->d, y->d, z->d)); v_wrote_dev(z);
	return 0;
}
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y)
{ (void)A; (void)x; (void)y; PB_ERR("MatMultTranspose is not offered (the reference passes explicit restriction matrices, src/solver.c:1035-1094)"); return 1; }
PetscErrorCode MatResidual(Mat A, Vec b, Vec x, Vec r) { MatMult(A, x, r); return VecAYPX(r, -1.0, b); }    /* r = b - A x */
PetscErrorCode MatRestrict(Mat A, Vec x, Vec y) { if (A->n == x->n) return MatMult(A, x, y); return MatMultTranspose(A, x, y); }
PetscErrorCode MatInterpolateAdd(Mat A, Vec x, Vec y, Vec w) { if (A->n != x->n) PB_ERR("MatInterpolateAdd with a transposed interpolation is not offered"); return MatMultAdd(A, x, y, w); }
PetscErrorCode MatScale(Mat A, PetscScalar s)
{
	need_assembled(A, "MatScale");
	for (PetscInt k = 0; k < A->ia[A->m]; k++) A->va[k] *= s;
	GPU(mgb_dcsr_scale(A->d, s));
	return 0;
}
PetscErrorCode MatGetDiagonal(Mat A, Vec d)
{
	need_assembled(A, "MatGetDiagonal");
	v_host(d);
	for (PetscInt i = 0; i < A->m; i++) {
		d->h[i] = 0.0;
		for (PetscInt k = A->ia[i]; k < A->ia[i + 1]; k++) if (A->ja[k] == i) d->h[i] = A->va[k];
	}
	v_wrote_host(d);
	return 0;
}
/* C = A B, set up once per solve by the additive cycle's filter (src/solver.c:1759): symbolic + numeric product on the
 * host pattern, the products accumulated in (row of A, entry of B's row) order; the result lives in HBM like any Mat */
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat *C)
{
	(void)scall; (void)fill;
	need_assembled(A, "MatMatMult"); need_assembled(B, "MatMatMult");
	if (A->n != B->m) PB_ERR("MatMatMult: size mismatch");
	Mat c; MatCreateSeqAIJ(0, A->m, B->n, 16, NULL, &c);
	for (PetscInt i = 0; i < A->m; i++)
		for (PetscInt k = A->ia[i]; k < A->ia[i + 1]; k++) {
			const PetscInt r = A->ja[k];
			for (PetscInt t = B->ia[r]; t < B->ia[r + 1]; t++) MatSetValue(c, i, B->ja[t], A->va[k] * B->va[t], ADD_VALUES);
		}
	MatAssemblyEnd(c, MAT_FINAL_ASSEMBLY);
	*C = c; return 0;
}
PetscErrorCode MatSOR(Mat A, Vec b, PetscReal omega, MatSORType flag, PetscReal shift, PetscInt its, PetscInt lits, Vec x)
{
	need_assembled(A, "MatSOR");
	v_dev(b); v_dev(x); GPU(mgb_dcsr_sor(A->d, b->d, omega, (int)flag, shift, its, lits, x->d)); v_wrote_dev(x);
	return 0;
}
PetscErrorCode MatView(Mat A, PetscViewer v)
{
	if (v != PETSC_VIEWER_STDOUT_WORLD) return 0;
	need_assembled(A, "MatView");
	for (PetscInt i = 0; i < A->m; i++) {
		printf("row %d:", i);
		for (PetscInt k = A->ia[i]; k < A->ia[i + 1]; k++) printf(" (%d, %g) ", A->ja[k], A->va[k]);
		printf("\n");
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------------ PC, KSP objects */
typedef struct { KSP smooth; Mat restrct, interpolate; Vec b, x, r; int own_b, own_x, own_r; } MGLevel;
struct _p_PC {
	char type[32], prefix[128]; int type_set, setup;
	Mat A;
	Vec dinv;                                                  /* Jacobi */
	PetscReal omega, fshift; PetscInt its, lits; MatSORType sym;   /* SOR */
	PetscInt nlevels, cycles; MGLevel *lev; PetscReal mg_ttol; /* MG: level 0 = coarsest, as in PETSc */
};
struct _p_KSP {
	char type[32], prefix[128];
	Mat A; PC pc;
	KSPNormType normtype;
	PetscReal rtol, abstol, dtol, scale; PetscInt max_it;
	PetscBool guess_nonzero;
	Vec sol, rhs, work[4];
	PetscInt its; KSPConvergedReason reason; PetscReal rnorm, rnorm0, ttol;
	PetscReal *hist; PetscInt hist_len, hist_max;
	PetscErrorCode (*mon[4])(KSP, PetscInt, PetscReal, void *); void *mctx[4]; PetscInt nmon;
	int print_monitor;
};

PetscErrorCode PCSetType(PC pc, PCType t)
{
	if (pc->type_set && !strcmp(pc->type, t)) return 0;
	strncpy(pc->type, t, sizeof pc->type - 1); pc->type_set = 1; pc->setup = 0;
	return 0;
}
PetscErrorCode KSPCreate(MPI_Comm c, KSP *out)
{
	(void)c;
	KSP k = calloc(1, sizeof *k);
	strcpy(k->type, KSPGMRES);                                 /* PETSc's default; the reference always sets a type */
	k->pc = calloc(1, sizeof *k->pc);
	k->pc->omega = 1.0; k->pc->its = 1; k->pc->lits = 1; k->pc->sym = SOR_LOCAL_SYMMETRIC_SWEEP; k->pc->cycles = 1;
	k->normtype = KSP_NORM_PRECONDITIONED;
	k->rtol = 1e-5; k->abstol = 1e-50; k->dtol = 1e4; k->max_it = 10000; k->scale = 1.0;
	*out = k; return 0;
}
PetscErrorCode KSPDestroy(KSP *pk)
{
	if (!pk || !*pk) return 0;
	KSP k = *pk; PC pc = k->pc;
	for (int i = 0; i < 4; i++) VecDestroy(&k->work[i]);
	VecDestroy(&pc->dinv);
	if (pc->lev) {
		for (PetscInt l = 0; l < pc->nlevels; l++) {
			KSPDestroy(&pc->lev[l].smooth);
			if (pc->lev[l].own_b) VecDestroy(&pc->lev[l].b);
			if (pc->lev[l].own_x) VecDestroy(&pc->lev[l].x);
			if (pc->lev[l].own_r) VecDestroy(&pc->lev[l].r);
		}
		free(pc->lev);
	}
	free(pc); free(k); *pk = NULL; return 0;
}
PetscErrorCode PetscObjectSetOptionsPrefix(PetscObject obj, const char *prefix)
{ KSP k = (KSP)obj; strncpy(k->prefix, prefix ? prefix : "", sizeof k->prefix - 1); return 0; }   /* only KSPs get one (src/solver.c:1624-1643) */
PetscErrorCode KSPSetType(KSP k, KSPType t) { strncpy(k->type, t, sizeof k->type - 1); return 0; }
PetscErrorCode KSPSetOperators(KSP k, Mat A, Mat P) { (void)P; k->A = A; if (k->pc->A != A) { k->pc->A = A; k->pc->setup = 0; } return 0; }
PetscErrorCode KSPSetNormType(KSP k, KSPNormType t) { k->normtype = t; return 0; }
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
PetscErrorCode KSPSetResidualHistory(KSP k, PetscReal a[], PetscInt na, PetscBool reset) { (void)reset; k->hist = a; k->hist_max = na; k->hist_len = 0; return 0; }
PetscErrorCode KSPGetResidualHistory(KSP k, PetscReal *a[], PetscInt *na) { if (a) *a = k->hist; if (na) *na = k->hist_len; return 0; }
PetscErrorCode KSPMonitorSet(KSP k, PetscErrorCode (*m)(KSP, PetscInt, PetscReal, void *), void *ctx, PetscErrorCode (*d)(void **))
{ (void)d; if (k->nmon >= 4) PB_ERR("KSPMonitorSet: too many monitors"); k->mon[k->nmon] = m; k->mctx[k->nmon] = ctx; k->nmon++; return 0; }

static void ksp_options(KSP k, const char *pre)
{
	char buf[64]; PetscBool set; PC pc = k->pc;
	PetscOptionsGetString(NULL, pre, "-ksp_type", buf, sizeof buf, &set); if (set) KSPSetType(k, buf);
	PetscOptionsGetInt(NULL, pre, "-ksp_max_it", &k->max_it, NULL);
	PetscOptionsGetReal(NULL, pre, "-ksp_rtol", &k->rtol, NULL);
	PetscOptionsGetReal(NULL, pre, "-ksp_atol", &k->abstol, NULL);
	PetscOptionsGetReal(NULL, pre, "-ksp_divtol", &k->dtol, NULL);
	PetscOptionsGetReal(NULL, pre, "-ksp_richardson_scale", &k->scale, NULL);
	PetscOptionsGetString(NULL, pre, "-ksp_norm_type", buf, sizeof buf, &set);
	if (set) {
		if (!strcmp(buf, "none")) k->normtype = KSP_NORM_NONE;
		else if (!strcmp(buf, "preconditioned")) k->normtype = KSP_NORM_PRECONDITIONED;
		else if (!strcmp(buf, "unpreconditioned")) k->normtype = KSP_NORM_UNPRECONDITIONED;
		else if (!strcmp(buf, "natural")) k->normtype = KSP_NORM_NATURAL;
		else PB_ERR("unknown -ksp_norm_type %s", buf);
	}
	int f; opt_get(pre, "-ksp_initial_guess_nonzero", &f);
	if (f) k->guess_nonzero = opt_flag(pre, "-ksp_initial_guess_nonzero") ? PETSC_TRUE : PETSC_FALSE;
	if (opt_flag(pre, "-ksp_monitor")) k->print_monitor = 1;
	strncpy(pc->prefix, pre, sizeof pc->prefix - 1);
	PetscOptionsGetString(NULL, pre, "-pc_type", buf, sizeof buf, &set); if (set) PCSetType(pc, buf);
	PetscOptionsGetReal(NULL, pre, "-pc_sor_omega", &pc->omega, NULL);
	PetscOptionsGetInt(NULL, pre, "-pc_sor_its", &pc->its, NULL);
	PetscOptionsGetInt(NULL, pre, "-pc_sor_lits", &pc->lits, NULL);
	if (opt_flag(pre, "-pc_sor_symmetric")) pc->sym = SOR_SYMMETRIC_SWEEP;
	if (opt_flag(pre, "-pc_sor_backward")) pc->sym = SOR_BACKWARD_SWEEP;
	if (opt_flag(pre, "-pc_sor_forward")) pc->sym = SOR_FORWARD_SWEEP;
	if (opt_flag(pre, "-pc_sor_local_symmetric")) pc->sym = SOR_LOCAL_SYMMETRIC_SWEEP;
	if (opt_flag(pre, "-pc_sor_local_backward")) pc->sym = SOR_LOCAL_BACKWARD_SWEEP;
	if (opt_flag(pre, "-pc_sor_local_forward")) pc->sym = SOR_LOCAL_FORWARD_SWEEP;
}
PetscErrorCode KSPSetFromOptions(KSP k) { ksp_options(k, k->prefix); return 0; }

/* ------------------------------------------------------------------------------------------------ PCMG */
PetscErrorCode PCMGSetLevels(PC pc, PetscInt levels, MPI_Comm *comms)
{
	(void)comms;
	if (pc->lev) PB_ERR("PCMGSetLevels called twice");
	pc->nlevels = levels; pc->lev = calloc((size_t)levels, sizeof(MGLevel));
	for (PetscInt l = 0; l < levels; l++) {
		KSP k; KSPCreate(0, &k);
		pc->lev[l].smooth = k; k->normtype = KSP_NORM_NONE;
		if (l == 0) { KSPSetType(k, KSPPREONLY); PCSetType(k->pc, PCLU); k->max_it = 1; }       /* coarse: preonly + LU (mg_coarse_) */
		else { KSPSetType(k, "chebyshev"); PCSetType(k->pc, PCSOR); k->max_it = 2; }             /* PETSc's default smoother (mg_levels_) */
	}
	return 0;
}
PetscErrorCode PCMGGetCoarseSolve(PC pc, KSP *k) { *k = pc->lev[0].smooth; return 0; }
PetscErrorCode PCMGGetSmoother(PC pc, PetscInt l, KSP *k) { *k = pc->lev[l].smooth; return 0; }
PetscErrorCode PCMGSetInterpolation(PC pc, PetscInt l, Mat m) { if (l <= 0) PB_ERR("PCMGSetInterpolation on level 0"); pc->lev[l].interpolate = m; return 0; }
PetscErrorCode PCMGSetRestriction(PC pc, PetscInt l, Mat m) { if (l <= 0) PB_ERR("PCMGSetRestriction on level 0"); pc->lev[l].restrct = m; return 0; }
PetscErrorCode PCMGSetR(PC pc, PetscInt l, Vec c) { pc->lev[l].r = c; return 0; }
PetscErrorCode PCMGSetRhs(PC pc, PetscInt l, Vec c) { pc->lev[l].b = c; return 0; }
PetscErrorCode PCMGSetX(PC pc, PetscInt l, Vec c) { pc->lev[l].x = c; return 0; }
PetscErrorCode PCMGSetNumberSmoothUp(PC pc, PetscInt n) { for (PetscInt l = 1; l < pc->nlevels; l++) pc->lev[l].smooth->max_it = n; return 0; }
PetscErrorCode PCMGSetNumberSmoothDown(PC pc, PetscInt n) { return PCMGSetNumberSmoothUp(pc, n); }
PetscErrorCode PCASMSetType(PC pc, PCASMType t) { (void)pc; (void)t; PB_ERR("PCASM is not offered (the reference's calls are commented out)"); return 1; }
PetscErrorCode PCASMSetOverlap(PC pc, PetscInt o) { (void)pc; (void)o; PB_ERR("PCASM is not offered"); return 1; }
PetscErrorCode PCASMSetTotalSubdomains(PC pc, PetscInt N, IS a[], IS b[]) { (void)pc; (void)N; (void)a; (void)b; PB_ERR("PCASM is not offered"); return 1; }

static void pc_setup_mg(PC pc)
{
	const PetscInt n = pc->nlevels;
	if (!pc->lev) PB_ERR("PCMG: PCMGSetLevels was not called");
	for (PetscInt l = 0; l < n; l++) {
		MGLevel *L = &pc->lev[l]; char pre[160];
		snprintf(pre, sizeof pre, l == 0 ? "%smg_coarse_" : "%smg_levels_", pc->prefix);
		ksp_options(L->smooth, pre);
		if (l > 0) { snprintf(pre, sizeof pre, "%smg_levels_%d_", pc->prefix, l); ksp_options(L->smooth, pre); }
		if (!L->smooth->A) {
			if (l == n - 1) KSPSetOperators(L->smooth, pc->A, pc->A);
			else PB_ERR("PCMG: no operator on level %d (Galerkin coarsening is not offered)", l);
		}
		if (l > 0 && (!L->restrct || !L->interpolate)) PB_ERR("PCMG: missing restriction / interpolation on level %d", l);
		const PetscInt m = L->smooth->A->m;
		if (l < n - 1) {
			if (!L->b) { VecCreateSeq(0, m, &L->b); L->own_b = 1; }
			if (!L->x) { VecCreateSeq(0, m, &L->x); L->own_x = 1; }
		}
		if (l > 0 && !L->r) { VecCreateSeq(0, m, &L->r); L->own_r = 1; }
		if (l > 0) KSPSetInitialGuessNonzero(L->smooth, PETSC_TRUE);     /* PCSetUp_MG: the level smoother continues from x */
	}
}
static void pc_setup(PC pc)
{
	if (pc->setup) return;
	if (!pc->A) PB_ERR("PCSetUp: no operator set");
	if (!pc->type_set) { strcpy(pc->type, PCILU); pc->type_set = 1; }    /* PETSc's default for AIJ on one rank */
	if (!strcmp(pc->type, PCJACOBI)) {
		VecDestroy(&pc->dinv); VecCreateSeq(0, pc->A->m, &pc->dinv);
		need_assembled(pc->A, "PCSetUp_Jacobi");
		GPU(mgb_dcsr_inverse_diagonal(pc->A->d, pc->dinv->d)); v_wrote_dev(pc->dinv);
	} else if (!strcmp(pc->type, PCILU)) { need_assembled(pc->A, "PCSetUp_ILU"); GPU(mgb_dcsr_ilu0_factor(pc->A->d)); }
	else if (!strcmp(pc->type, PCLU)) { need_assembled(pc->A, "PCSetUp_LU"); GPU(mgb_dcsr_lu_factor(pc->A->d)); }
	else if (!strcmp(pc->type, PCMG)) pc_setup_mg(pc);
	else if (!strcmp(pc->type, PCSOR) || !strcmp(pc->type, PCNONE)) { /* nothing to build */ }
	else PB_ERR("PC type '%s' is not offered (have: none jacobi sor ilu lu mg)", pc->type);
	pc->setup = 1;
}
static void mg_cycle(PC pc, PetscInt l, int *converged)           /* PCMGMCycle_Private */
{
	MGLevel *L = &pc->lev[l];
	KSPSolve(L->smooth, L->b, L->x);                                /* pre-smooth (level 0: the coarse solve) */
	if (l > 0) {
		MatResidual(L->smooth->A, L->b, L->x, L->r);
		if (l == pc->nlevels - 1 && pc->mg_ttol > 0.0 && converged) {
			PetscReal rn; VecNorm(L->r, NORM_2, &rn);
			if (rn <= pc->mg_ttol) { *converged = 1; return; }
		}
		MGLevel *C = &pc->lev[l - 1];
		MatRestrict(L->restrct, L->r, C->b);
		VecSet(C->x, 0.0);
		PetscInt cyc = (l == 1) ? 1 : pc->cycles;
		while (cyc--) mg_cycle(pc, l - 1, converged);
		MatInterpolateAdd(L->interpolate, C->x, L->x, L->x);
		KSPSolve(L->smooth, L->b, L->x);                            /* post-smooth */
	}
}
PetscErrorCode PCApply(PC pc, Vec x, Vec y)
{
	pc_setup(pc);
	if (!strcmp(pc->type, PCNONE)) return VecCopy(x, y);
	if (!strcmp(pc->type, PCJACOBI)) return VecPointwiseMult(y, x, pc->dinv);
	if (!strcmp(pc->type, PCSOR)) return MatSOR(pc->A, x, pc->omega, (MatSORType)(pc->sym | SOR_ZERO_INITIAL_GUESS), pc->fshift, pc->its, pc->lits, y);
	if (!strcmp(pc->type, PCILU)) { v_dev(x); GPU(mgb_dcsr_ilu0_solve(pc->A->d, x->d, y->d)); v_wrote_dev(y); return 0; }
	if (!strcmp(pc->type, PCLU)) { v_dev(x); GPU(mgb_dcsr_lu_solve(pc->A->d, x->d, y->d)); v_wrote_dev(y); return 0; }
	if (!strcmp(pc->type, PCMG)) {                                  /* PCApply_MG, multiplicative: x = 0, one cycle */
		MGLevel *F = &pc->lev[pc->nlevels - 1];
		F->b = x; F->x = y; VecSet(y, 0.0); pc->mg_ttol = 0.0;
		mg_cycle(pc, pc->nlevels - 1, NULL);
		return 0;
	}
	PB_ERR("PCApply: unknown type %s", pc->type);
	return 1;
}

/* ------------------------------------------------------------------------------------------------ KSPSolve */
static void ksp_work(KSP k, PetscInt nw, PetscInt n)
{
	for (PetscInt i = 0; i < nw; i++) {
		if (k->work[i] && k->work[i]->n != n) VecDestroy(&k->work[i]);
		if (!k->work[i]) VecCreateSeq(0, n, &k->work[i]);
	}
}
static void ksp_log(KSP k, PetscReal rn) { if (k->hist && k->hist_max > k->hist_len) k->hist[k->hist_len++] = rn; }
static void ksp_monitor(KSP k, PetscInt it, PetscReal rn)
{
	for (PetscInt i = 0; i < k->nmon; i++) k->mon[i](k, it, rn, k->mctx[i]);
	if (k->print_monitor) printf("%3d KSP Residual norm %14.12e\n", it, rn);
}
static KSPConvergedReason ksp_converged(KSP k, PetscInt it, PetscReal rn)      /* KSPConvergedDefault / Skip */
{
	if (k->normtype == KSP_NORM_NONE) return KSP_CONVERGED_ITERATING;
	if (it == 0) { k->rnorm0 = rn; k->ttol = fmax(k->rtol * rn, k->abstol); }
	if (rn != rn) return KSP_DIVERGED_DTOL;
	if (rn <= k->ttol) return (rn < k->abstol) ? KSP_CONVERGED_ATOL : KSP_CONVERGED_RTOL;
	if (rn >= k->dtol * k->rnorm0) return KSP_DIVERGED_DTOL;
	return KSP_CONVERGED_ITERATING;
}
static void solve_richardson(KSP k)
{
	Vec x = k->sol, b = k->rhs; PC pc = k->pc;
	const PetscInt maxit = k->max_it;
	ksp_work(k, 2, b->n);
	Vec r = k->work[0], z = k->work[1];
	pc_setup(pc);
	const int has_apply_richardson = !strcmp(pc->type, PCSOR) || !strcmp(pc->type, PCMG);
	if (has_apply_richardson && maxit > 0 && k->scale == 1.0 && k->nmon == 0 && !k->print_monitor) {
		if (!strcmp(pc->type, PCSOR)) {                             /* PCApplyRichardson_SOR: the whole loop is one MatSOR */
			MatSORType st = pc->sym;
			if (!k->guess_nonzero) st = (MatSORType)(st | SOR_ZERO_INITIAL_GUESS);
			MatSOR(pc->A, b, pc->omega, st, pc->fshift, maxit * pc->its, pc->lits, x);
			k->its = maxit; k->reason = KSP_CONVERGED_ITS;
			return;
		}
		MGLevel *F = &pc->lev[pc->nlevels - 1];                     /* PCApplyRichardson_MG */
		F->b = b; F->x = x;
		if (k->rtol) {
			PetscReal rn;
			if (!k->guess_nonzero) VecNorm(b, NORM_2, &rn);
			else { MatResidual(F->smooth->A, b, x, r); VecNorm(r, NORM_2, &rn); }
			pc->mg_ttol = fmax(k->rtol * rn, k->abstol);
		} else pc->mg_ttol = k->abstol;
		int conv = 0; PetscInt i;
		for (i = 0; i < maxit; i++) { mg_cycle(pc, pc->nlevels - 1, &conv); if (conv) break; }
		k->reason = conv ? KSP_CONVERGED_RTOL : KSP_CONVERGED_ITS; k->its = i;
		return;
	}
	if (k->guess_nonzero) { MatMult(k->A, x, r); VecAYPX(r, -1.0, b); } else VecCopy(b, r);       /* r = b - A x */
	k->its = 0;
	PetscReal rnorm = 0.0;
	for (PetscInt i = 0; i < maxit; i++) {
		if (k->normtype == KSP_NORM_UNPRECONDITIONED) {
			VecNorm(r, NORM_2, &rnorm); ksp_monitor(k, i, rnorm); k->rnorm = rnorm; ksp_log(k, rnorm);
			k->reason = ksp_converged(k, i, rnorm); if (k->reason) break;
		}
		PCApply(pc, r, z);                                                                         /* z = B r */
		if (k->normtype == KSP_NORM_PRECONDITIONED) {
			VecNorm(z, NORM_2, &rnorm); ksp_monitor(k, i, rnorm); k->rnorm = rnorm; ksp_log(k, rnorm);
			k->reason = ksp_converged(k, i, rnorm); if (k->reason) break;
		}
		VecAXPY(x, k->scale, z);                                                                   /* x = x + scale z */
		k->its++;
		if (i + 1 < maxit || k->normtype != KSP_NORM_NONE) { MatMult(k->A, x, r); VecAYPX(r, -1.0, b); }
	}
	if (!k->reason) {
		if (k->normtype != KSP_NORM_NONE) {
			if (k->normtype == KSP_NORM_UNPRECONDITIONED) VecNorm(r, NORM_2, &rnorm);
			else { PCApply(pc, r, z); VecNorm(z, NORM_2, &rnorm); }
			k->rnorm = rnorm; ksp_log(k, rnorm); ksp_monitor(k, k->its, rnorm);
		}
		if (k->its >= k->max_it) {
			if (k->normtype != KSP_NORM_NONE) { k->reason = ksp_converged(k, k->its, rnorm); if (!k->reason) k->reason = KSP_DIVERGED_ITS; }
			else k->reason = KSP_CONVERGED_ITS;
		}
	}
}
static void solve_cg(KSP k)                                        /* KSPSolve_CG, left preconditioning */
{
	Vec X = k->sol, B = k->rhs; PC pc = k->pc;
	ksp_work(k, 4, B->n);
	Vec R = k->work[0], Z = k->work[1], P = k->work[2], W = k->work[3];
	PetscScalar a, beta = 0.0, betaold = 1.0, b, dpi = 0.0, dpiold;
	PetscReal dp = 0.0;
	pc_setup(pc);
	k->its = 0;
	if (k->guess_nonzero) { MatMult(k->A, X, R); VecAYPX(R, -1.0, B); } else VecCopy(B, R);
	switch (k->normtype) {
	case KSP_NORM_PRECONDITIONED: PCApply(pc, R, Z); VecNorm(Z, NORM_2, &dp); break;
	case KSP_NORM_UNPRECONDITIONED: VecNorm(R, NORM_2, &dp); break;
	case KSP_NORM_NONE: dp = 0.0; break;
	default: PB_ERR("KSPCG: norm type %d is not offered", (int)k->normtype);
	}
	ksp_log(k, dp); ksp_monitor(k, 0, dp); k->rnorm = dp;
	k->reason = ksp_converged(k, 0, dp);
	if (k->reason) return;
	if (k->normtype != KSP_NORM_PRECONDITIONED) PCApply(pc, R, Z);
	VecDot(Z, R, &beta);
	PetscInt i = 0;
	do {
		k->its = i + 1;
		if (beta == 0.0) { k->reason = KSP_CONVERGED_ATOL; break; }
		else if (i > 0 && beta * betaold < 0.0) { k->reason = KSP_DIVERGED_INDEFINITE_PC; break; }
		if (!i) { VecCopy(Z, P); b = 0.0; } else { b = beta / betaold; VecAYPX(P, b, Z); }      /* p = z + b p */
		dpiold = dpi;
		MatMult(k->A, P, W);
		VecDot(P, W, &dpi);
		betaold = beta;
		if (dpi == 0.0 || (i > 0 && dpi * dpiold <= 0.0)) { k->reason = KSP_DIVERGED_INDEFINITE_MAT; break; }
		a = beta / dpi;
		VecAXPY(X, a, P);
		VecAXPY(R, -a, W);
		if (k->normtype == KSP_NORM_PRECONDITIONED) { PCApply(pc, R, Z); VecNorm(Z, NORM_2, &dp); }
		else if (k->normtype == KSP_NORM_UNPRECONDITIONED) VecNorm(R, NORM_2, &dp);
		else dp = 0.0;
		k->rnorm = dp; ksp_log(k, dp); ksp_monitor(k, i + 1, dp);
		k->reason = ksp_converged(k, i + 1, dp);
		if (k->reason) break;
		if (k->normtype != KSP_NORM_PRECONDITIONED) PCApply(pc, R, Z);
		VecDot(Z, R, &beta);
		i++;
	} while (i < k->max_it);
	if (i >= k->max_it && !k->reason) k->reason = (k->normtype == KSP_NORM_NONE) ? KSP_CONVERGED_ITS : KSP_DIVERGED_ITS;
}
PetscErrorCode KSPSolve(KSP k, Vec b, Vec x)
{
	if (!k->A) PB_ERR("KSPSolve: no operator");
	if (b == x) PB_ERR("KSPSolve: b and x must differ");
	k->rhs = b; k->sol = x;
	k->reason = KSP_CONVERGED_ITERATING; k->its = 0; k->hist_len = 0;
	if (!k->guess_nonzero) VecSet(x, 0.0);
	if (!strcmp(k->type, KSPRICHARDSON)) solve_richardson(k);
	else if (!strcmp(k->type, KSPCG)) solve_cg(k);
	else if (!strcmp(k->type, KSPPREONLY)) {
		if (k->guess_nonzero) PB_ERR("KSPPREONLY with a nonzero initial guess is not allowed");
		PCApply(k->pc, b, x); k->its = 1; k->reason = KSP_CONVERGED_ITS;
	} else if (!strcmp(k->type, "chebyshev"))
		PB_ERR("KSP type 'chebyshev' (PCMG's default level smoother: eigenvalue estimates from GMRES on a random right-hand side) is "
		       "not reproducible and not offered: pass -mg_levels_ksp_type richardson -mg_levels_pc_type {jacobi,sor} "
		       "[-mg_levels_ksp_richardson_scale w] -mg_levels_ksp_max_it nu");
	else PB_ERR("KSP type '%s' is not offered (have: richardson cg preonly)", k->type);
	return 0;
}
PetscErrorCode KSPBuildResidual(KSP k, Vec t, Vec v, Vec *V)      /* KSPBuildResidualDefault: v = b - A x of the last solve */
{
	(void)t;
	if (!k->sol || !k->rhs) PB_ERR("KSPBuildResidual before KSPSolve");
	if (!v) PB_ERR("KSPBuildResidual: a result vector must be supplied");
	MatMult(k->A, k->sol, v); VecAYPX(v, -1.0, k->rhs);
	if (V) *V = v;
	return 0;
}
PetscErrorCode KSPView(KSP k, PetscViewer viewer)
{
	(void)viewer; PC pc = k->pc;
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
	if (k->A) printf("  linear system matrix: rows=%d, cols=%d (in HBM; %lld GPU kernel launches so far)\n", k->A->m, k->A->n, PetscB200LaunchCount());
	return 0;
}
