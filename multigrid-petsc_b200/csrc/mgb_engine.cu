// mgb_engine.cu -- the B200 multigrid engine behind include/mgb200.h.
//
// Host-side orchestration (level hierarchy, HBM allocation, kernel sequencing of the V-cycle and of the
// MG-preconditioned Krylov wrapper, CUDA-graph replay) plus the extern "C" entry points.  The kernels are in
// mgb_stencil.cuh / mgb_transfer.cuh / mgb_blas.cuh / mgb_csr.cuh / mgb_coarse.cuh.  There is NO CPU fallback:
// every entry point fails with MGB_ECUDA when no device is usable.
//
// Reference sequencing that this file reproduces ("ref:" = /root/reference):
//   cycle 0  MultigridVcycle      ref: src/solver.c:1414-1575 (hot loop :1530-1550)
//   cycle 8  MultigridPetscPCMG   ref: src/solver.c:1884-1989, with PETSc's KSPCG / KSPRICHARDSON / PCMG
//            semantics as restated in oracle/minipetsc/minipetsc.c ([PETSc-upstream], unpinned).
#include "../../include/mgb200.h"
#include "mgb_common.cuh"
#include "mgb_stencil.cuh"
#include "mgb_transfer.cuh"
#include "mgb_blas.cuh"
#include "mgb_csr.cuh"
#include "mgb_coarse.cuh"

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
	return code;
}
#define CU(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) \
	return fail(MGB_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); } while (0)
#define TRY(call) do { int _r = (call); if (_r != MGB_OK) return _r; } while (0)

// ------------------------------------------------------------------------------------------------ engine state
struct Csr {
	int m = 0, n = 0; long long nnz = 0;
	int *rowptr = nullptr, *col = nullptr; double *val = nullptr;
};

struct Level {
	int ni = 0, nj = 0, pitch = 0;
	size_t alloc = 0, origin = 0;            // doubles per vector allocation, offset of element (0,0)
	double *base[MGB_NVEC] = {};
	double *v[MGB_NVEC] = {};                // v[k] = base[k] + origin (the two ping-pong buffers may swap)
	std::vector<double> coef_host;           // ni * MGB_COEF_STRIDE
	double *coef = nullptr;
	bool coef_set = false;
	int uniform = 1;
	Csr A, R, P;                             // R, P: this level (fine) <-> level+1
	BandLU lu;                               // coarse LU (cycle 8, coarsest level only)
};

struct mgb_engine {
	mgb_config cfg;
	std::vector<Level> lev;
	Stencil3 R3, P3; bool transfer_set = false;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	double *partial = nullptr;               // partial_cap doubles
	size_t partial_cap = 0;
	double *scal = nullptr;                  // device scalars
	double *scal_host = nullptr;             // pinned mirror
	double *tab_x = nullptr, *tab_y = nullptr; // device tables for separable functions
	double sor_omega = -1.0;                 // omega for which coef[6] (idiag) is valid
	long long launches = 0;
	double last_solve_ms = 0.0;              // CUDA-event time of the last solve loop
	bool csr_built = false;
	// graph replay of the cycle (two graphs: the Jacobi ping-pong state alternates between cycles)
	cudaGraphExec_t gexec[2] = {nullptr, nullptr};
};

static LevelDev ldev(const Level &L)
{
	LevelDev d; d.ni = L.ni; d.nj = L.nj; d.pitch = L.pitch; d.i0 = 0; d.uniform = L.uniform; d.coef = L.coef;
	return d;
}
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
#define RY_STREAM 32
static dim3 stream_grid(const Level &L) { return dim3(cdiv(L.pitch, MGB_SB_COLS), cdiv(L.ni, RY_STREAM)); }
#define LAUNCHED(e) do { (e)->launches++; } while (0)
#define KCHECK() CU(cudaGetLastError())

static int vec_ptr(mgb_engine *e, int which, int level, double **out)
{
	if (level < 0 || level >= (int)e->lev.size()) return fail(MGB_EINVAL, "level %d out of range", level);
	if (which < 0 || which >= MGB_NVEC) return fail(MGB_EINVAL, "vector id %d out of range", which);
	Level &L = e->lev[level];
	if (!L.base[which]) {
		CU(cudaMalloc(&L.base[which], L.alloc * sizeof(double)));
		CU(cudaMemsetAsync(L.base[which], 0, L.alloc * sizeof(double), e->stream));
		L.v[which] = L.base[which] + L.origin;
	}
	*out = L.v[which];
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ lifetime
extern "C" int mgb_version(void) { return 100; }
extern "C" const char *mgb_last_error(void) { return g_err; }

extern "C" int mgb_create(const mgb_config *cfg, mgb_engine **out)
{
	if (!cfg || !out) return fail(MGB_EINVAL, "null argument");
	if (cfg->levels < 1 || cfg->ni < 1 || cfg->nj < 1) return fail(MGB_EINVAL, "levels, ni, nj must be positive");
	if (cfg->nranks > 1) return fail(MGB_EINVAL, "strip decomposition goes through mgb_create_strip (not in this build)");
	int ndev = 0;
	cudaError_t ce = cudaGetDeviceCount(&ndev);
	if (ce != cudaSuccess || ndev < 1)
		return fail(MGB_ECUDA, "no CUDA device available (%s): the B200 engine has no CPU fallback", cudaGetErrorString(ce));
	if (cfg->device >= 0) CU(cudaSetDevice(cfg->device));
	mgb_engine *e = new mgb_engine();
	e->cfg = *cfg;
	e->lev.resize(cfg->levels);
	// level sizes: n_l = (N-1)/2^l - 1 with N-1 = n_0 + 1   (ref: src/matbuild.c:64-66)
	for (int l = 0; l < cfg->levels; ++l) {
		Level &L = e->lev[l];
		L.ni = (cfg->ni + 1) / (1 << l) - 1;
		L.nj = (cfg->nj + 1) / (1 << l) - 1;
		if (L.ni < 1 || L.nj < 1) { delete e; return fail(MGB_EINVAL, "level %d has no interior points", l); }
		L.pitch = ((L.nj + 1 + 15) / 16) * 16;
		L.origin = (size_t)MGB_GHOST_ROWS * L.pitch + 16;
		L.alloc = (size_t)(L.ni + 2 * MGB_GHOST_ROWS + 1) * L.pitch + 32;
		L.coef_host.assign((size_t)L.ni * MGB_COEF_STRIDE, 0.0);
	}
	CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
	CU(cudaEventCreate(&e->ev0)); CU(cudaEventCreate(&e->ev1));
	{
		// partial sums: one per block of the largest streaming grid (fine level) or 3 per block of k_error
		const Level &F = e->lev[0];
		size_t nb = (size_t)cdiv(F.pitch, MGB_SB_COLS) * cdiv(F.ni, RY_STREAM);
		e->partial_cap = nb > 3 * (size_t)MGB_RED_MAXBLOCKS ? nb : 3 * (size_t)MGB_RED_MAXBLOCKS;
		CU(cudaMalloc(&e->partial, sizeof(double) * e->partial_cap));
	}
	CU(cudaMalloc(&e->scal, sizeof(double) * 64));
	CU(cudaMallocHost(&e->scal_host, sizeof(double) * 64));
	for (int l = 0; l < cfg->levels; ++l) {
		Level &L = e->lev[l];
		CU(cudaMalloc(&L.coef, sizeof(double) * L.coef_host.size()));
		double *p;
		for (int k = 0; k <= MGB_VEC_W; ++k) { int r = vec_ptr(e, k, l, &p); if (r) { return r; } }
	}
	CU(cudaMalloc(&e->tab_x, sizeof(double) * (size_t)(cfg->nj + 16)));
	CU(cudaMalloc(&e->tab_y, sizeof(double) * (size_t)(cfg->ni + 16)));
	CU(cudaStreamSynchronize(e->stream));
	*out = e;
	return MGB_OK;
}

static void free_csr(Csr &c) { cudaFree(c.rowptr); cudaFree(c.col); cudaFree(c.val); c = Csr(); }

extern "C" int mgb_destroy(mgb_engine *e)
{
	if (!e) return MGB_OK;
	cudaStreamSynchronize(e->stream);
	for (int g = 0; g < 2; ++g) if (e->gexec[g]) cudaGraphExecDestroy(e->gexec[g]);
	for (auto &L : e->lev) {
		for (int k = 0; k < MGB_NVEC; ++k) cudaFree(L.base[k]);
		cudaFree(L.coef);
		free_csr(L.A); free_csr(L.R); free_csr(L.P);
		bandlu_free(L.lu);
	}
	cudaFree(e->partial); cudaFree(e->scal); cudaFreeHost(e->scal_host);
	cudaFree(e->tab_x); cudaFree(e->tab_y);
	cudaEventDestroy(e->ev0); cudaEventDestroy(e->ev1);
	cudaStreamDestroy(e->stream);
	delete e;
	return MGB_OK;
}

extern "C" int mgb_level_dims(const mgb_engine *e, int level, int *ni, int *nj)
{
	if (!e || level < 0 || level >= (int)e->lev.size()) return fail(MGB_EINVAL, "level out of range");
	if (ni) *ni = e->lev[level].ni;
	if (nj) *nj = e->lev[level].nj;
	return MGB_OK;
}
extern "C" long long mgb_launch_count(const mgb_engine *e) { return e ? e->launches : 0; }
extern "C" double mgb_last_solve_ms(const mgb_engine *e) { return e ? e->last_solve_ms : 0.0; }

// ------------------------------------------------------------------------------------------------ operators
static int upload_coef(mgb_engine *e, Level &L)
{
	CU(cudaMemcpyAsync(L.coef, L.coef_host.data(), sizeof(double) * L.coef_host.size(), cudaMemcpyHostToDevice, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	return MGB_OK;
}

extern "C" int mgb_set_level_operator(mgb_engine *e, int level, const double *row_coeff)
{
	if (!e || !row_coeff) return fail(MGB_EINVAL, "null argument");
	if (level < 0 || level >= (int)e->lev.size()) return fail(MGB_EINVAL, "level %d out of range", level);
	Level &L = e->lev[level];
	L.uniform = 1;
	for (int i = 0; i < L.ni; ++i) {
		double *c = &L.coef_host[(size_t)i * MGB_COEF_STRIDE];
		for (int k = 0; k < 5; ++k) c[k] = row_coeff[i * 5 + k];
		if (c[2] == 0.0) return fail(MGB_EINVAL, "zero diagonal on level %d grid row %d", level, i);
		c[5] = 1.0 / c[2];          // PCSetUp_Jacobi: reciprocal of the diagonal
		c[6] = 1.0 / c[2];          // MatInvertDiagonal_SeqAIJ with omega == 1
		c[7] = c[2];                // mdiag
		if (memcmp(c, &L.coef_host[0], 5 * sizeof(double)) != 0) L.uniform = 0;
	}
	L.coef_set = true;
	bandlu_free(L.lu);
	e->sor_omega = 1.0;
	e->csr_built = false;
	return upload_coef(e, L);
}

// idiag = omega / diag for the SOR kernels (MatInvertDiagonal_SeqAIJ: 1/d when omega == 1 and fshift == 0)
static int set_sor_omega(mgb_engine *e, double omega)
{
	if (e->sor_omega == omega) return MGB_OK;
	for (auto &L : e->lev) {
		for (int i = 0; i < L.ni; ++i) {
			double *c = &L.coef_host[(size_t)i * MGB_COEF_STRIDE];
			c[6] = (omega == 1.0) ? 1.0 / c[2] : omega / (0.0 + c[2]);
		}
		TRY(upload_coef(e, L));
	}
	e->sor_omega = omega;
	return MGB_OK;
}

extern "C" int mgb_set_transfer(mgb_engine *e, const double res3[9], const double pro3[9])
{
	if (!e || !res3 || !pro3) return fail(MGB_EINVAL, "null argument");
	for (int k = 0; k < 9; ++k) {
		if (res3[k] == 0.0 || pro3[k] == 0.0)
			return fail(MGB_EINVAL, "zero transfer weight: the reference drops such entries from res/pro (src/solver.c:1086), not supported");
		e->R3.w[k] = res3[k]; e->P3.w[k] = pro3[k];
	}
	e->transfer_set = true;
	e->csr_built = false;
	return MGB_OK;
}

static int require_ops(mgb_engine *e, bool transfer)
{
	if (!e) return fail(MGB_EINVAL, "null engine");
	for (size_t l = 0; l < e->lev.size(); ++l)
		if (!e->lev[l].coef_set) return fail(MGB_ESTATE, "mgb_set_level_operator was not called for level %d", (int)l);
	if (transfer && e->lev.size() > 1 && !e->transfer_set) return fail(MGB_ESTATE, "mgb_set_transfer was not called");
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ CSR
static int alloc_csr(Csr &c, int m, int n, long long nnz)
{
	free_csr(c);
	if (nnz > 2147483647LL) return fail(MGB_EINVAL, "nnz %lld overflows the 32-bit PetscInt of the reference", nnz);
	c.m = m; c.n = n; c.nnz = nnz;
	CU(cudaMalloc(&c.rowptr, sizeof(int) * ((size_t)m + 1)));
	CU(cudaMalloc(&c.col, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
	CU(cudaMalloc(&c.val, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
	return MGB_OK;
}

static long long host_touch_prefix(int t, int nc)
{
	long long s = 0;
	for (int K = 0; K < nc; ++K) { int d = t - 2 * K; s += d < 0 ? 0 : (d > 3 ? 3 : d); }
	return s;
}

extern "C" int mgb_assemble_csr(mgb_engine *e)
{
	TRY(require_ops(e, true));
	if (e->cfg.red_black_numbering)
		return fail(MGB_EINVAL, "CSR assembly is offered for the reference's natural numbering only (-map 0,1,2), not the -map 3 extension");
	const int Lc = (int)e->lev.size();
	for (int l = 0; l < Lc; ++l) {
		Level &L = e->lev[l];
		const long long N = (long long)L.ni * L.nj;
		if (N > 2147483647LL) return fail(MGB_EINVAL, "level %d has more rows than a 32-bit PetscInt holds", l);
		const long long nnz = 5 * N - 2LL * L.ni - 2LL * L.nj;
		TRY(alloc_csr(L.A, (int)N, (int)N, nnz));
		dim3 g(cdiv(L.nj, 256), L.ni);
		k_csr_A<<<g, 256, 0, e->stream>>>(L.A.rowptr, L.A.col, L.A.val, L.ni, L.nj, L.coef);
		LAUNCHED(e); KCHECK();
		if (l + 1 < Lc) {
			Level &C = e->lev[l + 1];
			const long long NC = (long long)C.ni * C.nj;
			TRY(alloc_csr(L.R, (int)NC, (int)N, 9 * NC));
			dim3 gr(cdiv(C.nj, 256), C.ni);
			k_csr_R<<<gr, 256, 0, e->stream>>>(L.R.rowptr, L.R.col, L.R.val, C.ni, C.nj, L.nj, e->R3);
			LAUNCHED(e); KCHECK();
			const long long pnnz = host_touch_prefix(L.ni, C.ni) * host_touch_prefix(L.nj, C.nj);
			TRY(alloc_csr(L.P, (int)N, (int)NC, pnnz));
			k_csr_P<<<g, 256, 0, e->stream>>>(L.P.rowptr, L.P.col, L.P.val, L.ni, L.nj, C.ni, C.nj, e->P3);
			LAUNCHED(e); KCHECK();
		}
	}
	CU(cudaStreamSynchronize(e->stream));
	e->csr_built = true;
	return MGB_OK;
}

static int pick_csr(const mgb_engine *e, int which, int level, const Csr **out)
{
	if (!e) return fail(MGB_EINVAL, "null engine");
	if (!e->csr_built) return fail(MGB_ESTATE, "mgb_assemble_csr was not called");
	if (level < 0 || level >= (int)e->lev.size()) return fail(MGB_EINVAL, "level %d out of range", level);
	const Level &L = e->lev[level];
	if (which == MGB_MAT_A) *out = &L.A;
	else if (level + 1 >= (int)e->lev.size()) return fail(MGB_EINVAL, "no transfer operator below the coarsest level");
	else if (which == MGB_MAT_RES) *out = &L.R;
	else if (which == MGB_MAT_PRO) *out = &L.P;
	else return fail(MGB_EINVAL, "matrix id %d out of range", which);
	return MGB_OK;
}

extern "C" int mgb_csr_dims(const mgb_engine *e, int which, int level, int *m, int *n, long long *nnz)
{
	const Csr *c; TRY(pick_csr(e, which, level, &c));
	if (m) *m = c->m;
	if (n) *n = c->n;
	if (nnz) *nnz = c->nnz;
	return MGB_OK;
}

extern "C" int mgb_csr_get(const mgb_engine *e, int which, int level, int *rowptr, int *col, double *val)
{
	const Csr *c; TRY(pick_csr(e, which, level, &c));
	if (rowptr) CU(cudaMemcpy(rowptr, c->rowptr, sizeof(int) * ((size_t)c->m + 1), cudaMemcpyDeviceToHost));
	if (col) CU(cudaMemcpy(col, c->col, sizeof(int) * (size_t)c->nnz, cudaMemcpyDeviceToHost));
	if (val) CU(cudaMemcpy(val, c->val, sizeof(double) * (size_t)c->nnz, cudaMemcpyDeviceToHost));
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ vectors
extern "C" int mgb_vec_set(mgb_engine *e, int which, int level, const double *host)
{
	if (!e || !host) return fail(MGB_EINVAL, "null argument");
	double *d; TRY(vec_ptr(e, which, level, &d));
	Level &L = e->lev[level];
	CU(cudaMemcpy2DAsync(d, sizeof(double) * L.pitch, host, sizeof(double) * L.nj, sizeof(double) * L.nj, L.ni,
	                     cudaMemcpyHostToDevice, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	return MGB_OK;
}
extern "C" int mgb_vec_get(mgb_engine *e, int which, int level, double *host)
{
	if (!e || !host) return fail(MGB_EINVAL, "null argument");
	double *d; TRY(vec_ptr(e, which, level, &d));
	Level &L = e->lev[level];
	CU(cudaMemcpy2DAsync(host, sizeof(double) * L.nj, d, sizeof(double) * L.pitch, sizeof(double) * L.nj, L.ni,
	                     cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	return MGB_OK;
}
static int vec_zero(mgb_engine *e, int which, int level)
{
	double *d; TRY(vec_ptr(e, which, level, &d));
	Level &L = e->lev[level];
	const size_t n2 = (size_t)L.ni * L.pitch / 2;
	const int blocks = (int)((n2 + 255) / 256 < 2368 ? (n2 + 255) / 256 : 2368);
	k_axpy<3><<<blocks, 256, 0, e->stream>>>(d, nullptr, n2, 0.0, nullptr, 0.0);
	LAUNCHED(e); KCHECK();
	return MGB_OK;
}
extern "C" int mgb_vec_zero(mgb_engine *e, int which, int level)
{
	if (!e) return fail(MGB_EINVAL, "null engine");
	TRY(vec_zero(e, which, level));
	CU(cudaStreamSynchronize(e->stream));
	return MGB_OK;
}
extern "C" int mgb_set_rhs(mgb_engine *e, const double *b0) { return mgb_vec_set(e, MGB_VEC_B, 0, b0); }
extern "C" int mgb_get_solution(mgb_engine *e, double *u0) { return mgb_vec_get(e, MGB_VEC_U, 0, u0); }

static int upload_tables(mgb_engine *e, const double *tx, const double *ty)
{
	Level &L = e->lev[0];
	CU(cudaMemcpyAsync(e->tab_x, tx, sizeof(double) * L.nj, cudaMemcpyHostToDevice, e->stream));
	CU(cudaMemcpyAsync(e->tab_y, ty, sizeof(double) * L.ni, cudaMemcpyHostToDevice, e->stream));
	return MGB_OK;
}
extern "C" int mgb_set_rhs_separable(mgb_engine *e, const double *gx, const double *gy)
{
	if (!e || !gx || !gy) return fail(MGB_EINVAL, "null argument");
	double *b; TRY(vec_ptr(e, MGB_VEC_B, 0, &b));
	TRY(upload_tables(e, gx, gy));
	Level &L = e->lev[0];
	dim3 g(cdiv(L.pitch, 256), L.ni);
	k_outer<<<g, 256, 0, e->stream>>>(b, e->tab_x, e->tab_y, ldev(L));
	LAUNCHED(e); KCHECK();
	CU(cudaStreamSynchronize(e->stream));
	return MGB_OK;
}
extern "C" int mgb_error_norms_separable(mgb_engine *e, const double *sx, const double *sy, double error[3])
{
	if (!e || !sx || !sy || !error) return fail(MGB_EINVAL, "null argument");
	double *u; TRY(vec_ptr(e, MGB_VEC_U, 0, &u));
	TRY(upload_tables(e, sx, sy));
	Level &L = e->lev[0];
	const size_t total = (size_t)L.ni * L.pitch;
	const int blocks = (int)((total + MGB_RED_THREADS - 1) / MGB_RED_THREADS < 1184 ? (total + MGB_RED_THREADS - 1) / MGB_RED_THREADS : 1184);
	k_error<<<blocks, MGB_RED_THREADS, 0, e->stream>>>(u, e->tab_x, e->tab_y, ldev(L), e->partial);
	LAUNCHED(e); KCHECK();
	k_error2<<<1, 32, 0, e->stream>>>(e->partial, blocks, e->scal + 8);
	LAUNCHED(e); KCHECK();
	CU(cudaMemcpyAsync(e->scal_host + 8, e->scal + 8, 3 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	for (int k = 0; k < 3; ++k) error[k] = e->scal_host[8 + k];
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ launch helpers
// All helpers enqueue on e->stream and do not synchronise.

static int k_apply(mgb_engine *e, int l, const double *x, double *y)
{
	Level &L = e->lev[l];
	k_stream5<ST_APPLY, RY_STREAM><<<stream_grid(L), MGB_SB_THREADS, 0, e->stream>>>(x, nullptr, y, ldev(L), 0.0, nullptr);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}
static int k_residual(mgb_engine *e, int l, const double *x, const double *b, double *r)
{
	Level &L = e->lev[l];
	k_stream5<ST_RESID, RY_STREAM><<<stream_grid(L), MGB_SB_THREADS, 0, e->stream>>>(x, b, r, ldev(L), 0.0, nullptr);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}
// scal[slot] = || b - A x ||_2
static int k_resnorm(mgb_engine *e, int l, const double *x, const double *b, int slot)
{
	Level &L = e->lev[l];
	const dim3 g = stream_grid(L);
	if ((size_t)g.x * g.y > e->partial_cap) return fail(MGB_EINVAL, "grid too large for the partial-sum buffer");
	k_stream5<ST_RESNORM, RY_STREAM><<<g, MGB_SB_THREADS, 0, e->stream>>>(x, b, nullptr, ldev(L), 0.0, e->partial);
	LAUNCHED(e); KCHECK();
	k_reduce2<<<1, 1024, 0, e->stream>>>(e->partial, (int)(g.x * g.y), e->scal, slot, 1);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}
static int k_reduce(mgb_engine *e, int l, const double *x, const double *y, int slot, int take_sqrt)
{
	Level &L = e->lev[l];
	const size_t n2 = (size_t)L.ni * L.pitch / 2;
	size_t want = (n2 + MGB_RED_THREADS * 4 - 1) / (MGB_RED_THREADS * 4);
	const int blocks = (int)(want < 1 ? 1 : (want > MGB_RED_MAXBLOCKS ? MGB_RED_MAXBLOCKS : want));
	if (y) k_reduce1<1><<<blocks, MGB_RED_THREADS, 0, e->stream>>>(x, y, n2, e->partial);
	else   k_reduce1<0><<<blocks, MGB_RED_THREADS, 0, e->stream>>>(x, nullptr, n2, e->partial);
	LAUNCHED(e); KCHECK();
	k_reduce2<<<1, 1024, 0, e->stream>>>(e->partial, blocks, e->scal, slot, take_sqrt);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}
template <int KIND>
static int k_vecop(mgb_engine *e, int l, double *y, const double *x, double alpha)
{
	Level &L = e->lev[l];
	const size_t n2 = (size_t)L.ni * L.pitch / 2;
	size_t want = (n2 + 255) / 256;
	const int blocks = (int)(want < 1 ? 1 : (want > 148 * 32 ? 148 * 32 : want));
	k_axpy<KIND><<<blocks, 256, 0, e->stream>>>(y, x, n2, alpha, nullptr, 0.0);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}
static int read_scalars(mgb_engine *e, int first, int count)
{
	CU(cudaMemcpyAsync(e->scal_host + first, e->scal + first, sizeof(double) * count, cudaMemcpyDeviceToHost, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	return MGB_OK;
}

// one red-black half sweep of colour c (0 = red: (i+j) even), in place
static int k_rb(mgb_engine *e, int l, double *x, const double *b, int colour, double omega, int variant)
{
	Level &L = e->lev[l];
	if (variant == 0) k_rb_half<0, RY_STREAM><<<stream_grid(L), MGB_SB_THREADS, 0, e->stream>>>(x, b, ldev(L), colour, omega);
	else              k_rb_half<1, RY_STREAM><<<stream_grid(L), MGB_SB_THREADS, 0, e->stream>>>(x, b, ldev(L), colour, omega);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}

// The level smoother: KSPSolve(KSPRICHARDSON, KSP_NORM_NONE, max_it = its) on (b, x) -- exactly `its`
// iterations, no convergence test (ref: src/solver.c:1463-1510; semantics in SURVEY.md appendix A).
// xv / sv: vector ids of the iterate and of the Jacobi ping-pong scratch; on return the iterate is in v[xv]
// (the two device pointers are swapped when the sweep count is odd).
static int smooth(mgb_engine *e, int l, const mgb_smoother *s, int its, bool guess_zero, int bv, int xv, int sv)
{
	Level &L = e->lev[l];
	double *b, *x, *w;
	TRY(vec_ptr(e, bv, l, &b)); TRY(vec_ptr(e, xv, l, &x)); TRY(vec_ptr(e, sv, l, &w));
	if (s->type == MGB_SMOOTH_JACOBI) {
		int k = 0;
		if (guess_zero) {
			if (its <= 0) return vec_zero(e, xv, l);
			dim3 g(cdiv(L.pitch, 512), L.ni);
			k_jacobi_first<<<g, 256, 0, e->stream>>>(b, x, ldev(L), s->scale);
			LAUNCHED(e); KCHECK();
			k = 1;
		}
		for (; k < its; ++k) {
			k_stream5<ST_JACOBI, RY_STREAM><<<stream_grid(L), MGB_SB_THREADS, 0, e->stream>>>(x, b, w, ldev(L), s->scale, nullptr);
			LAUNCHED(e); KCHECK();
			double *t = x; x = w; w = t;
		}
		L.v[xv] = x; L.v[sv] = w;
		return MGB_OK;
	}
	if (s->type == MGB_SMOOTH_RBSOR) {
		// KSPSolve_Richardson hands the whole loop to PCApplyRichardson_SOR only when scale == 1:
		// MatSOR(its * pc_its * lits sweeps).  Other scales go through PCApply_SOR per iteration: not offered.
		if (s->scale != 1.0) return fail(MGB_EINVAL, "red-black SOR needs -ksp_richardson_scale 1 (PCApplyRichardson_SOR path)");
		if (guess_zero) TRY(vec_zero(e, xv, l));
		const int total = its * (s->sor_its > 0 ? s->sor_its : 1);
		const double om = s->omega;
		for (int k = 0; k < total; ++k) {
			if (s->sor_sweep == MGB_SOR_SYMMETRIC) {
				// forward: red, black ; backward: black (no-op when omega == 1: x = t * idiag again), red.
				// A red half sweep directly after a red half sweep recomputes the same values when omega == 1.
				const bool prev_red = (k > 0);
				if (!(prev_red && om == 1.0)) TRY(k_rb(e, l, x, b, 0, om, 0));
				TRY(k_rb(e, l, x, b, 1, om, 0));
				if (om != 1.0) TRY(k_rb(e, l, x, b, 1, om, 0));
				TRY(k_rb(e, l, x, b, 0, om, 0));
			} else if (s->sor_sweep == MGB_SOR_FORWARD) {
				TRY(k_rb(e, l, x, b, 0, om, 0));
				TRY(k_rb(e, l, x, b, 1, om, 0));
			} else if (s->sor_sweep == MGB_SOR_BACKWARD) {
				const int variant = (guess_zero && k == 0) ? 0 : 1;
				TRY(k_rb(e, l, x, b, 1, om, variant));
				TRY(k_rb(e, l, x, b, 0, om, variant));
			} else return fail(MGB_EINVAL, "unknown sor_sweep %d", s->sor_sweep);
		}
		return MGB_OK;
	}
	return fail(MGB_EINVAL, "unknown smoother type %d", s->type);
}

static int check_smoother(mgb_engine *e, const mgb_smoother *s)
{
	if (!s) return fail(MGB_EINVAL, "null smoother");
	if (s->type == MGB_SMOOTH_RBSOR) {
		if (s->omega <= 0.0 || s->omega >= 2.0) return fail(MGB_EINVAL, "SOR omega must be in (0,2)");
		TRY(set_sor_omega(e, s->omega));
	} else if (s->type != MGB_SMOOTH_JACOBI) return fail(MGB_EINVAL, "unknown smoother type %d", s->type);
	return MGB_OK;
}

// b[l+1] = res[l] * (b[l] - A[l] x[l])   fused (ref: src/solver.c:1534-1535)
static int restrict_fused(mgb_engine *e, int l, int bv, int xv)
{
	Level &F = e->lev[l], &C = e->lev[l + 1];
	double *b, *x, *bc;
	TRY(vec_ptr(e, bv, l, &b)); TRY(vec_ptr(e, xv, l, &x)); TRY(vec_ptr(e, MGB_VEC_B, l + 1, &bc));
	dim3 blk(32, 4), g(cdiv(C.pitch, 32), cdiv(C.ni, 4));
	k_restrict<1><<<g, blk, 0, e->stream>>>(x, b, nullptr, bc, ldev(F), ldev(C), e->R3);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}
// x[l] += pro[l] * u[l+1]   (ref: src/solver.c:1540-1541 ; PCMG: MatInterpolateAdd)
static int prolong_add(mgb_engine *e, int l, int xv, bool multadd)
{
	Level &F = e->lev[l], &C = e->lev[l + 1];
	double *x, *uc;
	TRY(vec_ptr(e, xv, l, &x)); TRY(vec_ptr(e, MGB_VEC_U, l + 1, &uc));
	dim3 blk(32, 4), g(cdiv(F.pitch, 64), cdiv(F.ni, 4));
	if (multadd) k_prolong_add<1><<<g, blk, 0, e->stream>>>(x, uc, ldev(F), ldev(C), e->P3);
	else         k_prolong_add<0><<<g, blk, 0, e->stream>>>(x, uc, ldev(F), ldev(C), e->P3);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ single ops (C-ABI)
#define NEED(e) do { if (!(e)) return fail(MGB_EINVAL, "null engine"); } while (0)
static int sync(mgb_engine *e) { CU(cudaStreamSynchronize(e->stream)); return MGB_OK; }

extern "C" int mgb_op_apply(mgb_engine *e, int level, int x_vec, int y_vec)
{
	NEED(e); TRY(require_ops(e, false));
	if (x_vec == y_vec) return fail(MGB_EINVAL, "x and y must differ");
	double *x, *y; TRY(vec_ptr(e, x_vec, level, &x)); TRY(vec_ptr(e, y_vec, level, &y));
	TRY(k_apply(e, level, x, y));
	return sync(e);
}
extern "C" int mgb_op_residual(mgb_engine *e, int level)
{
	NEED(e); TRY(require_ops(e, false));
	double *x, *b, *r;
	TRY(vec_ptr(e, MGB_VEC_U, level, &x)); TRY(vec_ptr(e, MGB_VEC_B, level, &b)); TRY(vec_ptr(e, MGB_VEC_R, level, &r));
	TRY(k_residual(e, level, x, b, r));
	return sync(e);
}
extern "C" int mgb_op_residual_norm(mgb_engine *e, int level, double *norm)
{
	NEED(e); TRY(require_ops(e, false));
	double *x, *b; TRY(vec_ptr(e, MGB_VEC_U, level, &x)); TRY(vec_ptr(e, MGB_VEC_B, level, &b));
	TRY(k_resnorm(e, level, x, b, 0));
	TRY(read_scalars(e, 0, 1));
	*norm = e->scal_host[0];
	return MGB_OK;
}
extern "C" int mgb_op_smooth(mgb_engine *e, int level, const mgb_smoother *s, int its, int guess_zero)
{
	NEED(e); TRY(require_ops(e, false)); TRY(check_smoother(e, s));
	if (level < 0 || level >= (int)e->lev.size()) return fail(MGB_EINVAL, "level %d out of range", level);
	TRY(smooth(e, level, s, its, guess_zero != 0, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W));
	return sync(e);
}
extern "C" int mgb_op_restrict(mgb_engine *e, int level, int fused)
{
	NEED(e); TRY(require_ops(e, true));
	if (level < 0 || level + 1 >= (int)e->lev.size()) return fail(MGB_EINVAL, "level %d has no coarser level", level);
	if (fused) { TRY(restrict_fused(e, level, MGB_VEC_B, MGB_VEC_U)); return sync(e); }
	Level &F = e->lev[level], &C = e->lev[level + 1];
	double *r, *bc; TRY(vec_ptr(e, MGB_VEC_R, level, &r)); TRY(vec_ptr(e, MGB_VEC_B, level + 1, &bc));
	dim3 blk(32, 4), g(cdiv(C.pitch, 32), cdiv(C.ni, 4));
	k_restrict<0><<<g, blk, 0, e->stream>>>(nullptr, nullptr, r, bc, ldev(F), ldev(C), e->R3);
	LAUNCHED(e); KCHECK();
	return sync(e);
}
extern "C" int mgb_op_prolong(mgb_engine *e, int level, int multadd)
{
	NEED(e); TRY(require_ops(e, true));
	if (level < 0 || level + 1 >= (int)e->lev.size()) return fail(MGB_EINVAL, "level %d has no coarser level", level);
	TRY(prolong_add(e, level, MGB_VEC_U, multadd != 0));
	return sync(e);
}
extern "C" int mgb_op_norm2(mgb_engine *e, int which, int level, double *out)
{
	NEED(e); double *x; TRY(vec_ptr(e, which, level, &x));
	TRY(k_reduce(e, level, x, nullptr, 0, 1)); TRY(read_scalars(e, 0, 1));
	*out = e->scal_host[0]; return MGB_OK;
}
extern "C" int mgb_op_dot(mgb_engine *e, int xw, int yw, int level, double *out)
{
	NEED(e); double *x, *y; TRY(vec_ptr(e, xw, level, &x)); TRY(vec_ptr(e, yw, level, &y));
	TRY(k_reduce(e, level, x, y, 0, 0)); TRY(read_scalars(e, 0, 1));
	*out = e->scal_host[0]; return MGB_OK;
}
extern "C" int mgb_op_axpy(mgb_engine *e, int yw, double alpha, int xw, int level)
{
	NEED(e); double *x, *y; TRY(vec_ptr(e, xw, level, &x)); TRY(vec_ptr(e, yw, level, &y));
	TRY(k_vecop<0>(e, level, y, x, alpha)); return sync(e);
}
extern "C" int mgb_op_aypx(mgb_engine *e, int yw, double beta, int xw, int level)
{
	NEED(e); double *x, *y; TRY(vec_ptr(e, xw, level, &x)); TRY(vec_ptr(e, yw, level, &y));
	TRY(k_vecop<1>(e, level, y, x, beta)); return sync(e);
}

static int csr_spmv_dev(mgb_engine *e, const Csr *c, const double *x, int xn, int xp, double *y, int yn, int yp)
{
	k_csr_spmv<<<cdiv(c->m, 256), 256, 0, e->stream>>>(c->rowptr, c->col, c->val, c->m, x, xn, xp, y, yn, yp);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}
extern "C" int mgb_csr_spmv_vec(mgb_engine *e, int which, int level, int x_vec, int y_vec)
{
	const Csr *c; TRY(pick_csr(e, which, level, &c));
	const int lx = (which == MGB_MAT_PRO) ? level + 1 : level;
	const int ly = (which == MGB_MAT_RES) ? level + 1 : level;
	if (lx == ly && x_vec == y_vec) return fail(MGB_EINVAL, "x and y must differ");
	double *x, *y; TRY(vec_ptr(e, x_vec, lx, &x)); TRY(vec_ptr(e, y_vec, ly, &y));
	TRY(csr_spmv_dev(e, c, x, e->lev[lx].nj, e->lev[lx].pitch, y, e->lev[ly].nj, e->lev[ly].pitch));
	return sync(e);
}
extern "C" int mgb_csr_spmv(mgb_engine *e, int which, int level, const double *x, double *y)
{
	const Csr *c; TRY(pick_csr(e, which, level, &c));
	if (!x || !y) return fail(MGB_EINVAL, "null argument");
	double *dx, *dy;
	CU(cudaMalloc(&dx, sizeof(double) * (size_t)c->n)); CU(cudaMalloc(&dy, sizeof(double) * (size_t)c->m));
	CU(cudaMemcpyAsync(dx, x, sizeof(double) * (size_t)c->n, cudaMemcpyHostToDevice, e->stream));
	int r = csr_spmv_dev(e, c, dx, c->n, 0, dy, c->m, 0);   // dense vectors: one "row" of length n
	if (r == MGB_OK) {
		cudaError_t ce = cudaMemcpyAsync(y, dy, sizeof(double) * (size_t)c->m, cudaMemcpyDeviceToHost, e->stream);
		if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
		if (ce != cudaSuccess) r = fail(MGB_ECUDA, "csr spmv copy back: %s", cudaGetErrorString(ce));
	}
	cudaFree(dx); cudaFree(dy);
	return r;
}

// ------------------------------------------------------------------------------------------------ cycle 0
// One V-cycle exactly as the body of the reference's while loop (ref: src/solver.c:1531-1546), followed by
// the fine residual norm into scal[0].
static int vcycle_body(mgb_engine *e, const mgb_vcycle_params *p, bool first)
{
	const int Lc = (int)e->lev.size();
	const mgb_smoother *s = &p->smoother;
	TRY(smooth(e, 0, s, p->v0, first, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W));                          // :1531-1532
	for (int l = 1; l < Lc; ++l) {
		TRY(restrict_fused(e, l - 1, MGB_VEC_B, MGB_VEC_U));                                      // :1534-1535
		TRY(smooth(e, l, s, (l == Lc - 1) ? p->v1 : p->v0, true, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W)); // :1536
	}
	for (int l = Lc - 2; l >= 0; --l) {
		TRY(prolong_add(e, l, MGB_VEC_U, false));                                                 // :1540-1541
		TRY(smooth(e, l, s, p->v0, false, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W));                       // :1542
	}
	double *u, *b; TRY(vec_ptr(e, MGB_VEC_U, 0, &u)); TRY(vec_ptr(e, MGB_VEC_B, 0, &b));
	TRY(k_resnorm(e, 0, u, b, 0));                                                                // :1545-1546
	CU(cudaMemcpyAsync(e->scal_host, e->scal, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
	return MGB_OK;
}

static void drop_graphs(mgb_engine *e)
{
	for (int g = 0; g < 2; ++g) if (e->gexec[g]) { cudaGraphExecDestroy(e->gexec[g]); e->gexec[g] = nullptr; }
}

extern "C" int mgb_solve_vcycle(mgb_engine *e, const mgb_vcycle_params *p, double *rnorm, int *num_iter, double *seconds)
{
	NEED(e);
	if (!p || !rnorm || !num_iter) return fail(MGB_EINVAL, "null argument");
	TRY(require_ops(e, true)); TRY(check_smoother(e, &p->smoother));
	if (p->v0 < 0 || p->v1 < 0 || p->max_iter < 0) return fail(MGB_EINVAL, "negative sweep or iteration count");
	double *u, *b; TRY(vec_ptr(e, MGB_VEC_U, 0, &u)); TRY(vec_ptr(e, MGB_VEC_B, 0, &b));
	// bnorm = ||b0|| ; u0 = 0 ; rnorm[0] = ||A0 u0 - b0||                                      (:1512-1520)
	TRY(k_reduce(e, 0, b, nullptr, 1, 1));
	TRY(vec_zero(e, MGB_VEC_U, 0));
	TRY(k_resnorm(e, 0, u, b, 0));
	TRY(read_scalars(e, 0, 2));
	const double bnorm = e->scal_host[1];
	double rn = e->scal_host[0];
	rnorm[0] = rn;
	int iter = 0;
	drop_graphs(e);
	long long launches_per_graph[2] = {0, 0};
	double *state_u[2][64], *state_w[2][64];
	const int Lc = (int)e->lev.size();
	if (Lc > 64) return fail(MGB_EINVAL, "too many levels");
	const auto t0 = std::chrono::steady_clock::now();
	CU(cudaEventRecord(e->ev0, e->stream));
	int gphase = 0;
	while (iter < p->max_iter && 100000000.0 * bnorm > rn && rn > p->rtol * bnorm) {              // :1530
		if (!p->use_graph || iter == 0) {
			TRY(vcycle_body(e, p, iter == 0));
		} else {
			if (!e->gexec[gphase]) {
				// capture this phase: the pointer state before/after is a pure function of the phase
				cudaGraph_t g;
				const long long l0 = e->launches;
				CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
				int r = vcycle_body(e, p, false);
				cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
				if (r != MGB_OK) return r;
				if (ce != cudaSuccess) return fail(MGB_ECUDA, "graph capture failed: %s", cudaGetErrorString(ce));
				CU(cudaGraphInstantiate(&e->gexec[gphase], g, 0));
				cudaGraphDestroy(g);
				launches_per_graph[gphase] = e->launches - l0;
				e->launches = l0;
				for (int l = 0; l < Lc; ++l) { state_u[gphase][l] = e->lev[l].v[MGB_VEC_U]; state_w[gphase][l] = e->lev[l].v[MGB_VEC_W]; }
			} else {
				for (int l = 0; l < Lc; ++l) { e->lev[l].v[MGB_VEC_U] = state_u[gphase][l]; e->lev[l].v[MGB_VEC_W] = state_w[gphase][l]; }
			}
			CU(cudaGraphLaunch(e->gexec[gphase], e->stream));
			e->launches += launches_per_graph[gphase];
			gphase ^= 1;
		}
		CU(cudaStreamSynchronize(e->stream));
		rn = e->scal_host[0];
		iter = iter + 1;
		rnorm[iter] = rn;
	}
	CU(cudaEventRecord(e->ev1, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	const auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	{ float ms = 0.f; CU(cudaEventElapsedTime(&ms, e->ev0, e->ev1)); e->last_solve_ms = ms; }
	drop_graphs(e);
	const double r0 = rnorm[0];
	for (int i = 0; i <= iter; ++i) rnorm[i] = rnorm[i] / r0;                                     // :1554-1557
	*num_iter = iter;
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ cycle 8
// PCApply_MG (multiplicative V, one cycle, x = 0 on entry) on level l with right-hand side bv and iterate xv.
static int pcmg_cycle(mgb_engine *e, const mgb_pcmg_params *p, int l, int bv, int xv)
{
	const int Lc = (int)e->lev.size();
	if (l == Lc - 1) {
		if (p->coarse == MGB_COARSE_RICHARDSON)
			return smooth(e, l, &p->coarse_smoother, p->coarse_its, true, bv, xv, MGB_VEC_W);
		Level &L = e->lev[l];
		double *b, *x; TRY(vec_ptr(e, bv, l, &b)); TRY(vec_ptr(e, xv, l, &x));
		TRY(bandlu_solve(L.lu, b, x, L.ni, L.nj, L.pitch, e->stream));
		LAUNCHED(e);
		return MGB_OK;
	}
	TRY(smooth(e, l, &p->level_smoother, p->level_its, true, bv, xv, MGB_VEC_W));     // pre-smooth from x = 0
	TRY(restrict_fused(e, l, bv, xv));                                                // b_c = R (b - A x)
	TRY(pcmg_cycle(e, p, l + 1, MGB_VEC_B, MGB_VEC_U));                               // x_c = 0 ; recurse
	TRY(prolong_add(e, l, xv, true));                                                 // x = x + P x_c (MatMultAdd)
	TRY(smooth(e, l, &p->level_smoother, p->level_its, false, bv, xv, MGB_VEC_W));    // post-smooth
	return MGB_OK;
}

// KSPConvergedDefault
static int ksp_converged(const mgb_pcmg_params *p, int it, double rn, double *rnorm0, double *ttol)
{
	if (it == 0) { *rnorm0 = rn; *ttol = fmax(p->rtol * rn, p->abstol); }
	if (rn != rn) return -4;                       // KSP_DIVERGED_DTOL (nan)
	if (rn <= *ttol) return (rn < p->abstol) ? 3 : 2;
	if (rn >= p->dtol * (*rnorm0)) return -4;
	return 0;
}

extern "C" int mgb_solve_pcmg(mgb_engine *e, const mgb_pcmg_params *p, double *rnorm, int *num_iter, int *reason_out, double *seconds)
{
	NEED(e);
	if (!p || !rnorm || !num_iter) return fail(MGB_EINVAL, "null argument");
	TRY(require_ops(e, true));
	const int Lc = (int)e->lev.size();
	if (Lc < 2) return fail(MGB_EINVAL, "cycle 8 needs at least two levels");
	TRY(check_smoother(e, &p->level_smoother));
	if (p->coarse == MGB_COARSE_RICHARDSON) {
		TRY(check_smoother(e, &p->coarse_smoother));
		if (p->coarse_smoother.type == MGB_SMOOTH_RBSOR && p->level_smoother.type == MGB_SMOOTH_RBSOR &&
		    p->coarse_smoother.omega != p->level_smoother.omega)
			return fail(MGB_EINVAL, "different SOR omegas on levels and coarse grid are not supported");
	} else if (p->coarse == MGB_COARSE_LU) {
		Level &C = e->lev[Lc - 1];
		TRY(bandlu_factor(C.lu, C.coef_host.data(), C.ni, C.nj, g_err, sizeof g_err));
	} else return fail(MGB_EINVAL, "unknown coarse solver %d", p->coarse);
	if (p->outer != MGB_KSP_CG && p->outer != MGB_KSP_RICHARDSON) return fail(MGB_EINVAL, "unknown outer KSP %d", p->outer);

	double *X, *B, *R, *Z, *P, *Q;
	TRY(vec_ptr(e, MGB_VEC_U, 0, &X)); TRY(vec_ptr(e, MGB_VEC_B, 0, &B)); TRY(vec_ptr(e, MGB_VEC_R, 0, &R));
	TRY(vec_ptr(e, MGB_VEC_Z, 0, &Z)); TRY(vec_ptr(e, MGB_VEC_P, 0, &P)); TRY(vec_ptr(e, MGB_VEC_Q, 0, &Q));
	CU(cudaStreamSynchronize(e->stream));
	const auto t0 = std::chrono::steady_clock::now();
	CU(cudaEventRecord(e->ev0, e->stream));
	int reason = 0, its = 0, nlog = 0;
	double rnorm0 = 0.0, ttol = 0.0, dp = 0.0;
	auto logr = [&](double v) { if (nlog < p->max_iter) rnorm[nlog++] = v; };   // KSPSetResidualHistory(na = numIter)
	for (int i = 0; i <= p->max_iter; ++i) rnorm[i] = NAN;

	TRY(vec_zero(e, MGB_VEC_U, 0));                          // KSPSolve: zero initial guess
	TRY(k_vecop<2>(e, 0, R, B, 0.0));                        // r = b
	if (p->outer == MGB_KSP_CG) {
		// KSPSolve_CG, KSP_NORM_UNPRECONDITIONED (ref: src/solver.c:1922)
		double beta = 0.0, betaold = 1.0, dpi = 0.0, dpiold;
		TRY(k_reduce(e, 0, R, nullptr, 0, 1)); TRY(read_scalars(e, 0, 1)); dp = e->scal_host[0];
		logr(dp);
		reason = ksp_converged(p, 0, dp, &rnorm0, &ttol);
		if (!reason) {
			TRY(pcmg_cycle(e, p, 0, MGB_VEC_R, MGB_VEC_Z)); TRY(vec_ptr(e, MGB_VEC_Z, 0, &Z));     // z = B r
			TRY(k_reduce(e, 0, Z, R, 0, 0)); TRY(read_scalars(e, 0, 1)); beta = e->scal_host[0];   // beta = z'r
			int i = 0;
			do {
				its = i + 1;
				if (beta == 0.0) { reason = 3; break; }
				else if (i > 0 && beta * betaold < 0.0) { reason = -8; break; }                    // KSP_DIVERGED_INDEFINITE_PC
				if (i == 0) { TRY(k_vecop<2>(e, 0, P, Z, 0.0)); }                                  // p = z
				else { TRY(k_vecop<1>(e, 0, P, Z, beta / betaold)); }                              // p = z + b p
				dpiold = dpi;
				TRY(k_apply(e, 0, P, Q));                                                          // w = A p
				TRY(k_reduce(e, 0, P, Q, 0, 0)); TRY(read_scalars(e, 0, 1)); dpi = e->scal_host[0];
				betaold = beta;
				if (dpi == 0.0 || (i > 0 && dpi * dpiold <= 0.0)) { reason = -10; break; }         // KSP_DIVERGED_INDEFINITE_MAT
				const double a = beta / dpi;
				TRY(k_vecop<0>(e, 0, X, P, a));                                                    // x = x + a p
				TRY(k_vecop<0>(e, 0, R, Q, -a));                                                   // r = r - a w
				TRY(k_reduce(e, 0, R, nullptr, 0, 1)); TRY(read_scalars(e, 0, 1)); dp = e->scal_host[0];
				logr(dp);
				reason = ksp_converged(p, i + 1, dp, &rnorm0, &ttol);
				if (reason) break;
				TRY(pcmg_cycle(e, p, 0, MGB_VEC_R, MGB_VEC_Z)); TRY(vec_ptr(e, MGB_VEC_Z, 0, &Z));
				TRY(k_reduce(e, 0, Z, R, 0, 0)); TRY(read_scalars(e, 0, 1)); beta = e->scal_host[0];
				i++;
			} while (i < p->max_iter);
			if (i >= p->max_iter && !reason) reason = -3;                                          // KSP_DIVERGED_ITS
		}
	} else {
		// KSPSolve_Richardson, general path (residual norm logged every iteration), scale 1
		for (int i = 0; i < p->max_iter; ++i) {
			TRY(k_reduce(e, 0, R, nullptr, 0, 1)); TRY(read_scalars(e, 0, 1)); dp = e->scal_host[0];
			logr(dp);
			reason = ksp_converged(p, i, dp, &rnorm0, &ttol);
			if (reason) break;
			TRY(pcmg_cycle(e, p, 0, MGB_VEC_R, MGB_VEC_Z)); TRY(vec_ptr(e, MGB_VEC_Z, 0, &Z));     // z = B r
			TRY(k_vecop<0>(e, 0, X, Z, 1.0));                                                      // x = x + scale z
			its++;
			TRY(k_residual(e, 0, X, B, R));                                                        // r = b - A x
		}
		if (!reason) {
			TRY(k_reduce(e, 0, R, nullptr, 0, 1)); TRY(read_scalars(e, 0, 1)); dp = e->scal_host[0];
			logr(dp);
			if (its >= p->max_iter) { reason = ksp_converged(p, its, dp, &rnorm0, &ttol); if (!reason) reason = -3; }
		}
	}
	CU(cudaEventRecord(e->ev1, e->stream));
	CU(cudaStreamSynchronize(e->stream));
	const auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	{ float ms = 0.f; CU(cudaEventElapsedTime(&ms, e->ev0, e->ev1)); e->last_solve_ms = ms; }
	// ref: src/solver.c:1971-1976 -- numIter = KSPGetIterationNumber ; rnorm[i] /= rnorm[0]
	const double r0 = rnorm[0];
	for (int i = 0; i < its + 1 && i <= p->max_iter; ++i) rnorm[i] = rnorm[i] / r0;
	*num_iter = its;
	if (reason_out) *reason_out = reason;
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ measurement
extern "C" int mgb_time_op(mgb_engine *e, int op, int level, int reps, double *ms_per_launch)
{
	NEED(e); TRY(require_ops(e, true));
	if (level < 0 || level >= (int)e->lev.size()) return fail(MGB_EINVAL, "level %d out of range", level);
	if (reps < 1 || !ms_per_launch) return fail(MGB_EINVAL, "bad reps / null output");
	double *u, *b, *r, *w;
	TRY(vec_ptr(e, MGB_VEC_U, level, &u)); TRY(vec_ptr(e, MGB_VEC_B, level, &b));
	TRY(vec_ptr(e, MGB_VEC_R, level, &r)); TRY(vec_ptr(e, MGB_VEC_W, level, &w));
	mgb_smoother jac = {MGB_SMOOTH_JACOBI, 0.8, 1.0, 0, 1};
	const bool has_coarse = level + 1 < (int)e->lev.size();
	if ((op == 4 || op == 5) && !has_coarse) return fail(MGB_EINVAL, "level %d has no coarser level", level);
	if (op == 7 && !e->csr_built) return fail(MGB_ESTATE, "mgb_assemble_csr was not called");
	TRY(set_sor_omega(e, 1.0));
	for (int pass = 0; pass < 2; ++pass) {
		const int n = pass == 0 ? 2 : reps;                   // pass 0: warm-up
		if (pass == 1) CU(cudaEventRecord(e->ev0, e->stream));
		for (int k = 0; k < n; ++k) {
			switch (op) {
			case 0: TRY(k_apply(e, level, u, r)); break;
			case 1: TRY(k_residual(e, level, u, b, r)); break;
			case 2: TRY(smooth(e, level, &jac, 1, false, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W)); break;
			case 3: TRY(k_rb(e, level, u, b, 0, 1.0, 0)); TRY(k_rb(e, level, u, b, 1, 1.0, 0)); break;
			case 4: TRY(restrict_fused(e, level, MGB_VEC_B, MGB_VEC_U)); break;
			case 5: TRY(prolong_add(e, level, MGB_VEC_U, false)); break;
			case 6: TRY(k_resnorm(e, level, u, b, 0)); break;
			case 7: TRY(csr_spmv_dev(e, &e->lev[level].A, u, e->lev[level].nj, e->lev[level].pitch, r, e->lev[level].nj, e->lev[level].pitch)); break;
			case 8: TRY(k_reduce(e, level, u, nullptr, 0, 1)); break;
			case 9: TRY(k_reduce(e, level, u, b, 0, 0)); break;
			case 10: TRY(k_vecop<0>(e, level, r, u, 0.5)); break;
			default: return fail(MGB_EINVAL, "unknown op %d", op);
			}
		}
		if (pass == 1) CU(cudaEventRecord(e->ev1, e->stream));
		CU(cudaStreamSynchronize(e->stream));
	}
	float ms = 0.f;
	CU(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
	*ms_per_launch = (double)ms / reps;
	return MGB_OK;
}
