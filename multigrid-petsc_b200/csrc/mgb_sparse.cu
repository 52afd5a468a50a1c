// mgb_sparse.cu -- general CSR matrices and dense vectors in HBM behind include/mgb200_sparse.h.
//
// The reference's unmodified src/solver.c drives PETSc objects; the PETSc-subset layer in host/petsc_b200/ maps each of
// them onto one object of this file and each PETSc operation onto one entry point.  Nothing here is specialised to the
// 5-point stencil (that is mgb_engine.cu's job): these kernels take whatever matrix the reference assembled -- several
// grids per level, the research cycles' combined operators -- and reproduce PETSc's SeqAIJ arithmetic on it.
//
// Arithmetic: compiled with --fmad=false; products and sums are written with __dmul_rn / __dadd_rn / __dsub_rn in the order
// of MatMult_SeqAIJ, MatSOR_SeqAIJ, MatLUFactorNumeric (ILU(0), inverted pivots), as restated in oracle/minipetsc
// [PETSc-upstream].  Sequential sweeps (SOR, ILU(0), triangular solves) run level set by level set: a row starts when every
// row it depends on is complete, so it reads exactly the operands the sequential loop would have read.
#include "../../include/mgb200.h"
#include "../../include/mgb200_sparse.h"
#include "mgb_common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

extern "C" void mgb__set_error(const char *msg);     // mgb_engine.cu: the text behind mgb_last_error()

static int sfail(int code, const char *fmt, ...)
{
	char buf[512];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	mgb__set_error(buf);
	return code;
}
#define SCU(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) \
	return sfail(MGB_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); } while (0)
#define STRY(call) do { int _r = (call); if (_r != MGB_OK) return _r; } while (0)

static long long g_sparse_launches = 0;
#define SLAUNCHED() do { ++g_sparse_launches; } while (0)
#define SKCHECK() SCU(cudaGetLastError())
extern "C" long long mgb_sparse_launch_count(void) { return g_sparse_launches; }

#define SP_RED_BLK 4096          // elements per reduction block (the order the CPU checker uses, so tests can demand equal bits)
#define SP_SWEEP_THREADS 1024

struct mgb_dvec { int n; double *d; };
struct mgb_dindex { int n; int *d; };
struct Levels {                                          // level sets of one triangular part
	int nlev = 0; int *ptr = nullptr, *rows = nullptr;   // device: ptr[nlev+1], rows[m] (rows of level l: rows[ptr[l] .. ptr[l+1]))
};
struct mgb_dcsr {
	int m, n; long long nnz;
	int *ia, *ja, *diag; double *va;                     // device
	std::vector<int> h_ia, h_ja, h_diag;                 // host copy of the pattern (level sets, checks)
	bool sched = false; Levels fwd, bwd;
	double *idiag = nullptr, *mdiag = nullptr, *ssor_t = nullptr; double sor_omega = 0.0, sor_fshift = 0.0; bool idiag_valid = false;
	double *fac = nullptr, *tmp = nullptr; bool ilu_valid = false;      // ILU(0) factors on the pattern of A
	double *dense = nullptr; bool lu_valid = false;                      // dense LU (small coarse problems)
	int *status = nullptr;                                               // device word: != 0 after a zero pivot / missing entry
};

static int need_device()
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return sfail(MGB_ECUDA, "no CUDA device: the B200 engine has no CPU fallback");
	return MGB_OK;
}

// ---------------------------------------------------------------------------------------------------------------- vectors
enum { EW_SET = 0, EW_AXPY, EW_AYPX, EW_WAXPY, EW_AXPBYPCZ, EW_SCALE };
template <int OP>
__global__ void __launch_bounds__(256)
k_sp_ew(double *__restrict__ out, const double *__restrict__ x, const double *__restrict__ y, int n, double a, double b, double c)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		double r;
		if (OP == EW_SET) r = a;
		else if (OP == EW_AXPY) r = add(out[i], mul(a, x[i]));                     // y = y + alpha x
		else if (OP == EW_AYPX) r = add(x[i], mul(a, out[i]));                     // y = x + beta y
		else if (OP == EW_WAXPY) r = add(mul(a, x[i]), y[i]);                      // w = alpha x + y
		else if (OP == EW_AXPBYPCZ) r = add(add(mul(c, out[i]), mul(a, x[i])), mul(b, y[i]));   // z = gamma z + alpha x + beta y
		else r = mul(a, out[i]);
		out[i] = r;
	}
}
__global__ void __launch_bounds__(256)
k_sp_pmult(double *w, const double *x, const double *y, int n)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) w[i] = mul(x[i], y[i]);
}
static int ew_grid(int n) { int g = (n + 255) / 256; return g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g); }

extern "C" int mgb_dvec_create(int n, mgb_dvec **out)
{
	if (!out || n < 0) return sfail(MGB_EINVAL, "mgb_dvec_create: bad arguments");
	STRY(need_device());
	mgb_dvec *v = new mgb_dvec; v->n = n; v->d = nullptr;
	cudaError_t e = cudaMalloc(&v->d, sizeof(double) * (size_t)(n > 0 ? n : 1));
	if (e == cudaSuccess) e = cudaMemset(v->d, 0, sizeof(double) * (size_t)(n > 0 ? n : 1));
	if (e != cudaSuccess) { cudaFree(v->d); delete v; return sfail(MGB_ECUDA, "mgb_dvec_create(%d): %s", n, cudaGetErrorString(e)); }
	*out = v; return MGB_OK;
}
extern "C" int mgb_dvec_destroy(mgb_dvec *v) { if (v) { cudaFree(v->d); delete v; } return MGB_OK; }
extern "C" int mgb_dvec_size(const mgb_dvec *v) { return v ? v->n : 0; }
extern "C" int mgb_dvec_upload(mgb_dvec *v, const double *h)
{
	if (!v || !h) return sfail(MGB_EINVAL, "mgb_dvec_upload: null argument");
	SCU(cudaMemcpy(v->d, h, sizeof(double) * (size_t)v->n, cudaMemcpyHostToDevice)); return MGB_OK;
}
extern "C" int mgb_dvec_download(const mgb_dvec *v, double *h)
{
	if (!v || !h) return sfail(MGB_EINVAL, "mgb_dvec_download: null argument");
	SCU(cudaMemcpy(h, v->d, sizeof(double) * (size_t)v->n, cudaMemcpyDeviceToHost)); return MGB_OK;
}
#define SAME_N(a, b, what) do { if (!(a) || !(b) || (a)->n != (b)->n) return sfail(MGB_EINVAL, what ": null or size mismatch"); } while (0)
extern "C" int mgb_dvec_set(mgb_dvec *v, double alpha)
{
	if (!v) return sfail(MGB_EINVAL, "mgb_dvec_set: null");
	k_sp_ew<EW_SET><<<ew_grid(v->n), 256>>>(v->d, nullptr, nullptr, v->n, alpha, 0, 0); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dvec_copy(mgb_dvec *dst, const mgb_dvec *src)
{
	SAME_N(dst, src, "mgb_dvec_copy");
	if (dst != src) SCU(cudaMemcpyAsync(dst->d, src->d, sizeof(double) * (size_t)src->n, cudaMemcpyDeviceToDevice, 0));
	return MGB_OK;
}
extern "C" int mgb_dvec_axpy(mgb_dvec *y, double alpha, const mgb_dvec *x)
{
	SAME_N(y, x, "mgb_dvec_axpy");
	k_sp_ew<EW_AXPY><<<ew_grid(y->n), 256>>>(y->d, x->d, nullptr, y->n, alpha, 0, 0); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dvec_aypx(mgb_dvec *y, double beta, const mgb_dvec *x)
{
	SAME_N(y, x, "mgb_dvec_aypx");
	k_sp_ew<EW_AYPX><<<ew_grid(y->n), 256>>>(y->d, x->d, nullptr, y->n, beta, 0, 0); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dvec_waxpy(mgb_dvec *w, double alpha, const mgb_dvec *x, const mgb_dvec *y)
{
	SAME_N(w, x, "mgb_dvec_waxpy"); SAME_N(w, y, "mgb_dvec_waxpy");
	if (w == x || w == y) return sfail(MGB_EINVAL, "mgb_dvec_waxpy: w must differ from x and y");
	k_sp_ew<EW_WAXPY><<<ew_grid(w->n), 256>>>(w->d, x->d, y->d, w->n, alpha, 0, 0); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dvec_axpbypcz(mgb_dvec *z, double alpha, double beta, double gamma, const mgb_dvec *x, const mgb_dvec *y)
{
	SAME_N(z, x, "mgb_dvec_axpbypcz"); SAME_N(z, y, "mgb_dvec_axpbypcz");
	if (z == x || z == y) return sfail(MGB_EINVAL, "mgb_dvec_axpbypcz: z must differ from x and y");
	k_sp_ew<EW_AXPBYPCZ><<<ew_grid(z->n), 256>>>(z->d, x->d, y->d, z->n, alpha, beta, gamma); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dvec_pointwise_mult(mgb_dvec *w, const mgb_dvec *x, const mgb_dvec *y)
{
	SAME_N(w, x, "mgb_dvec_pointwise_mult"); SAME_N(w, y, "mgb_dvec_pointwise_mult");
	// element-wise, each element read before it is written by the same thread: w may alias x or y (no __restrict__ here)
	k_sp_pmult<<<ew_grid(w->n), 256>>>(w->d, x->d, y->d, w->n); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dvec_scale(mgb_dvec *x, double alpha)
{
	if (!x) return sfail(MGB_EINVAL, "mgb_dvec_scale: null");
	k_sp_ew<EW_SCALE><<<ew_grid(x->n), 256>>>(x->d, nullptr, nullptr, x->n, alpha, 0, 0); SLAUNCHED(); SKCHECK(); return MGB_OK;
}

// blocked reduction: thread b sums block b left to right; one thread then sums the block sums left to right.
// MODE 0: sum x*y   1: sum |x|   2: max |x|
template <int MODE>
__global__ void __launch_bounds__(128)
k_sp_red1(const double *__restrict__ x, const double *__restrict__ y, int n, double *__restrict__ part)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	const int lo = b * SP_RED_BLK;
	if (lo >= n) return;
	const int hi = min(lo + SP_RED_BLK, n);
	double s = 0.0;
	for (int i = lo; i < hi; ++i) {
		if (MODE == 0) s = add(s, mul(x[i], y[i]));
		else if (MODE == 1) s = add(s, fabs(x[i]));
		else s = fmax(s, fabs(x[i]));
	}
	part[b] = s;
}
template <int MODE>
__global__ void k_sp_red2(const double *__restrict__ part, int nb, double *__restrict__ out)
{
	double s = 0.0;
	for (int b = 0; b < nb; ++b) s = (MODE == 2) ? fmax(s, part[b]) : add(s, part[b]);
	out[0] = s;
}
// scratch of the reductions (block sums + the result), grown on demand and kept: a VecNorm per cycle must not cost a cudaMalloc
static double *g_red_part = nullptr; static int g_red_cap = 0;
static std::mutex g_red_mu;                              // the scratch is shared by every vector of the process
static int reduce(int mode, const mgb_dvec *x, const mgb_dvec *y, double *out)
{
	std::lock_guard<std::mutex> lk(g_red_mu);
	const int n = x->n, nb = (n + SP_RED_BLK - 1) / SP_RED_BLK;
	if (n == 0) { *out = 0.0; return MGB_OK; }
	if (nb + 1 > g_red_cap) {
		cudaFree(g_red_part); g_red_part = nullptr; g_red_cap = 0;
		SCU(cudaMalloc(&g_red_part, sizeof(double) * (size_t)(2 * nb + 2)));
		g_red_cap = 2 * nb + 2;
	}
	double *part = g_red_part;
	const int gr = (nb + 127) / 128;
	if (mode == 0) { k_sp_red1<0><<<gr, 128>>>(x->d, y->d, n, part); k_sp_red2<0><<<1, 1>>>(part, nb, part + nb); }
	else if (mode == 1) { k_sp_red1<1><<<gr, 128>>>(x->d, nullptr, n, part); k_sp_red2<1><<<1, 1>>>(part, nb, part + nb); }
	else { k_sp_red1<2><<<gr, 128>>>(x->d, nullptr, n, part); k_sp_red2<2><<<1, 1>>>(part, nb, part + nb); }
	SLAUNCHED(); SLAUNCHED();
	SKCHECK();
	SCU(cudaMemcpy(out, part + nb, sizeof(double), cudaMemcpyDeviceToHost));
	return MGB_OK;
}
extern "C" int mgb_dvec_dot(const mgb_dvec *x, const mgb_dvec *y, double *out)
{
	SAME_N(x, y, "mgb_dvec_dot");
	if (!out) return sfail(MGB_EINVAL, "mgb_dvec_dot: null output");
	return reduce(0, x, y, out);
}
extern "C" int mgb_dvec_norm(const mgb_dvec *x, int type, double *out)
{
	if (!x || !out) return sfail(MGB_EINVAL, "mgb_dvec_norm: null argument");
	if (type == MGB_NORM_2 || type == 2) { double s; STRY(reduce(0, x, x, &s)); *out = sqrt(s); return MGB_OK; }
	if (type == MGB_NORM_1) return reduce(1, x, nullptr, out);
	if (type == MGB_NORM_INF) return reduce(2, x, nullptr, out);
	return sfail(MGB_EINVAL, "mgb_dvec_norm: unknown norm type %d", type);
}

// ---------------------------------------------------------------------------------------------------------------- index lists
__global__ void __launch_bounds__(256)
k_sp_gather(double *__restrict__ sub, const double *__restrict__ x, const int *__restrict__ idx, int n, int scatter, double *__restrict__ xw)
{
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		if (scatter) xw[idx[k]] = sub[k]; else sub[k] = x[idx[k]];
	}
}
extern "C" int mgb_dindex_create(int n, const int *idx, mgb_dindex **out)
{
	if (!out || n < 0 || (n > 0 && !idx)) return sfail(MGB_EINVAL, "mgb_dindex_create: bad arguments");
	STRY(need_device());
	mgb_dindex *s = new mgb_dindex; s->n = n; s->d = nullptr;
	cudaError_t e = cudaMalloc(&s->d, sizeof(int) * (size_t)(n > 0 ? n : 1));
	if (e == cudaSuccess && n > 0) e = cudaMemcpy(s->d, idx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { cudaFree(s->d); delete s; return sfail(MGB_ECUDA, "mgb_dindex_create: %s", cudaGetErrorString(e)); }
	*out = s; return MGB_OK;
}
extern "C" int mgb_dindex_destroy(mgb_dindex *s) { if (s) { cudaFree(s->d); delete s; } return MGB_OK; }
extern "C" int mgb_dvec_gather(mgb_dvec *sub, const mgb_dvec *x, const mgb_dindex *is)
{
	if (!sub || !x || !is || sub->n != is->n) return sfail(MGB_EINVAL, "mgb_dvec_gather: null or size mismatch");
	k_sp_gather<<<ew_grid(is->n), 256>>>(sub->d, x->d, is->d, is->n, 0, nullptr); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dvec_scatter(mgb_dvec *x, const mgb_dvec *sub, const mgb_dindex *is)
{
	if (!sub || !x || !is || sub->n != is->n) return sfail(MGB_EINVAL, "mgb_dvec_scatter: null or size mismatch");
	k_sp_gather<<<ew_grid(is->n), 256>>>(sub->d, nullptr, is->d, is->n, 1, x->d); SLAUNCHED(); SKCHECK(); return MGB_OK;
}

// ---------------------------------------------------------------------------------------------------------------- CSR
extern "C" int mgb_dcsr_create(int m, int n, const int *ia, const int *ja, const double *va, mgb_dcsr **out)
{
	if (!out || m < 0 || n < 0 || !ia || (ia[m] > 0 && (!ja || !va))) return sfail(MGB_EINVAL, "mgb_dcsr_create: bad arguments");
	STRY(need_device());
	const long long nnz = ia[m];
	for (int i = 0; i < m; ++i) {
		if (ia[i + 1] < ia[i]) return sfail(MGB_EINVAL, "mgb_dcsr_create: row pointers decrease at row %d", i);
		for (int k = ia[i]; k < ia[i + 1]; ++k) {
			if (ja[k] < 0 || ja[k] >= n) return sfail(MGB_EINVAL, "mgb_dcsr_create: column %d out of range in row %d", ja[k], i);
			if (k > ia[i] && ja[k] <= ja[k - 1]) return sfail(MGB_EINVAL, "mgb_dcsr_create: columns of row %d are not ascending", i);
		}
	}
	mgb_dcsr *A = new mgb_dcsr; A->m = m; A->n = n; A->nnz = nnz; A->ia = A->ja = A->diag = nullptr; A->va = nullptr;
	A->h_ia.assign(ia, ia + m + 1); A->h_ja.assign(ja, ja + nnz); A->h_diag.assign((size_t)(m > 0 ? m : 1), -1);
	for (int i = 0; i < m; ++i) for (int k = ia[i]; k < ia[i + 1]; ++k) if (ja[k] == i) A->h_diag[i] = k;
	cudaError_t e = cudaMalloc(&A->ia, sizeof(int) * (size_t)(m + 1));
	if (e == cudaSuccess) e = cudaMalloc(&A->ja, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
	if (e == cudaSuccess) e = cudaMalloc(&A->va, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
	if (e == cudaSuccess) e = cudaMalloc(&A->diag, sizeof(int) * (size_t)(m > 0 ? m : 1));
	if (e == cudaSuccess) e = cudaMalloc(&A->status, sizeof(int));
	if (e == cudaSuccess) e = cudaMemset(A->status, 0, sizeof(int));
	if (e == cudaSuccess) e = cudaMemcpy(A->ia, ia, sizeof(int) * (size_t)(m + 1), cudaMemcpyHostToDevice);
	if (e == cudaSuccess && nnz > 0) e = cudaMemcpy(A->ja, ja, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice);
	if (e == cudaSuccess && nnz > 0) e = cudaMemcpy(A->va, va, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice);
	if (e == cudaSuccess && m > 0) e = cudaMemcpy(A->diag, A->h_diag.data(), sizeof(int) * (size_t)m, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { mgb_dcsr_destroy(A); return sfail(MGB_ECUDA, "mgb_dcsr_create(%d x %d, %lld nnz): %s", m, n, nnz, cudaGetErrorString(e)); }
	*out = A; return MGB_OK;
}
static void free_levels(Levels &L) { cudaFree(L.ptr); cudaFree(L.rows); L = Levels(); }
extern "C" int mgb_dcsr_destroy(mgb_dcsr *A)
{
	if (!A) return MGB_OK;
	cudaFree(A->ia); cudaFree(A->ja); cudaFree(A->va); cudaFree(A->diag); cudaFree(A->status);
	cudaFree(A->idiag); cudaFree(A->mdiag); cudaFree(A->ssor_t); cudaFree(A->fac); cudaFree(A->tmp); cudaFree(A->dense);
	free_levels(A->fwd); free_levels(A->bwd);
	delete A; return MGB_OK;
}

// y_i = [y_i +] sum_k a_ik x_k in ascending column order (MatMult_SeqAIJ / MatMultAdd_SeqAIJ); one thread per row
template <int ADD>
__global__ void __launch_bounds__(256)
k_sp_mult(const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ va, int m,
          const double *__restrict__ x, const double *y, double *z)      // z may be y (each thread reads y_i before it writes z_i)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	double sum = ADD ? y[i] : 0.0;
	for (int k = ia[i]; k < ia[i + 1]; ++k) sum = add(sum, mul(va[k], x[ja[k]]));
	z[i] = sum;
}
extern "C" int mgb_dcsr_mult(const mgb_dcsr *A, const mgb_dvec *x, mgb_dvec *y)
{
	if (!A || !x || !y || x->n != A->n || y->n != A->m) return sfail(MGB_EINVAL, "mgb_dcsr_mult: null or size mismatch");
	if (x == y) return sfail(MGB_EINVAL, "mgb_dcsr_mult: x and y must differ");
	if (A->m == 0) return MGB_OK;
	k_sp_mult<0><<<(A->m + 255) / 256, 256>>>(A->ia, A->ja, A->va, A->m, x->d, nullptr, y->d); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dcsr_mult_add(const mgb_dcsr *A, const mgb_dvec *x, const mgb_dvec *y, mgb_dvec *z)
{
	if (!A || !x || !y || !z || x->n != A->n || y->n != A->m || z->n != A->m) return sfail(MGB_EINVAL, "mgb_dcsr_mult_add: null or size mismatch");
	if (x == z) return sfail(MGB_EINVAL, "mgb_dcsr_mult_add: x and z must differ");
	if (A->m == 0) return MGB_OK;
	// z may be y (MatInterpolateAdd(P, xc, x, x)): each thread reads y_i before it writes z_i, nobody else touches element i
	k_sp_mult<1><<<(A->m + 255) / 256, 256>>>(A->ia, A->ja, A->va, A->m, x->d, y->d, z->d); SLAUNCHED(); SKCHECK(); return MGB_OK;
}
extern "C" int mgb_dcsr_scale(mgb_dcsr *A, double s)
{
	if (!A) return sfail(MGB_EINVAL, "mgb_dcsr_scale: null");
	if (A->nnz > 0) { k_sp_ew<EW_SCALE><<<ew_grid((int)A->nnz), 256>>>(A->va, nullptr, nullptr, (int)A->nnz, s, 0, 0); SLAUNCHED(); SKCHECK(); }
	A->idiag_valid = false; A->ilu_valid = false; A->lu_valid = false;
	return MGB_OK;
}
__global__ void __launch_bounds__(256)
k_sp_invdiag(const int *__restrict__ diag, const double *__restrict__ va, int m, double *__restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const double d = diag[i] >= 0 ? va[diag[i]] : 0.0;
	out[i] = (d != 0.0) ? __ddiv_rn(1.0, d) : 1.0;
}
extern "C" int mgb_dcsr_inverse_diagonal(const mgb_dcsr *A, mgb_dvec *dinv)
{
	if (!A || !dinv || dinv->n != A->m) return sfail(MGB_EINVAL, "mgb_dcsr_inverse_diagonal: null or size mismatch");
	if (A->m == 0) return MGB_OK;
	k_sp_invdiag<<<(A->m + 255) / 256, 256>>>(A->diag, A->va, A->m, dinv->d); SLAUNCHED(); SKCHECK(); return MGB_OK;
}

// ---- level sets of the triangular parts: level(i) = 1 + max level(j) over the rows j that row i reads updated values of
static int build_levels(mgb_dcsr *A)
{
	if (A->sched) return MGB_OK;
	const int m = A->m;
	if (A->m != A->n) return sfail(MGB_EINVAL, "sequential sweeps need a square matrix");
	for (int i = 0; i < m; ++i) if (A->h_diag[i] < 0) return sfail(MGB_EINVAL, "sequential sweeps: missing diagonal in row %d", i);
	// structural symmetry: (i,j) present  =>  (j,i) present.  With it, a row that row i reads OLD values of (the other
	// triangle) always lies in a later level set than row i, as in the sequential loop.
	for (int i = 0; i < m; ++i)
		for (int k = A->h_ia[i]; k < A->h_ia[i + 1]; ++k) {
			const int j = A->h_ja[k];
			if (j == i) continue;
			const int *b = A->h_ja.data() + A->h_ia[j], *e = A->h_ja.data() + A->h_ia[j + 1];
			bool found = false;
			while (b < e) { const int *mid = b + (e - b) / 2; if (*mid == i) { found = true; break; } if (*mid < i) b = mid + 1; else e = mid; }
			if (!found) return sfail(MGB_EINVAL, "sequential sweeps need a structurally symmetric matrix: (%d,%d) has no transpose entry", i, j);
		}
	auto make = [&](bool forward, Levels &L) -> int {
		std::vector<int> lev((size_t)(m > 0 ? m : 1), 0);
		int nlev = 0;
		if (forward) {
			for (int i = 0; i < m; ++i) {
				int l = 0;
				for (int k = A->h_ia[i]; k < A->h_diag[i]; ++k) l = lev[A->h_ja[k]] + 1 > l ? lev[A->h_ja[k]] + 1 : l;
				lev[i] = l; if (l + 1 > nlev) nlev = l + 1;
			}
		} else {
			for (int i = m - 1; i >= 0; --i) {
				int l = 0;
				for (int k = A->h_diag[i] + 1; k < A->h_ia[i + 1]; ++k) l = lev[A->h_ja[k]] + 1 > l ? lev[A->h_ja[k]] + 1 : l;
				lev[i] = l; if (l + 1 > nlev) nlev = l + 1;
			}
		}
		std::vector<int> ptr((size_t)nlev + 1, 0), rows((size_t)(m > 0 ? m : 1));
		for (int i = 0; i < m; ++i) ptr[lev[i] + 1]++;
		for (int l = 0; l < nlev; ++l) ptr[l + 1] += ptr[l];
		std::vector<int> fill(ptr.begin(), ptr.end() - 1);
		for (int i = 0; i < m; ++i) rows[fill[lev[i]]++] = i;
		L.nlev = nlev;
		SCU(cudaMalloc(&L.ptr, sizeof(int) * (size_t)(nlev + 1)));
		SCU(cudaMalloc(&L.rows, sizeof(int) * (size_t)(m > 0 ? m : 1)));
		SCU(cudaMemcpy(L.ptr, ptr.data(), sizeof(int) * (size_t)(nlev + 1), cudaMemcpyHostToDevice));
		if (m > 0) SCU(cudaMemcpy(L.rows, rows.data(), sizeof(int) * (size_t)m, cudaMemcpyHostToDevice));
		return MGB_OK;
	};
	STRY(make(true, A->fwd));
	STRY(make(false, A->bwd));
	A->sched = true;
	return MGB_OK;
}

// One launch of ONE thread block walks the level sets of a sweep (a __syncthreads between two sets; the data is L2 / L1
// resident and every row is written by exactly one thread).  op selects the row body.
enum { SW_SOR_ZF = 0,      // zero guess, forward:   t = b - L x ; x = t * idiag
       SW_SOR_ZB_T,        // zero guess, backward after a forward sweep (xb = t):  x = (1-w) x + (t - U x) * idiag
       SW_SOR_ZB_B,        // zero guess, backward only (xb = b):                  x = (b - U x) * idiag
       SW_SOR_F,           // forward:  t = b - L x ; x = (1-w) x + (t - U x) * idiag
       SW_SOR_B_T,         // backward after a forward sweep:  x = (1-w) x + (t - U x) * idiag
       SW_SOR_B_B,         // backward only: whole row, x = (1-w) x + (b - A x + d x) * idiag
       SW_ILU_FACTOR, SW_ILU_FWD, SW_ILU_BWD };
struct SweepArgs {
	int op, nlev; const int *ptr, *rows;
	const int *ia, *ja, *diag; const double *aa;
	const double *b; double *x, *t; const double *idiag, *mdiag; double om1;
	double *fac; int *status;
};
__global__ void __launch_bounds__(SP_SWEEP_THREADS)
k_sp_sweep(SweepArgs a)
{
	for (int l = 0; l < a.nlev; ++l) {
		const int p0 = a.ptr[l], p1 = a.ptr[l + 1];
		for (int q = p0 + threadIdx.x; q < p1; q += SP_SWEEP_THREADS) {
			const int i = a.rows[q];
			const int r0 = a.ia[i], rd = a.diag[i], r1 = a.ia[i + 1];
			double sum;
			switch (a.op) {
			case SW_SOR_ZF:
				sum = a.b[i];
				for (int k = r0; k < rd; ++k) sum = sub(sum, mul(a.aa[k], a.x[a.ja[k]]));
				a.t[i] = sum;
				a.x[i] = mul(sum, a.idiag[i]);
				break;
			case SW_SOR_ZB_T: case SW_SOR_B_T:
				sum = a.t[i];
				for (int k = rd + 1; k < r1; ++k) sum = sub(sum, mul(a.aa[k], a.x[a.ja[k]]));
				a.x[i] = add(mul(a.om1, a.x[i]), mul(sum, a.idiag[i]));
				break;
			case SW_SOR_ZB_B:
				sum = a.b[i];
				for (int k = rd + 1; k < r1; ++k) sum = sub(sum, mul(a.aa[k], a.x[a.ja[k]]));
				a.x[i] = mul(sum, a.idiag[i]);
				break;
			case SW_SOR_F:
				sum = a.b[i];
				for (int k = r0; k < rd; ++k) sum = sub(sum, mul(a.aa[k], a.x[a.ja[k]]));
				a.t[i] = sum;
				for (int k = rd + 1; k < r1; ++k) sum = sub(sum, mul(a.aa[k], a.x[a.ja[k]]));
				a.x[i] = add(mul(a.om1, a.x[i]), mul(sum, a.idiag[i]));
				break;
			case SW_SOR_B_B:
				sum = a.b[i];
				for (int k = r0; k < r1; ++k) sum = sub(sum, mul(a.aa[k], a.x[a.ja[k]]));
				a.x[i] = add(mul(a.om1, a.x[i]), mul(add(sum, mul(a.mdiag[i], a.x[i])), a.idiag[i]));
				break;
			case SW_ILU_FACTOR:
				// row i of MatLUFactorNumeric on the pattern of A (natural ordering, inverted pivots): for every lower entry k
				// (pivot row = ja[k], complete: it lies in an earlier level set) mult = a_ik * (1 / pivot); the pivot row's upper
				// entries update the entries of row i that exist in the pattern
				for (int k = r0; k < rd; ++k) {
					const int row = a.ja[k];
					if (a.fac[k] != 0.0) {
						const double mult = mul(a.fac[k], a.fac[a.diag[row]]);
						a.fac[k] = mult;
						for (int tt = a.diag[row] + 1; tt < a.ia[row + 1]; ++tt) {
							const int c = a.ja[tt];
							int lo = r0, hi = r1;                      // binary search of column c in row i
							while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.ja[mid] < c) lo = mid + 1; else hi = mid; }
							if (lo < r1 && a.ja[lo] == c) a.fac[lo] = sub(a.fac[lo], mul(mult, a.fac[tt]));
						}
					}
				}
				if (a.fac[rd] == 0.0) atomicExch(a.status, i + 1);
				a.fac[rd] = __ddiv_rn(1.0, a.fac[rd]);
				break;
			case SW_ILU_FWD:
				sum = a.b[i];
				for (int k = r0; k < rd; ++k) sum = sub(sum, mul(a.fac[k], a.t[a.ja[k]]));
				a.t[i] = sum;
				break;
			case SW_ILU_BWD:
				sum = a.t[i];
				for (int k = rd + 1; k < r1; ++k) sum = sub(sum, mul(a.fac[k], a.t[a.ja[k]]));
				sum = mul(sum, a.fac[rd]);
				a.t[i] = sum; a.x[i] = sum;
				break;
			}
		}
		__syncthreads();
	}
}
static int sweep(mgb_dcsr *A, int op, bool forward, const double *b, double *x, double *t, double om1)
{
	const Levels &L = forward ? A->fwd : A->bwd;
	SweepArgs a; memset(&a, 0, sizeof a);
	a.op = op; a.nlev = L.nlev; a.ptr = L.ptr; a.rows = L.rows;
	a.ia = A->ia; a.ja = A->ja; a.diag = A->diag; a.aa = A->va;
	a.b = b; a.x = x; a.t = t; a.idiag = A->idiag; a.mdiag = A->mdiag; a.om1 = om1; a.fac = A->fac; a.status = A->status;
	k_sp_sweep<<<1, SP_SWEEP_THREADS>>>(a); SLAUNCHED(); SKCHECK();
	return MGB_OK;
}

// MatInvertDiagonal_SeqAIJ: idiag = omega / (fshift + d) (1 / d when omega == 1 and fshift == 0), mdiag = d
__global__ void __launch_bounds__(256)
k_sp_sor_diag(const int *__restrict__ diag, const double *__restrict__ va, int m, double omega, double fshift,
              double *__restrict__ idiag, double *__restrict__ mdiag, int *status)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const double d = va[diag[i]];
	mdiag[i] = d;
	if (omega == 1.0 && fshift == 0.0) { if (d == 0.0) atomicExch(status, i + 1); idiag[i] = __ddiv_rn(1.0, d); }
	else idiag[i] = __ddiv_rn(omega, add(fshift, d));
}
static int check_status(mgb_dcsr *A, const char *what)
{
	int st = 0;
	SCU(cudaMemcpy(&st, A->status, sizeof(int), cudaMemcpyDeviceToHost));
	if (st) { cudaMemset(A->status, 0, sizeof(int)); return sfail(MGB_EINVAL, "%s: zero pivot / diagonal in row %d", what, st - 1); }
	return MGB_OK;
}

extern "C" int mgb_dcsr_sor(mgb_dcsr *A, const mgb_dvec *b, double omega, int flag, double fshift, int its, int lits, mgb_dvec *x)
{
	if (!A || !b || !x || b->n != A->m || x->n != A->m) return sfail(MGB_EINVAL, "mgb_dcsr_sor: null or size mismatch");
	if (flag & (32 | 64 | 128)) return sfail(MGB_EINVAL, "mgb_dcsr_sor: Eisenstat / apply-upper / apply-lower are not offered");
	if (its <= 0 || lits <= 0) return sfail(MGB_EINVAL, "mgb_dcsr_sor: its=%d lits=%d must be positive", its, lits);
	if (b == x) return sfail(MGB_EINVAL, "mgb_dcsr_sor: b and x must differ");
	if (A->m == 0) return MGB_OK;
	STRY(build_levels(A));
	const size_t mb = sizeof(double) * (size_t)A->m;
	if (!A->idiag) { SCU(cudaMalloc(&A->idiag, mb)); SCU(cudaMalloc(&A->mdiag, mb)); SCU(cudaMalloc(&A->ssor_t, mb)); }
	if (!A->idiag_valid || omega != A->sor_omega || fshift != A->sor_fshift) {
		k_sp_sor_diag<<<(A->m + 255) / 256, 256>>>(A->diag, A->va, A->m, omega, fshift, A->idiag, A->mdiag, A->status); SLAUNCHED(); SKCHECK();
		STRY(check_status(A, "MatSOR"));
		A->sor_omega = omega; A->sor_fshift = fshift; A->idiag_valid = true;
	}
	its = its * lits;
	const bool fwd = (flag & 1) || (flag & 4), bwd = (flag & 2) || (flag & 8);
	const double om1 = 1.0 - omega;
	if (flag & 16) {
		if (fwd) STRY(sweep(A, SW_SOR_ZF, true, b->d, x->d, A->ssor_t, om1));
		if (bwd) STRY(sweep(A, fwd ? SW_SOR_ZB_T : SW_SOR_ZB_B, false, b->d, x->d, A->ssor_t, om1));
		its--;
	}
	while (its-- > 0) {
		if (fwd) STRY(sweep(A, SW_SOR_F, true, b->d, x->d, A->ssor_t, om1));
		if (bwd) STRY(sweep(A, fwd ? SW_SOR_B_T : SW_SOR_B_B, false, b->d, x->d, A->ssor_t, om1));
	}
	return MGB_OK;
}

extern "C" int mgb_dcsr_ilu0_factor(mgb_dcsr *A)
{
	if (!A) return sfail(MGB_EINVAL, "mgb_dcsr_ilu0_factor: null");
	if (A->ilu_valid || A->m == 0) return MGB_OK;
	STRY(build_levels(A));
	if (!A->fac) { SCU(cudaMalloc(&A->fac, sizeof(double) * (size_t)(A->nnz > 0 ? A->nnz : 1))); }
	if (!A->tmp) { SCU(cudaMalloc(&A->tmp, sizeof(double) * (size_t)A->m)); SCU(cudaMemset(A->tmp, 0, sizeof(double) * (size_t)A->m)); }
	SCU(cudaMemcpyAsync(A->fac, A->va, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToDevice, 0));
	STRY(sweep(A, SW_ILU_FACTOR, true, nullptr, nullptr, nullptr, 0.0));
	STRY(check_status(A, "ILU(0)"));
	A->ilu_valid = true;
	return MGB_OK;
}
extern "C" int mgb_dcsr_ilu0_solve(mgb_dcsr *A, const mgb_dvec *b, mgb_dvec *x)
{
	if (!A || !b || !x || b->n != A->m || x->n != A->m) return sfail(MGB_EINVAL, "mgb_dcsr_ilu0_solve: null or size mismatch");
	if (!A->ilu_valid) return sfail(MGB_ESTATE, "mgb_dcsr_ilu0_solve before mgb_dcsr_ilu0_factor");
	if (A->m == 0) return MGB_OK;
	STRY(sweep(A, SW_ILU_FWD, true, b->d, x->d, A->tmp, 0.0));
	STRY(sweep(A, SW_ILU_BWD, false, b->d, x->d, A->tmp, 0.0));
	return MGB_OK;
}

// ---- dense LU without pivoting, natural ordering, inverted pivots (PCLU on a small coarse grid).  Right-looking: at step k
// every entry (i, j > k) of the trailing block receives its k-th update; per entry the updates arrive in ascending k and
// d[i][k], d[k][j] are final when used -- the operation sequence of the row-by-row loop, entry by entry.
__global__ void __launch_bounds__(256)
k_sp_dense_fill(const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ va, int m, double *__restrict__ d)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	for (int k = ia[i]; k < ia[i + 1]; ++k) d[(size_t)i * m + ja[k]] = va[k];
}
__global__ void __launch_bounds__(SP_SWEEP_THREADS)
k_sp_dense_lu(double *__restrict__ d, int m, int *status)
{
	for (int k = 0; k < m; ++k) {
		if (threadIdx.x == 0) {
			if (d[(size_t)k * m + k] == 0.0) atomicExch(status, k + 1);
			d[(size_t)k * m + k] = __ddiv_rn(1.0, d[(size_t)k * m + k]);
		}
		__syncthreads();
		const double piv = d[(size_t)k * m + k];
		for (int i = k + 1 + (int)threadIdx.x; i < m; i += SP_SWEEP_THREADS) {
			const double e = d[(size_t)i * m + k];
			if (e != 0.0) d[(size_t)i * m + k] = mul(e, piv);
		}
		__syncthreads();
		const int w = m - k - 1;
		for (long long q = threadIdx.x; q < (long long)w * w; q += SP_SWEEP_THREADS) {
			const int i = k + 1 + (int)(q / w), j = k + 1 + (int)(q % w);
			const double mult = d[(size_t)i * m + k];
			if (mult != 0.0) d[(size_t)i * m + j] = sub(d[(size_t)i * m + j], mul(mult, d[(size_t)k * m + j]));
		}
		__syncthreads();
	}
}
// forward substitution column by column (tmp_i -= d_ik tmp_k for all i > k: ascending k per row, as in the row loop), then the
// backward substitution row by row on one thread (its row sums run over ascending k, which needs every later row complete)
__global__ void __launch_bounds__(SP_SWEEP_THREADS)
k_sp_dense_solve(const double *__restrict__ d, int m, const double *__restrict__ b, double *__restrict__ tmp, double *__restrict__ x)
{
	for (int i = threadIdx.x; i < m; i += SP_SWEEP_THREADS) tmp[i] = b[i];
	__syncthreads();
	for (int k = 0; k < m; ++k) {
		const double tk = tmp[k];
		for (int i = k + 1 + (int)threadIdx.x; i < m; i += SP_SWEEP_THREADS) tmp[i] = sub(tmp[i], mul(d[(size_t)i * m + k], tk));
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		for (int i = m - 1; i >= 0; --i) {
			double sum = tmp[i];
			for (int k = i + 1; k < m; ++k) sum = sub(sum, mul(d[(size_t)i * m + k], tmp[k]));
			sum = mul(sum, d[(size_t)i * m + i]);
			tmp[i] = sum; x[i] = sum;
		}
	}
}
extern "C" int mgb_dcsr_lu_factor(mgb_dcsr *A)
{
	if (!A) return sfail(MGB_EINVAL, "mgb_dcsr_lu_factor: null");
	if (A->lu_valid || A->m == 0) return MGB_OK;
	if (A->m != A->n) return sfail(MGB_EINVAL, "LU: matrix not square");
	if (A->m > 4096) return sfail(MGB_EINVAL, "LU: the dense coarse solve handles at most 4096 unknowns (got %d)", A->m);
	const size_t mm = (size_t)A->m * A->m;
	if (!A->dense) SCU(cudaMalloc(&A->dense, sizeof(double) * mm));
	if (!A->tmp) { SCU(cudaMalloc(&A->tmp, sizeof(double) * (size_t)A->m)); }
	SCU(cudaMemset(A->dense, 0, sizeof(double) * mm));
	k_sp_dense_fill<<<(A->m + 255) / 256, 256>>>(A->ia, A->ja, A->va, A->m, A->dense); SLAUNCHED(); SKCHECK();
	k_sp_dense_lu<<<1, SP_SWEEP_THREADS>>>(A->dense, A->m, A->status); SLAUNCHED(); SKCHECK();
	STRY(check_status(A, "LU"));
	A->lu_valid = true;
	return MGB_OK;
}
extern "C" int mgb_dcsr_lu_solve(mgb_dcsr *A, const mgb_dvec *b, mgb_dvec *x)
{
	if (!A || !b || !x || b->n != A->m || x->n != A->m) return sfail(MGB_EINVAL, "mgb_dcsr_lu_solve: null or size mismatch");
	if (!A->lu_valid) return sfail(MGB_ESTATE, "mgb_dcsr_lu_solve before mgb_dcsr_lu_factor");
	if (A->m == 0) return MGB_OK;
	k_sp_dense_solve<<<1, SP_SWEEP_THREADS>>>(A->dense, A->m, b->d, A->tmp, x->d); SLAUNCHED(); SKCHECK();
	return MGB_OK;
}
