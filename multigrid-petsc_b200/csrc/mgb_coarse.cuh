// mgb_coarse.cuh -- exact coarse-grid solve kept resident on the GPU (PCMG's default coarse solver,
// -mg_coarse_ksp_type preonly -mg_coarse_pc_type lu; ref: src/solver.c:1931-1933 wires A[levels-1] into it).
//
// The coarsest 5-point operator (natural ordering, m = ni*nj unknowns, half bandwidth w = nj) is LU-factored
// without pivoting inside its band on the host once per solve setup (m * w^2 flops, m <= 4096), with inverted
// pivots as PETSc's MatLUFactorNumeric_SeqAIJ stores them.  The triangular solves run in one thread block with
// the running sums in shared memory, column-oriented: after x_k is final every row of the band column k is
// updated in parallel.  The forward sweep applies the updates of each row in ascending k exactly like the
// row-oriented loop of the oracle; the backward sweep applies them in descending k (the oracle: ascending),
// so the result agrees with it to rounding (~1e-16 relative), not bit for bit -- and PETSc's own LU uses a
// nested-dissection ordering, so no bit-level reference exists for this step (oracle/minipetsc/minipetsc.c).
#pragma once
#include "mgb_common.cuh"
#include <vector>
#include <cstdio>

#define MGB_LU_MAX 4096

struct BandLU {
	int m = 0, w = 0;
	double *band = nullptr;     // device: m * (2w+1), entry (r, c) at r*(2w+1) + (c - r + w); diagonal inverted
	bool valid = false;
};

static void bandlu_free(BandLU &f) { cudaFree(f.band); f = BandLU(); }

// coef: ni rows x MGB_COEF_STRIDE (S W C E N ...) of the coarsest level
static int bandlu_factor(BandLU &f, const double *coef, int ni, int nj, char *err, size_t errlen)
{
	if (f.valid) return 0;
	const int m = ni * nj, w = nj, bw = 2 * w + 1;
	if (m > MGB_LU_MAX) {
		snprintf(err, errlen, "coarse LU: %d unknowns on the coarsest level exceed the limit %d; add levels or use "
		                      "-mg_coarse_ksp_type richardson", m, MGB_LU_MAX);
		return -1;
	}
	std::vector<double> d((size_t)m * bw, 0.0);
	auto at = [&](int r, int c) -> double & { return d[(size_t)r * bw + (c - r + w)]; };
	for (int i = 0; i < ni; ++i)
		for (int j = 0; j < nj; ++j) {
			const int r = i * nj + j;
			const double *c = coef + (size_t)i * MGB_COEF_STRIDE;
			if (i > 0) at(r, r - nj) = c[0];
			if (j > 0) at(r, r - 1) = c[1];
			at(r, r) = c[2];
			if (j < nj - 1) at(r, r + 1) = c[3];
			if (i < ni - 1) at(r, r + nj) = c[4];
		}
	for (int i = 0; i < m; ++i) {
		for (int k = (i - w > 0 ? i - w : 0); k < i; ++k) {
			if (at(i, k) != 0.0) {
				const double mult = at(i, k) * at(k, k);
				at(i, k) = mult;
				const int jhi = (k + w < m - 1) ? k + w : m - 1;
				for (int j = k + 1; j <= jhi; ++j) at(i, j) -= mult * at(k, j);
			}
		}
		if (at(i, i) == 0.0) { snprintf(err, errlen, "coarse LU: zero pivot in row %d", i); return -1; }
		at(i, i) = 1.0 / at(i, i);
	}
	cudaError_t ce = cudaMalloc(&f.band, sizeof(double) * d.size());
	if (ce == cudaSuccess) ce = cudaMemcpy(f.band, d.data(), sizeof(double) * d.size(), cudaMemcpyHostToDevice);
	if (ce != cudaSuccess) { snprintf(err, errlen, "coarse LU upload: %s", cudaGetErrorString(ce)); return -2; }
	f.m = m; f.w = w; f.valid = true;
	return 0;
}

__global__ void __launch_bounds__(256)
k_bandlu_solve(const double *__restrict__ band, int m, int w, const double *__restrict__ b, double *__restrict__ x,
               int nj, int pitch)
{
	__shared__ double s[MGB_LU_MAX];
	const int bw = 2 * w + 1;
	for (int r = threadIdx.x; r < m; r += blockDim.x) s[r] = b[(size_t)(r / nj) * pitch + (r % nj)];
	__syncthreads();
	// forward: L has a unit diagonal; multipliers are stored in the strictly lower band
	for (int k = 0; k < m - 1; ++k) {
		const double tk = s[k];
		for (int t = threadIdx.x; t < w; t += blockDim.x) {
			const int i = k + 1 + t;
			if (i < m) s[i] = sub(s[i], mul(band[(size_t)i * bw + (k - i + w)], tk));
		}
		__syncthreads();
	}
	// backward: x_k = s_k * (1/u_kk), then the rows above it in column k
	for (int k = m - 1; k >= 0; --k) {
		if (threadIdx.x == 0) s[k] = mul(s[k], band[(size_t)k * bw + w]);
		__syncthreads();
		const double xk = s[k];
		for (int t = threadIdx.x; t < w; t += blockDim.x) {
			const int i = k - 1 - t;
			if (i >= 0) s[i] = sub(s[i], mul(band[(size_t)i * bw + (k - i + w)], xk));
		}
		__syncthreads();
	}
	for (int r = threadIdx.x; r < m; r += blockDim.x) x[(size_t)(r / nj) * pitch + (r % nj)] = s[r];
}

static int bandlu_solve(const BandLU &f, const double *b, double *x, int ni, int nj, int pitch, cudaStream_t st)
{
	(void)ni;
	k_bandlu_solve<<<1, 256, 0, st>>>(f.band, f.m, f.w, b, x, nj, pitch);
	return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
