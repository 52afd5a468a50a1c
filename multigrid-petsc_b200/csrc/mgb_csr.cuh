// mgb_csr.cuh -- device-side generation of the assembled (AIJ/CSR) operators and a CSR SpMV.
//
// The reference fills its matrices one MatSetValue at a time on the host (ref: src/solver.c:185-253 for A,
// :1035-1094 for res, :1096-1154 for pro) and PETSc compresses each row with ascending column indices.
// Here row pointers, column indices and values are produced in closed form by one thread per row -- no scan,
// no sort, no atomics -- and are bit-identical to the reference's matrix (tests/test_gpu_parity.py).
// Write-only traffic: 64 B per row of A (5 x 12 + 4), 112 B per row of res, <= 52 B per row of pro.
#pragma once
#include "mgb_common.cuh"
#include "mgb_transfer.cuh"

// number of entries in rows < (i, j) of the 5-point operator on an ni x nj grid (natural numbering).
// Every row has 5 entries minus one per missing neighbour (ref: src/solver.c:239-251 bounds checks).
__device__ __forceinline__ long long rowptr5(int i, int j, int ni, int nj)
{
	const long long row = (long long)i * nj + j;
	const long long dropS = (i == 0) ? j : nj;                 // rows of grid row 0 have no south neighbour
	const long long dropN = (i == ni - 1) ? j : 0;             // rows of the last grid row have no north neighbour
	const long long dropW = (long long)i + (j > 0 ? 1 : 0);    // one row per grid row has j == 0
	const long long dropE = (long long)i;                      // one row per completed grid row has j == nj-1
	return 5 * row - dropS - dropN - dropW - dropE;
}

// Rows of the grid rows [i0, i0 + nloc) only (a rank's row range, like the local rows of a PETSc MPIAIJ matrix):
// row pointers start at 0 for the first local row, column indices are global.
__global__ void __launch_bounds__(256)
k_csr_A(int *__restrict__ rowptr, int *__restrict__ col, double *__restrict__ val, int ni, int nj,
        const double *__restrict__ coef, int i0, int nloc)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int il = blockIdx.y;
	if (j >= nj) return;
	const int i = i0 + il;
	const int row = i * nj + j;
	const long long base = rowptr5(i0, 0, ni, nj);
	long long p = rowptr5(i, j, ni, nj) - base;
	rowptr[il * nj + j] = (int)p;
	const double *c = coef + (size_t)i * MGB_COEF_STRIDE;
	if (i > 0)      { col[p] = row - nj; val[p] = c[0]; ++p; }
	if (j > 0)      { col[p] = row - 1;  val[p] = c[1]; ++p; }
	                { col[p] = row;      val[p] = c[2]; ++p; }
	if (j < nj - 1) { col[p] = row + 1;  val[p] = c[3]; ++p; }
	if (i < ni - 1) { col[p] = row + nj; val[p] = c[4]; ++p; }
	if (il == nloc - 1 && j == nj - 1) rowptr[il * nj + j + 1] = (int)p;
}

// res[l]: coarse row (I,J) -> 9 entries at fine (2I+a, 2J+b), ascending (ref: src/solver.c:1078-1088)
__global__ void __launch_bounds__(256)
k_csr_R(int *__restrict__ rowptr, int *__restrict__ col, double *__restrict__ val, int nci, int ncj, int nfj, Stencil3 R, int I0)
{
	const int J = blockIdx.x * blockDim.x + threadIdx.x;
	const int Il = blockIdx.y;                       // local coarse grid row; nci = local coarse rows, I0 = first global one
	if (J >= ncj) return;
	const int row = Il * ncj + J;
	const long long p = 9LL * row;
	rowptr[row] = (int)p;
	if (Il == nci - 1 && J == ncj - 1) rowptr[row + 1] = (int)(p + 9);
#pragma unroll
	for (int a = 0; a < 3; ++a)
#pragma unroll
		for (int b = 0; b < 3; ++b) {
			col[p + a * 3 + b] = (2 * (I0 + Il) + a) * nfj + 2 * J + b;
			val[p + a * 3 + b] = R.w[a * 3 + b];
		}
}

// number of coarse indices K (0 <= K < nc) with 2K <= t <= 2K+2, i.e. how many coarse points touch fine index t
__device__ __forceinline__ int touch_count(int t, int nc)
{
	if (t & 1) return ((t - 1) / 2 < nc) ? 1 : 0;
	return ((t / 2 - 1 >= 0 && t / 2 - 1 < nc) ? 1 : 0) + ((t / 2 < nc) ? 1 : 0);
}
// sum_{t' < t} touch_count(t', nc) = number of pairs (K, c) with 2K + c < t, c in 0..2
__device__ __forceinline__ long long touch_prefix(int t, int nc)
{
	// coarse K contributes min(3, max(0, t - 2K)) ; K with t - 2K >= 3 <=> K <= (t-3)/2
	if (t <= 0) return 0;
	long long full = (t >= 3) ? ((t - 3) / 2 + 1) : 0;            // K = 0 .. (t-3)/2 contribute 3
	if (full > nc) full = nc;
	long long s = 3 * full;
	for (long long K = full; K < nc && 2 * K < t; ++K) s += (t - 2 * K);   // at most 2 partial terms
	return s;
}

// pro[l]: fine row (i,j) gets p[a][b] from coarse (I,J) with i = 2I+a, j = 2J+b (transpose pattern of res;
// ref: src/solver.c:1138-1148), entries in ascending coarse column order.
__global__ void __launch_bounds__(256)
k_csr_P(int *__restrict__ rowptr, int *__restrict__ col, double *__restrict__ val, int nfi, int nfj, int nci, int ncj, Stencil3 Pw,
        int i0, int nloc)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int il = blockIdx.y;                       // local fine grid row of the rows [i0, i0 + nloc); nfi, nci are global
	if (j >= nfj) return;
	const int i = i0 + il;
	const int row = il * nfj + j;
	const long long rowsum = touch_prefix(nfj, ncj);             // entries per unit of touch_count(i)
	long long p = (touch_prefix(i, nci) - touch_prefix(i0, nci)) * rowsum + (long long)touch_count(i, nci) * touch_prefix(j, ncj);
	rowptr[row] = (int)p;
	const int Ilo = (i & 1) ? (i - 1) / 2 : i / 2 - 1, Ihi = (i & 1) ? Ilo : i / 2;
	const int Jlo = (j & 1) ? (j - 1) / 2 : j / 2 - 1, Jhi = (j & 1) ? Jlo : j / 2;
	for (int I = Ilo; I <= Ihi; ++I) {
		if (I < 0 || I >= nci) continue;
		for (int J = Jlo; J <= Jhi; ++J) {
			if (J < 0 || J >= ncj) continue;
			col[p] = I * ncj + J;
			val[p] = Pw.w[(i - 2 * I) * 3 + (j - 2 * J)];
			++p;
		}
	}
	if (il == nloc - 1 && j == nfj - 1) rowptr[row + 1] = (int)p;
}

// y = M x, one thread per row, ascending columns, accumulation from 0.0 (MatMult_SeqAIJ).
// x and y are level vectors in the padded layout: natural index k <-> (k / n) * pitch + k % n.
__global__ void __launch_bounds__(256)
k_csr_spmv(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
           int m, const double *__restrict__ x, int xn, int xpitch, double *__restrict__ y, int yn, int ypitch)
{
	const int row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= m) return;
	double sum = 0.0;
	for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
		const int c = col[k];
		const int ci = c / xn;
		sum = add(sum, mul(val[k], x[(size_t)ci * xpitch + (c - ci * xn)]));
	}
	const int ri = row / yn;
	y[(size_t)ri * ypitch + (row - ri * yn)] = sum;
}
