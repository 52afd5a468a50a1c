// mgb_transfer.cuh -- full-weighting restriction and bilinear prolongation, fused with the residual and the
// correction (ref: src/solver.c:1534-1535 and :1540-1541; weights from op.res[0] / op.pro[0],
// src/matbuild.c:398-431).  Coarse point (I,J) sits on fine point (2I+1, 2J+1) (src/solver.c:231-232).
#pragma once
#include "mgb_common.cuh"

struct Stencil3 { double w[9]; };

// b_c[I][J] = sum_{a,b in 0..2} w[a][b] * r_f[2I+a][2J+b], accumulated in ascending fine column order
// (a-major, b-minor) as MatMult_SeqAIJ does on res[l].
// FUSED: r_f = b_f - A_f u_f is recomputed on the fly (no residual vector in HBM): 18 B per fine unknown
//        (8 u + 8 b read, 2 written) instead of 24 + 10.
// One thread per coarse point; the 3x3 fine residuals come from a 5x5 patch of u (L1/L2 resident re-reads).
template <int FUSED>
__global__ void __launch_bounds__(128)
k_restrict(const double *__restrict__ uf, const double *__restrict__ bf, const double *__restrict__ rf,
           double *__restrict__ bc, LevelDev F, LevelDev C, Stencil3 R)
{
	const int J = blockIdx.x * blockDim.x + threadIdx.x;
	const int I = blockIdx.y * blockDim.y + threadIdx.y;
	if (I >= C.ni || J >= C.pitch) return;
	double out = 0.0;
	if (J < C.nj) {
		const size_t P = (size_t)F.pitch;
		double sum = 0.0;
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const int i = 2 * I + a;
			double aS = 0, aW = 0, aC = 0, aE = 0, aN = 0;
			if (FUSED) {
				const double *cf = F.coef + (size_t)(F.i0 + i) * MGB_COEF_STRIDE;
				aS = cf[0]; aW = cf[1]; aC = cf[2]; aE = cf[3]; aN = cf[4];
			}
#pragma unroll
			for (int b = 0; b < 3; ++b) {
				const size_t o = (size_t)i * P + (2 * J + b);
				double r;
				if (FUSED) {
					const double t = stencil5(aS, aW, aC, aE, aN, uf[o - P], uf[o - 1], uf[o], uf[o + 1], uf[o + P]);
					r = sub(bf[o], t);
				} else {
					r = rf[o];
				}
				const double term = mul(R.w[a * 3 + b], r);
				sum = (a == 0 && b == 0) ? term : add(sum, term);
			}
		}
		out = sum;
	}
	bc[(size_t)I * C.pitch + J] = out;
}

// u_f += P u_c.  Gather form: fine (i,j) receives p[a][b] * u_c[I][J] for every coarse (I,J) with
// i = 2I + a, j = 2J + b; the terms are added in ascending coarse column order (I-major), which is the order of
// the entries of row (i,j) of pro[l].  Out-of-range coarse neighbours are the zero ghosts of the coarse array.
//   MULTADD == 0 (cycle 0):  rv = sum of terms from the first ; u = u + 1.0 * rv      (MatMult + VecAXPY)
//   MULTADD == 1 (PCMG):     sum = u ; sum += term ...                                  (MatMultAdd, MatInterpolateAdd)
// One thread per fine column pair (j0 even, j0+1 odd); 18 B per fine unknown (8 read + 8 written + 2 coarse).
template <int MULTADD>
__global__ void __launch_bounds__(128)
k_prolong_add(double *__restrict__ uf, const double *__restrict__ uc, LevelDev F, LevelDev C, Stencil3 Pw)
{
	const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
	const int i = blockIdx.y * blockDim.y + threadIdx.y;
	if (i >= F.ni || j0 >= F.pitch) return;
	const size_t o = (size_t)i * F.pitch + j0;
	double2 u = ld2(uf + o);
	const int J0 = j0 >> 1, Jm = J0 - 1;
	const size_t PC = (size_t)C.pitch;
	double e0, e1;       // corrections (MULTADD 0) or running sums (MULTADD 1) for columns j0, j0+1
	if (i & 1) {
		// a = 1: single coarse row I = (i-1)/2
		const double *c = uc + (size_t)((i - 1) >> 1) * PC;
		const double cm = c[Jm], c0 = c[J0];
		if (MULTADD) {
			e0 = add(add(u.x, mul(Pw.w[3 + 2], cm)), mul(Pw.w[3 + 0], c0));
			e1 = add(u.y, mul(Pw.w[3 + 1], c0));
		} else {
			e0 = add(mul(Pw.w[3 + 2], cm), mul(Pw.w[3 + 0], c0));
			e1 = mul(Pw.w[3 + 1], c0);
		}
	} else {
		// coarse rows I = i/2 - 1 (a = 2) then I = i/2 (a = 0)
		const double *cA = uc + (size_t)((i >> 1) - 1) * PC;   // row -1 is the zero ghost row
		const double *cB = cA + PC;
		const double am = cA[Jm], a0 = cA[J0], bm = cB[Jm], b0 = cB[J0];
		if (MULTADD) {
			e0 = add(add(add(add(u.x, mul(Pw.w[6 + 2], am)), mul(Pw.w[6 + 0], a0)), mul(Pw.w[0 + 2], bm)), mul(Pw.w[0 + 0], b0));
			e1 = add(add(u.y, mul(Pw.w[6 + 1], a0)), mul(Pw.w[0 + 1], b0));
		} else {
			e0 = add(add(add(mul(Pw.w[6 + 2], am), mul(Pw.w[6 + 0], a0)), mul(Pw.w[0 + 2], bm)), mul(Pw.w[0 + 0], b0));
			e1 = add(mul(Pw.w[6 + 1], a0), mul(Pw.w[0 + 1], b0));
		}
	}
	double2 out;
	if (MULTADD) { out.x = e0; out.y = e1; }
	else { out.x = add(u.x, mul(1.0, e0)); out.y = add(u.y, mul(1.0, e1)); }
	if (j0 >= F.nj) out.x = 0.0;
	if (j0 + 1 >= F.nj) out.y = 0.0;
	st2(uf + o, out);
}
