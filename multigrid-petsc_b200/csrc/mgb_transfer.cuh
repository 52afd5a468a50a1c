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
	pdl_enter();
	const int J = blockIdx.x * blockDim.x + threadIdx.x;
	const int I = blockIdx.y * blockDim.y + threadIdx.y;
	if (I >= C.ni || J >= C.pitch) return;
	double out = 0.0;
	if (J < C.nj) {
		const size_t P = (size_t)F.pitch;
		double r[9];
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
				if (FUSED) {
					// F.i0 is even (strips start on even fine rows), so the colour of fine (2I+a, 2J+b) is (a+b) & 1
					const int ord = F.rb ? 1 + ((a + b) & 1) : 0;
					const double t = stencil5_ord(ord, aS, aW, aC, aE, aN, uf[o - P], uf[o - 1], uf[o], uf[o + 1], uf[o + P]);
					r[a * 3 + b] = sub(bf[o], t);
				} else {
					r[a * 3 + b] = rf[o];
				}
			}
		}
		double sum;
		if (!F.rb) {
			sum = mul(R.w[0], r[0]);
#pragma unroll
			for (int k = 1; k < 9; ++k) sum = add(sum, mul(R.w[k], r[k]));
		} else {
			// red-first numbering of the fine grid: the five red fine points, then the four black ones
			sum = mul(R.w[0], r[0]);
			sum = add(sum, mul(R.w[2], r[2]));
			sum = add(sum, mul(R.w[4], r[4]));
			sum = add(sum, mul(R.w[6], r[6]));
			sum = add(sum, mul(R.w[8], r[8]));
			sum = add(sum, mul(R.w[1], r[1]));
			sum = add(sum, mul(R.w[3], r[3]));
			sum = add(sum, mul(R.w[5], r[5]));
			sum = add(sum, mul(R.w[7], r[7]));
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
	pdl_enter();
	const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
	const int i = blockIdx.y * blockDim.y + threadIdx.y;
	if (i >= F.ni || j0 >= F.pitch) return;
	const size_t o = (size_t)i * F.pitch + j0;
	double2 u = ld2(uf + o);
	const int J0 = j0 >> 1, Jm = J0 - 1;
	const size_t PC = (size_t)C.pitch;
	// up to four terms for column j0 (t0..t3) and two for column j0+1 (s0, s1), in the order of the entries of the
	// row of pro[l]: ascending coarse number -- natural, or red-first (-map 3) where a red coarse point precedes a black one
	double t0, t1, t2 = 0.0, t3 = 0.0, s0, s1 = 0.0;
	int nt, ns;
	if (i & 1) {
		// a = 1: single coarse row I = (i-1)/2
		const int I = (i - 1) >> 1;
		const double *c = uc + (size_t)I * PC;
		const double cm = mul(Pw.w[3 + 2], c[Jm]), c0 = mul(Pw.w[3 + 0], c[J0]);
		const bool swap = F.rb && (((C.i0 + I + Jm) & 1) != 0);      // (I,Jm) black: the red (I,J0) comes first
		t0 = swap ? c0 : cm; t1 = swap ? cm : c0; nt = 2;
		s0 = mul(Pw.w[3 + 1], c[J0]); ns = 1;
	} else {
		// coarse rows IA = i/2 - 1 (a = 2) then IB = i/2 (a = 0)
		const int IA = (i >> 1) - 1;
		const double *cA = uc + (size_t)IA * PC;     // row -1 is the ghost row above the strip
		const double *cB = cA + PC;
		const double am = mul(Pw.w[6 + 2], cA[Jm]), a0 = mul(Pw.w[6 + 0], cA[J0]);
		const double bm = mul(Pw.w[0 + 2], cB[Jm]), b0 = mul(Pw.w[0 + 0], cB[J0]);
		const double sa = mul(Pw.w[6 + 1], cA[J0]), sb = mul(Pw.w[0 + 1], cB[J0]);
		nt = 4; ns = 2;
		if (!F.rb) { t0 = am; t1 = a0; t2 = bm; t3 = b0; s0 = sa; s1 = sb; }
		else {
			const bool red_am = ((C.i0 + IA + Jm) & 1) == 0;          // (IA,Jm) and (IB,J0) share a colour
			if (red_am) { t0 = am; t1 = b0; t2 = a0; t3 = bm; s0 = sb; s1 = sa; }   // (IA,J0) black, (IB,J0) red
			else        { t0 = a0; t1 = bm; t2 = am; t3 = b0; s0 = sa; s1 = sb; }
		}
	}
	double e0, e1;
	if (MULTADD) {
		// MatMultAdd: sum = u ; sum += term ...
		e0 = add(add(u.x, t0), t1);
		if (nt == 4) e0 = add(add(e0, t2), t3);
		e1 = add(u.y, s0);
		if (ns == 2) e1 = add(e1, s1);
	} else {
		// MatMult then VecAXPY(u, 1.0, rv)
		e0 = add(t0, t1);
		if (nt == 4) e0 = add(add(e0, t2), t3);
		e1 = s0;
		if (ns == 2) e1 = add(e1, s1);
		e0 = add(u.x, mul(1.0, e0));
		e1 = add(u.y, mul(1.0, e1));
	}
	double2 out;
	out.x = (j0 < F.nj) ? e0 : 0.0;
	out.y = (j0 + 1 < F.nj) ? e1 : 0.0;
	st2(uf + o, out);
}
