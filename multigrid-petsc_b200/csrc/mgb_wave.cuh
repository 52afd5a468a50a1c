// mgb_wave.cuh -- lexicographic (natural-order) sweeps of the 5-point operator as pipelined wavefronts.
//
// PETSc's default smoothers on the reference's path are inherently sequential in the natural numbering:
//   -pc_type sor   MatSOR_SeqAIJ: forward / backward / symmetric Gauss-Seidel sweeps   (restated in oracle/minipetsc: MatSOR)
//   (no -pc_type)  PCILU: ILU(0) factorisation + forward / backward triangular solves   (pc_setup_ilu0 / pc_apply_ilu0)
// (ref: the shipped poisson.in sets no -pc_type; KSPSetFromOptions at src/solver.c:1476,1492,1509).  For the 5-point stencil
// row (i,j) depends on (i-1,j) and (i,j-1) only, so all points of an anti-diagonal i + j = const are independent: visiting
// the anti-diagonals in order performs EXACTLY the operations of the sequential loop, each with the same operands in the same
// order -- the results are bit-identical, only the schedule is parallel.
//
// Schedule.  One thread per grid column, WV_T columns per block ("panel"), one block per panel, all panels resident at the
// same time (<= 148 panels).  At step s thread c of a panel works on row s - c: its upper neighbour is the value it produced
// in the previous step (a register), its left neighbour was produced by thread c-1 in the previous step (shared memory, one
// __syncthreads per step).  The first thread of panel p takes its left neighbour from HBM after panel p-1 has published that
// it finished the row (a progress counter, release / acquire at GPU scope): panel p runs WV_T steps behind panel p-1, and a
// sweep takes rows + columns steps in total.  Backward sweeps run the same schedule on the mirrored grid.
#pragma once
#include "mgb_common.cuh"

#define WV_T 128
enum { WV_SOR_FWD = 0, WV_SOR_BWD_T = 1, WV_SOR_BWD_B = 2, WV_ILU_FACTOR = 3, WV_ILU_FWD = 4, WV_ILU_BWD = 5 };

struct WaveArgs {
	int op, reverse;         // reverse: rows and columns run from the last to the first (backward sweeps / solves)
	LevelDev L;
	double *x;               // SOR: the iterate (in place).  ILU solves: the vector solved in place (tmp, then x).  Factor: unused
	const double *b;         // SOR: right-hand side
	double *t;               // SOR: b - (lower part) x checkpoint, written by FWD and read by BWD_T (PETSc's ssor_work)
	double *invd, *mS, *mW;  // ILU(0): inverted pivots and the two multipliers per point (written by FACTOR)
	double omega;
	int *progress;           // one counter per panel: rows finished (zeroed before the launch)
	long long spin_limit;
	int *status;
};

__device__ __forceinline__ int ld_acquire_gpu_i32(const int *p)
{
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_gpu_i32(int *p, int v)
{
	asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }

__global__ void __launch_bounds__(WV_T)
k_wave(WaveArgs a)
{
	__shared__ double sh[2][WV_T];
	const LevelDev &L = a.L;
	const int c = threadIdx.x, panel = blockIdx.x;
	const int lc = panel * WV_T + c;                     // logical column
	const bool valid = lc < L.nj;
	const int j = a.reverse ? L.nj - 1 - lc : lc;        // physical column
	const size_t P = (size_t)L.pitch;
	const int nsteps = L.ni + WV_T - 1;
	const double om1 = sub(1.0, a.omega);
	double up = 0.0;                                     // the value this thread produced on the previous logical row
	sh[0][c] = 0.0; sh[1][c] = 0.0;
	__syncthreads();
	for (int s = 0; s < nsteps; ++s) {
		const int r = s - c;                             // logical row
		const bool act = valid && r >= 0 && r < L.ni;
		double v = 0.0;
		if (act) {
			const int i = a.reverse ? L.ni - 1 - r : r;  // physical row
			const size_t o = (size_t)i * P + j;
			// the neighbour produced at the previous step by the thread on my left (logical column lc-1, same logical row)
			double left;
			if (lc == 0) left = 0.0;                     // outside the grid: the zero Dirichlet ghost
			else if (c > 0) left = sh[(s + 1) & 1][c - 1];
			else {
				const long long t0 = clock64();
				while (ld_acquire_gpu_i32(a.progress + panel - 1) <= r)
					if (clock64() - t0 > a.spin_limit) { atomicExch(a.status, 2); break; }
				const double *dep = (a.op == WV_ILU_FACTOR) ? a.invd : a.x;
				left = ld_cg(dep + (size_t)i * P + (a.reverse ? j + 1 : j - 1));
			}
			const double *cf = L.coef + (size_t)(L.i0 + i) * MGB_COEF_STRIDE;
			const double aS = cf[0], aW = cf[1], aC = cf[2], aE = cf[3], aN = cf[4];
			const bool hasS = i > 0, hasW = j > 0, hasE = j < L.nj - 1, hasN = i < L.ni - 1;
			// in a forward pass the new values are S (up) and W (left), in a reversed pass N (up) and E (left)
			if (a.op == WV_SOR_FWD) {
				const double xC = ld_cg(a.x + o);
				const double xE = hasE ? ld_cg(a.x + o + 1) : 0.0, xN = hasN ? ld_cg(a.x + o + P) : 0.0;
				double sum = ld_cg(a.b + o);
				sum = sub(sum, mul(aS, up));
				sum = sub(sum, mul(aW, left));
				if (a.t) a.t[o] = sum;
				sum = sub(sum, mul(aE, xE));
				sum = sub(sum, mul(aN, xN));
				v = add(mul(om1, xC), mul(sum, cf[6]));
				a.x[o] = v;
			} else if (a.op == WV_SOR_BWD_T) {
				const double xC = ld_cg(a.x + o);
				double sum = ld_cg(a.t + o);
				sum = sub(sum, mul(aE, left));
				sum = sub(sum, mul(aN, up));
				v = add(mul(om1, xC), mul(sum, cf[6]));
				a.x[o] = v;
			} else if (a.op == WV_SOR_BWD_B) {
				const double xC = ld_cg(a.x + o);
				const double xS = hasS ? ld_cg(a.x + o - P) : 0.0, xW = hasW ? ld_cg(a.x + o - 1) : 0.0;
				double sum = ld_cg(a.b + o);
				sum = sub(sum, mul(aS, xS));
				sum = sub(sum, mul(aW, xW));
				sum = sub(sum, mul(aC, xC));
				sum = sub(sum, mul(aE, left));
				sum = sub(sum, mul(aN, up));
				v = add(mul(om1, xC), mul(add(sum, mul(cf[7], xC)), cf[6]));
				a.x[o] = v;
			} else if (a.op == WV_ILU_FACTOR) {
				// MatLUFactorNumeric_SeqAIJ on the 5-point pattern: the pivot of row (i,j) takes one update from each lower
				// entry -- S: mult = aS * invd(i-1,j), pivot -= mult * aN(i-1) ; W: mult = aW * invd(i,j-1), pivot -= mult * aE(i)
				// (the coefficients depend on the row only); no other fill position lies inside the pattern (n >= 3)
				double piv = aC, mS = 0.0, mW = 0.0;
				if (hasS) {
					const double *cfS = cf - MGB_COEF_STRIDE;
					mS = mul(aS, up);
					piv = sub(piv, mul(mS, cfS[4]));
				}
				if (hasW) {
					mW = mul(aW, left);
					piv = sub(piv, mul(mW, aE));
				}
				v = 1.0 / piv;
				a.invd[o] = v; a.mS[o] = mS; a.mW[o] = mW;
			} else if (a.op == WV_ILU_FWD) {
				// tmp(i) = b(i) - L_S tmp(S) - L_W tmp(W), in place on x
				double sum = ld_cg(a.x + o);
				if (hasS) sum = sub(sum, mul(ld_cg(a.mS + o), up));
				if (hasW) sum = sub(sum, mul(ld_cg(a.mW + o), left));
				v = sum;
				a.x[o] = v;
			} else {
				// x(i) = (tmp(i) - U_E x(E) - U_N x(N)) * invd(i), in place on x
				double sum = ld_cg(a.x + o);
				if (hasE) sum = sub(sum, mul(aE, left));
				if (hasN) sum = sub(sum, mul(aN, up));
				v = mul(sum, ld_cg(a.invd + o));
				a.x[o] = v;
			}
			up = v;
		}
		sh[s & 1][c] = v;
		// the next panel needs the column of this panel's last thread
		if (c == WV_T - 1 && act) { __threadfence(); st_release_gpu_i32(a.progress + panel, r + 1); }
		__syncthreads();
	}
}
