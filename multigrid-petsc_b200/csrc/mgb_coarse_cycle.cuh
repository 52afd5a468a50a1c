// mgb_coarse_cycle.cuh -- the bottom of the V-cycle (every level with at most `coarse_threshold` rows, resident
// whole on one GPU) as ONE persistent launch: a single thread-block cluster of 8 CTAs walks down to the coarsest
// level and back up, one cluster barrier per phase, instead of two launches per level.
//
// Replaces, for levels lp .. L-1, the sequence of the reference's loop body (ref: src/solver.c:1533-1544):
//   down:  KSPSolve (zero guess, its sweeps)  ->  KSPBuildResidual + MatMult(res)           per level
//   coarsest: KSPSolve (zero guess, v1 sweeps)        (the reference has no exact coarse solve in cycle 0, :1508)
//   up:    MatMult(pro) + VecAXPY  ->  KSPSolve (nonzero guess, its sweeps)                  per level
// These levels hold 1/12 of the unknowns but cost ~20 launches of ~10-30 us each; here the whole thing is a few
// dozen phases of ~1 us.  The arithmetic per value is that of the one-sweep kernels (same operation order, no FMA),
// so the results are bit-identical.  Data stays in global memory (L2-resident: <= 0.5 MB per vector); loads use
// ld.global.cg so that no stale L1 line of another CTA's output can be read; cluster.sync() orders the phases.
#pragma once
#include "mgb_common.cuh"
#include "mgb_transfer.cuh"
#include <cooperative_groups.h>

#define CC_CTAS 8
#define CC_THREADS 1024
#define CC_MAXLEV 16

struct CLevel {
	double *x, *w, *b;       // iterate, ping-pong scratch, right-hand side: element (0,0)
	const double *coef;      // MGB_COEF_STRIDE doubles per grid row
	int ni, nj, pitch, uniform;
	int its_down, its_up;    // sweeps on the way down (zero guess) / up (after the correction); coarsest: its_down only
	double scale;            // Richardson damping of this level's smoother
};
struct CoarseArgs {
	int nlev;                // levels lev[0] (finest of the bottom part) .. lev[nlev-1] (coarsest)
	int multadd;             // 1: PCMG's MatInterpolateAdd order in the correction step
	CLevel lev[CC_MAXLEV];
	Stencil3 R3, P3;
};

namespace ccy {
// G = true: data in global memory, possibly written by another CTA of the cluster in the previous phase (ld.global.cg);
// G = false: data in this CTA's shared memory (plain loads)
template <bool G> __device__ __forceinline__ double ld(const double *p) { return G ? __ldcg(p) : *p; }

struct Rows { int gw, GW, lane; };   // this warp's index in the cluster, warps in the cluster, lane

// x = scale * (b * dinv)  (first Richardson iteration from a zero guess)
template <bool G>
__device__ __forceinline__ void first_sweep(const CLevel &L, double *x, const Rows &R)
{
	for (int i = R.gw; i < L.ni; i += R.GW) {
		const double dinv = L.coef[(size_t)i * MGB_COEF_STRIDE + 5];
		for (int j = R.lane; j < L.pitch; j += 32) {
			const size_t o = (size_t)i * L.pitch + j;
			x[o] = (j < L.nj) ? mul(L.scale, mul(ld<G>(L.b + o), dinv)) : 0.0;
		}
	}
}
// w = x + scale * ((b - A x) * dinv)
template <bool G>
__device__ __forceinline__ void sweep(const CLevel &L, const double *x, double *w, const Rows &R)
{
	const ptrdiff_t P = L.pitch;
	if (G && (L.pitch & 127) == 0) {
		// wide rows: four 32-column chunks at a time, all loads issued before the arithmetic (one L2 round trip per
		// 128 columns).  Pad columns are computed and discarded: every address is inside the padded row.
		for (int i = R.gw; i < L.ni; i += R.GW) {
			const double *cf = L.coef + (size_t)i * MGB_COEF_STRIDE;
			const double aS = cf[0], aW = cf[1], aC = cf[2], aE = cf[3], aN = cf[4], dinv = cf[5];
			for (int jb = R.lane; jb < L.pitch; jb += 128) {
				double xs[4], xw[4], xc[4], xe[4], xn[4], bb[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const ptrdiff_t o = (ptrdiff_t)i * P + jb + 32 * k;
					xs[k] = ld<G>(x + o - P); xw[k] = ld<G>(x + o - 1); xc[k] = ld<G>(x + o);
					xe[k] = ld<G>(x + o + 1); xn[k] = ld<G>(x + o + P); bb[k] = ld<G>(L.b + o);
				}
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const int j = jb + 32 * k;
					const double t = stencil5(aS, aW, aC, aE, aN, xs[k], xw[k], xc[k], xe[k], xn[k]);
					const double out = add(xc[k], mul(L.scale, mul(sub(bb[k], t), dinv)));
					w[(ptrdiff_t)i * P + j] = (j < L.nj) ? out : 0.0;
				}
			}
		}
		return;
	}
	for (int i = R.gw; i < L.ni; i += R.GW) {
		const double *cf = L.coef + (size_t)i * MGB_COEF_STRIDE;
		const double aS = cf[0], aW = cf[1], aC = cf[2], aE = cf[3], aN = cf[4], dinv = cf[5];
		for (int j = R.lane; j < L.pitch; j += 32) {
			const ptrdiff_t o = (ptrdiff_t)i * P + j;
			double out = 0.0;
			if (j < L.nj) {
				const double xc = ld<G>(x + o);
				const double t = stencil5(aS, aW, aC, aE, aN, ld<G>(x + o - P), ld<G>(x + o - 1), xc, ld<G>(x + o + 1), ld<G>(x + o + P));
				out = add(xc, mul(L.scale, mul(sub(ld<G>(L.b + o), t), dinv)));
			}
			w[o] = out;
		}
	}
}
// bc = res * (b - A x)   (k_restrict<1>, natural numbering)
template <bool G>
__device__ __forceinline__ void residual_restrict(const CLevel &F, const double *x, const CLevel &C, const Stencil3 &Rw, const Rows &R)
{
	const ptrdiff_t P = F.pitch;
	for (int I = R.gw; I < C.ni; I += R.GW) {
		for (int J = R.lane; J < C.pitch; J += 32) {
			double out = 0.0;
			if (J < C.nj) {
				double sum = 0.0;
#pragma unroll
				for (int a = 0; a < 3; ++a) {
					const int i = 2 * I + a;
					const double *cf = F.coef + (size_t)i * MGB_COEF_STRIDE;
					const double aS = cf[0], aW = cf[1], aC = cf[2], aE = cf[3], aN = cf[4];
#pragma unroll
					for (int b = 0; b < 3; ++b) {
						const ptrdiff_t o = (ptrdiff_t)i * P + (2 * J + b);
						const double t = stencil5(aS, aW, aC, aE, aN, ld<G>(x + o - P), ld<G>(x + o - 1), ld<G>(x + o), ld<G>(x + o + 1), ld<G>(x + o + P));
						const double term = mul(Rw.w[a * 3 + b], sub(ld<G>(F.b + o), t));
						sum = (a == 0 && b == 0) ? term : add(sum, term);
					}
				}
				out = sum;
			}
			C.b[(size_t)I * C.pitch + J] = out;
		}
	}
}
// x += pro * xc   (k_prolong_add, natural numbering; one thread per fine point)
template <int MULTADD, bool G>
__device__ __forceinline__ void prolong_add(const CLevel &F, double *x, const CLevel &C, const double *xc, const Stencil3 &Pw, const Rows &R)
{
	const ptrdiff_t PC = C.pitch;
	for (int i = R.gw; i < F.ni; i += R.GW) {
		for (int j = R.lane; j < F.pitch; j += 32) {
			const size_t o = (size_t)i * F.pitch + j;
			double out = 0.0;
			if (j < F.nj) {
				const double u = ld<G>(x + o);
				const int J0 = j >> 1, Jm = J0 - 1;
				if (i & 1) {
					const double *c = xc + (ptrdiff_t)((i - 1) >> 1) * PC;
					if (j & 1) {
						const double s0 = mul(Pw.w[3 + 1], ld<G>(c + J0));
						out = MULTADD ? add(u, s0) : add(u, mul(1.0, s0));
					} else {
						const double cm = mul(Pw.w[3 + 2], ld<G>(c + Jm)), c0 = mul(Pw.w[3 + 0], ld<G>(c + J0));
						out = MULTADD ? add(add(u, cm), c0) : add(u, mul(1.0, add(cm, c0)));
					}
				} else {
					const double *cA = xc + (ptrdiff_t)((i >> 1) - 1) * PC;     // row -1: the zero ghost row
					const double *cB = cA + PC;
					if (j & 1) {
						const double sa = mul(Pw.w[6 + 1], ld<G>(cA + J0)), sb = mul(Pw.w[0 + 1], ld<G>(cB + J0));
						out = MULTADD ? add(add(u, sa), sb) : add(u, mul(1.0, add(sa, sb)));
					} else {
						const double am = mul(Pw.w[6 + 2], ld<G>(cA + Jm)), a0 = mul(Pw.w[6 + 0], ld<G>(cA + J0));
						const double bm = mul(Pw.w[0 + 2], ld<G>(cB + Jm)), b0 = mul(Pw.w[0 + 0], ld<G>(cB + J0));
						out = MULTADD ? add(add(add(add(u, am), a0), bm), b0) : add(u, mul(1.0, add(add(add(am, a0), bm), b0)));
					}
				}
			}
			x[o] = out;
		}
	}
}
}  // namespace ccy

// the whole sub-cycle of levels lo .. nlev-1 inside ONE CTA with everything in shared memory (levels of at most
// CC_TINY rows): __syncthreads between phases instead of cluster barriers and L2 round trips
#define CC_TINY 31
#define CC_TINY_DOUBLES 1792              // per vector, all tiny levels together (ghost row above and below each level)
#define CC_TINY_COEF 512                  // coefficient rows of all tiny levels (8 doubles per grid row)
template <int MULTADD>
__device__ void tiny_cycle(const CoarseArgs &A, int lo, double *sx, double *sw, double *sb, double *sc, const ccy::Rows &R)
{
	CLevel T[6];
	const int nt = A.nlev - lo;
	size_t off = 0, coff = 0;
	for (int k = 0; k < nt; ++k) {
		T[k] = A.lev[lo + k];
		// stencil coefficients into shared memory too: a global load per row and phase costs more than the phase
		for (int q = threadIdx.x; q < T[k].ni * MGB_COEF_STRIDE; q += blockDim.x) sc[coff + q] = A.lev[lo + k].coef[q];
		T[k].coef = sc + coff;
		coff += (size_t)T[k].ni * MGB_COEF_STRIDE;
		off += T[k].pitch;                                   // ghost row above (its last element is the left ghost of row 0)
		T[k].x = sx + off; T[k].w = sw + off; T[k].b = sb + off;
		off += (size_t)(T[k].ni + 1) * T[k].pitch;           // rows 0..ni-1 and the ghost row below
	}
	for (size_t k = threadIdx.x; k < off; k += blockDim.x) { sx[k] = 0.0; sw[k] = 0.0; sb[k] = 0.0; }
	__syncthreads();
	// right-hand side of the first tiny level: written to global memory by the whole cluster in the previous phase
	for (int k = threadIdx.x; k < T[0].ni * T[0].pitch; k += blockDim.x) T[0].b[k] = __ldcg(A.lev[lo].b + k);
	__syncthreads();
	unsigned swapped = 0u;
	auto X = [&](int l) { return (swapped >> l) & 1u ? T[l].w : T[l].x; };
	auto W = [&](int l) { return (swapped >> l) & 1u ? T[l].x : T[l].w; };
	for (int l = 0; l < nt; ++l) {
		ccy::first_sweep<false>(T[l], X(l), R);
		__syncthreads();
		for (int k = 1; k < T[l].its_down; ++k) { ccy::sweep<false>(T[l], X(l), W(l), R); __syncthreads(); swapped ^= 1u << l; }
		if (l + 1 < nt) { ccy::residual_restrict<false>(T[l], X(l), T[l + 1], A.R3, R); __syncthreads(); }
	}
	for (int l = nt - 2; l >= 0; --l) {
		ccy::prolong_add<MULTADD, false>(T[l], X(l), T[l + 1], X(l + 1), A.P3, R);
		__syncthreads();
		for (int k = 0; k < T[l].its_up; ++k) { ccy::sweep<false>(T[l], X(l), W(l), R); __syncthreads(); swapped ^= 1u << l; }
	}
	// the correction of the first tiny level goes back to global memory, into the buffer the host expects after the
	// same number of swaps; the deeper levels' vectors are scratch and need not be written back
	double *gx = (swapped & 1u) ? A.lev[lo].w : A.lev[lo].x;
	double *sxf = X(0);
	for (int k = threadIdx.x; k < T[0].ni * T[0].pitch; k += blockDim.x) gx[k] = sxf[k];
}

__global__ void __cluster_dims__(CC_CTAS, 1, 1) __launch_bounds__(CC_THREADS)
k_coarse_cycle(CoarseArgs A)
{
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	__shared__ double sx[CC_TINY_DOUBLES], sw[CC_TINY_DOUBLES], sb[CC_TINY_DOUBLES], sc[CC_TINY_COEF];
	ccy::Rows R;
	R.lane = threadIdx.x & 31;
	R.gw = (int)cluster.block_rank() * (CC_THREADS / 32) + (threadIdx.x >> 5);
	R.GW = CC_CTAS * (CC_THREADS / 32);
	ccy::Rows R1 = R; R1.gw = threadIdx.x >> 5; R1.GW = CC_THREADS / 32;     // CTA-local row distribution
	// first level of the tiny tail (levels of at most CC_TINY rows, at most 6 of them, fitting the shared arrays)
	int lt = A.nlev;
	{
		size_t need = 0, cneed = 0;
		for (int l = A.nlev - 1; l >= 0; --l) {
			if (A.lev[l].ni > CC_TINY || A.lev[l].nj > CC_TINY || A.nlev - l > 6) break;
			need += (size_t)(A.lev[l].ni + 2) * A.lev[l].pitch;
			cneed += (size_t)A.lev[l].ni * MGB_COEF_STRIDE;
			if (need > CC_TINY_DOUBLES || cneed > CC_TINY_COEF) break;
			lt = l;
		}
	}
	// the ping-pong swaps iterate and scratch after every out-of-place sweep: bit l of `swapped` = level l currently
	// has its iterate in lev[l].w (the host applies the same number of swaps to its own pointers after the launch)
	unsigned swapped = 0u;
	auto X = [&](int l) { return (swapped >> l) & 1u ? A.lev[l].w : A.lev[l].x; };
	auto W = [&](int l) { return (swapped >> l) & 1u ? A.lev[l].x : A.lev[l].w; };
	// ---- down (levels 0 .. lt-1 by the whole cluster)
	for (int l = 0; l < lt; ++l) {
		const CLevel &L = A.lev[l];
		ccy::first_sweep<true>(L, X(l), R);
		cluster.sync();
		for (int k = 1; k < L.its_down; ++k) {
			ccy::sweep<true>(L, X(l), W(l), R);
			cluster.sync();
			swapped ^= 1u << l;
		}
		if (l + 1 < A.nlev) {
			ccy::residual_restrict<true>(L, X(l), A.lev[l + 1], A.R3, R);
			cluster.sync();
		}
	}
	// ---- the tiny tail in one CTA's shared memory
	if (lt < A.nlev) {
		if (cluster.block_rank() == 0) {
			if (A.multadd) tiny_cycle<1>(A, lt, sx, sw, sb, sc, R1);
			else           tiny_cycle<0>(A, lt, sx, sw, sb, sc, R1);
		}
		// host-side bookkeeping counts the swaps of every level; the first tiny level's result was stored accordingly
		{
			const CLevel &L = A.lev[lt];
			const int sw_count = (lt == A.nlev - 1) ? L.its_down - 1 : (L.its_down - 1) + L.its_up;
			if (sw_count & 1) swapped ^= 1u << lt;
		}
		cluster.sync();
	}
	// ---- up
	for (int l = lt - 1; l >= 0; --l) {
		if (l + 1 >= A.nlev) continue;
		const CLevel &L = A.lev[l];
		if (A.multadd) ccy::prolong_add<1, true>(L, X(l), A.lev[l + 1], X(l + 1), A.P3, R);
		else           ccy::prolong_add<0, true>(L, X(l), A.lev[l + 1], X(l + 1), A.P3, R);
		cluster.sync();
		for (int k = 0; k < L.its_up; ++k) {
			ccy::sweep<true>(L, X(l), W(l), R);
			cluster.sync();
			swapped ^= 1u << l;
		}
	}
}
