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
__device__ __forceinline__ double ldg(const double *p) { return __ldcg(p); }

struct Rows { int gw, GW, lane; };   // this warp's index in the cluster, warps in the cluster, lane

// x = scale * (b * dinv)  (first Richardson iteration from a zero guess)
__device__ __forceinline__ void first_sweep(const CLevel &L, double *x, const Rows &R)
{
	for (int i = R.gw; i < L.ni; i += R.GW) {
		const double dinv = L.coef[(size_t)i * MGB_COEF_STRIDE + 5];
		for (int j = R.lane; j < L.pitch; j += 32) {
			const size_t o = (size_t)i * L.pitch + j;
			x[o] = (j < L.nj) ? mul(L.scale, mul(ldg(L.b + o), dinv)) : 0.0;
		}
	}
}
// w = x + scale * ((b - A x) * dinv)
__device__ __forceinline__ void sweep(const CLevel &L, const double *x, double *w, const Rows &R)
{
	const ptrdiff_t P = L.pitch;
	for (int i = R.gw; i < L.ni; i += R.GW) {
		const double *cf = L.coef + (size_t)i * MGB_COEF_STRIDE;
		const double aS = cf[0], aW = cf[1], aC = cf[2], aE = cf[3], aN = cf[4], dinv = cf[5];
		for (int j = R.lane; j < L.pitch; j += 32) {
			const ptrdiff_t o = (ptrdiff_t)i * P + j;
			double out = 0.0;
			if (j < L.nj) {
				const double xc = ldg(x + o);
				const double t = stencil5(aS, aW, aC, aE, aN, ldg(x + o - P), ldg(x + o - 1), xc, ldg(x + o + 1), ldg(x + o + P));
				out = add(xc, mul(L.scale, mul(sub(ldg(L.b + o), t), dinv)));
			}
			w[o] = out;
		}
	}
}
// bc = res * (b - A x)   (k_restrict<1>, natural numbering)
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
						const double t = stencil5(aS, aW, aC, aE, aN, ldg(x + o - P), ldg(x + o - 1), ldg(x + o), ldg(x + o + 1), ldg(x + o + P));
						const double term = mul(Rw.w[a * 3 + b], sub(ldg(F.b + o), t));
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
template <int MULTADD>
__device__ __forceinline__ void prolong_add(const CLevel &F, double *x, const CLevel &C, const double *xc, const Stencil3 &Pw, const Rows &R)
{
	const ptrdiff_t PC = C.pitch;
	for (int i = R.gw; i < F.ni; i += R.GW) {
		for (int j = R.lane; j < F.pitch; j += 32) {
			const size_t o = (size_t)i * F.pitch + j;
			double out = 0.0;
			if (j < F.nj) {
				const double u = ldg(x + o);
				const int J0 = j >> 1, Jm = J0 - 1;
				if (i & 1) {
					const double *c = xc + (ptrdiff_t)((i - 1) >> 1) * PC;
					if (j & 1) {
						const double s0 = mul(Pw.w[3 + 1], ldg(c + J0));
						out = MULTADD ? add(u, s0) : add(u, mul(1.0, s0));
					} else {
						const double cm = mul(Pw.w[3 + 2], ldg(c + Jm)), c0 = mul(Pw.w[3 + 0], ldg(c + J0));
						out = MULTADD ? add(add(u, cm), c0) : add(u, mul(1.0, add(cm, c0)));
					}
				} else {
					const double *cA = xc + (ptrdiff_t)((i >> 1) - 1) * PC;     // row -1: the zero ghost row
					const double *cB = cA + PC;
					if (j & 1) {
						const double sa = mul(Pw.w[6 + 1], ldg(cA + J0)), sb = mul(Pw.w[0 + 1], ldg(cB + J0));
						out = MULTADD ? add(add(u, sa), sb) : add(u, mul(1.0, add(sa, sb)));
					} else {
						const double am = mul(Pw.w[6 + 2], ldg(cA + Jm)), a0 = mul(Pw.w[6 + 0], ldg(cA + J0));
						const double bm = mul(Pw.w[0 + 2], ldg(cB + Jm)), b0 = mul(Pw.w[0 + 0], ldg(cB + J0));
						out = MULTADD ? add(add(add(add(u, am), a0), bm), b0) : add(u, mul(1.0, add(add(add(am, a0), bm), b0)));
					}
				}
			}
			x[o] = out;
		}
	}
}
}  // namespace ccy

__global__ void __cluster_dims__(CC_CTAS, 1, 1) __launch_bounds__(CC_THREADS)
k_coarse_cycle(CoarseArgs A)
{
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	ccy::Rows R;
	R.lane = threadIdx.x & 31;
	R.gw = (int)cluster.block_rank() * (CC_THREADS / 32) + (threadIdx.x >> 5);
	R.GW = CC_CTAS * (CC_THREADS / 32);
	// the ping-pong swaps iterate and scratch after every out-of-place sweep: bit l of `swapped` = level l currently
	// has its iterate in lev[l].w (the host applies the same number of swaps to its own pointers after the launch)
	unsigned swapped = 0u;
	auto X = [&](int l) { return (swapped >> l) & 1u ? A.lev[l].w : A.lev[l].x; };
	auto W = [&](int l) { return (swapped >> l) & 1u ? A.lev[l].x : A.lev[l].w; };
	// ---- down
	for (int l = 0; l < A.nlev; ++l) {
		const CLevel &L = A.lev[l];
		ccy::first_sweep(L, X(l), R);
		cluster.sync();
		for (int k = 1; k < L.its_down; ++k) {
			ccy::sweep(L, X(l), W(l), R);
			cluster.sync();
			swapped ^= 1u << l;
		}
		if (l + 1 < A.nlev) {
			ccy::residual_restrict(L, X(l), A.lev[l + 1], A.R3, R);
			cluster.sync();
		}
	}
	// ---- up
	for (int l = A.nlev - 2; l >= 0; --l) {
		const CLevel &L = A.lev[l];
		if (A.multadd) ccy::prolong_add<1>(L, X(l), A.lev[l + 1], X(l + 1), A.P3, R);
		else           ccy::prolong_add<0>(L, X(l), A.lev[l + 1], X(l + 1), A.P3, R);
		cluster.sync();
		for (int k = 0; k < L.its_up; ++k) {
			ccy::sweep(L, X(l), W(l), R);
			cluster.sync();
			swapped ^= 1u << l;
		}
	}
}
