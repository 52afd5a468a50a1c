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
// so the results are bit-identical.  Smoothers: weighted Jacobi (out of place, ping-pong) and red-black SOR on the
// red-first numbering (-map 3; in-place half sweeps, sums in red-first order).  Data stays in global memory (L2-resident: <= 0.5 MB per vector); loads use
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
	unsigned mask_down, mask_up;   // red-black SOR: its_* counts HALF sweeps, bit k = colour of half sweep k (0 red, 1 black)
};
struct CoarseArgs {
	int nlev;                // levels lev[0] (finest of the bottom part) .. lev[nlev-1] (coarsest)
	int multadd;             // 1: PCMG's MatInterpolateAdd order in the correction step
	int rb;                  // 1: red-black SOR on the red-first numbering (-map 3): in-place half sweeps, row sums and transfer
	                         //    sums in red-first order (k_rb_half variant 0, k_restrict<1> / k_prolong_add with L.rb)
	double omega;            // SOR relaxation factor (coef[6] holds omega / diag)
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
// one red-black half sweep of `colour`, in place (k_rb_half variant 0).  zero: the iterate is taken as zero (the first half
// sweep of a zero-guess smoothing; the same arithmetic on zeros) and the points of the other colour are set to zero.
template <bool G>
__device__ __forceinline__ void rb_half(const CLevel &L, double *x, int colour, double om1, bool zero, const Rows &R)
{
	const ptrdiff_t P = L.pitch;
	for (int i = R.gw; i < L.ni; i += R.GW) {
		const double *cf = L.coef + (size_t)i * MGB_COEF_STRIDE;
		const double aS = cf[0], aW = cf[1], aE = cf[3], aN = cf[4], idiag = cf[6];
		for (int j = R.lane; j < L.pitch; j += 32) {
			const ptrdiff_t o = (ptrdiff_t)i * P + j;
			if (j >= L.nj) { if (zero) x[o] = 0.0; continue; }
			if (((i + j + colour) & 1) == 0) {
				double xS = 0.0, xW = 0.0, xC = 0.0, xE = 0.0, xN = 0.0;
				if (!zero) { xS = ld<G>(x + o - P); xW = ld<G>(x + o - 1); xC = ld<G>(x + o); xE = ld<G>(x + o + 1); xN = ld<G>(x + o + P); }
				double sum = ld<G>(L.b + o);
				sum = sub(sum, mul(aS, xS));
				sum = sub(sum, mul(aW, xW));
				sum = sub(sum, mul(aE, xE));
				sum = sub(sum, mul(aN, xN));
				x[o] = add(mul(om1, xC), mul(sum, idiag));
			} else if (zero) x[o] = 0.0;
		}
	}
}
// bc = res * (b - A x)   (k_restrict<1>; RB: row sums and the nine-term sum in red-first order)
template <bool G, bool RB>
__device__ __forceinline__ void residual_restrict(const CLevel &F, const double *x, const CLevel &C, const Stencil3 &Rw, const Rows &R)
{
	const ptrdiff_t P = F.pitch;
	for (int I = R.gw; I < C.ni; I += R.GW) {
		for (int J = R.lane; J < C.pitch; J += 32) {
			double out = 0.0;
			if (J < C.nj) {
				double sum = 0.0;
				double r[9];
#pragma unroll
				for (int a = 0; a < 3; ++a) {
					const int i = 2 * I + a;
					const double *cf = F.coef + (size_t)i * MGB_COEF_STRIDE;
					const double aS = cf[0], aW = cf[1], aC = cf[2], aE = cf[3], aN = cf[4];
#pragma unroll
					for (int b = 0; b < 3; ++b) {
						const ptrdiff_t o = (ptrdiff_t)i * P + (2 * J + b);
						if (RB) {
							// the colour of fine point (2I+a, 2J+b) is (a+b) & 1: a red row sums its diagonal first, a black row last
							const double t = stencil5_ord(1 + ((a + b) & 1), aS, aW, aC, aE, aN, ld<G>(x + o - P), ld<G>(x + o - 1), ld<G>(x + o), ld<G>(x + o + 1), ld<G>(x + o + P));
							r[a * 3 + b] = sub(ld<G>(F.b + o), t);
						} else {
							const double t = stencil5(aS, aW, aC, aE, aN, ld<G>(x + o - P), ld<G>(x + o - 1), ld<G>(x + o), ld<G>(x + o + 1), ld<G>(x + o + P));
							const double term = mul(Rw.w[a * 3 + b], sub(ld<G>(F.b + o), t));
							sum = (a == 0 && b == 0) ? term : add(sum, term);
						}
					}
				}
				if (RB) {
					// the five red fine points, then the four black ones
					sum = mul(Rw.w[0], r[0]);
					sum = add(sum, mul(Rw.w[2], r[2])); sum = add(sum, mul(Rw.w[4], r[4])); sum = add(sum, mul(Rw.w[6], r[6])); sum = add(sum, mul(Rw.w[8], r[8]));
					sum = add(sum, mul(Rw.w[1], r[1])); sum = add(sum, mul(Rw.w[3], r[3])); sum = add(sum, mul(Rw.w[5], r[5])); sum = add(sum, mul(Rw.w[7], r[7]));
				}
				out = sum;
			}
			C.b[(size_t)I * C.pitch + J] = out;
		}
	}
}
// x += pro * xc   (k_prolong_add; one thread per fine point; RB: the terms in ascending red-first coarse number)
template <int MULTADD, bool G, bool RB>
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
						const bool sw = RB && (((((i - 1) >> 1) + Jm) & 1) != 0);   // (I,Jm) black: the red (I,J0) comes first
						const double t0 = sw ? c0 : cm, t1 = sw ? cm : c0;
						out = MULTADD ? add(add(u, t0), t1) : add(u, mul(1.0, add(t0, t1)));
					}
				} else {
					const double *cA = xc + (ptrdiff_t)((i >> 1) - 1) * PC;     // row -1: the zero ghost row
					const double *cB = cA + PC;
					const bool red_am = !RB || (((((i >> 1) - 1) + Jm) & 1) == 0);     // (IA,Jm) and (IB,J0) share a colour
					if (j & 1) {
						const double sa = mul(Pw.w[6 + 1], ld<G>(cA + J0)), sb = mul(Pw.w[0 + 1], ld<G>(cB + J0));
						const double s0 = (RB && red_am) ? sb : sa, s1 = (RB && red_am) ? sa : sb;   // (IA,J0) black, (IB,J0) red
						out = MULTADD ? add(add(u, s0), s1) : add(u, mul(1.0, add(s0, s1)));
					} else {
						const double am = mul(Pw.w[6 + 2], ld<G>(cA + Jm)), a0 = mul(Pw.w[6 + 0], ld<G>(cA + J0));
						const double bm = mul(Pw.w[0 + 2], ld<G>(cB + Jm)), b0 = mul(Pw.w[0 + 0], ld<G>(cB + J0));
						double t0 = am, t1 = a0, t2 = bm, t3 = b0;
						if (RB) { if (red_am) { t0 = am; t1 = b0; t2 = a0; t3 = bm; } else { t0 = a0; t1 = bm; t2 = am; t3 = b0; } }
						out = MULTADD ? add(add(add(add(u, t0), t1), t2), t3) : add(u, mul(1.0, add(add(add(t0, t1), t2), t3)));
					}
				}
			}
			x[o] = out;
		}
	}
}
}  // namespace ccy

// One smoothing call on a level: Jacobi = first sweep from the zero guess (down) + out-of-place sweeps that swap x and w;
// red-black SOR = in-place half sweeps in the colours of `mask`, the first one of a zero-guess call on an implied zero
// iterate.  SYNC() separates the phases; *swapped toggles with every out-of-place sweep.
template <bool G, bool RB, class SYNC>
__device__ __forceinline__ void cc_smooth(const CoarseArgs &A, const CLevel &L, bool zero_guess, int its, unsigned mask, unsigned *swapped, int bit,
                                          const ccy::Rows &R, SYNC sync)
{
	auto X = [&]() { return (*swapped >> bit) & 1u ? L.w : L.x; };
	auto W = [&]() { return (*swapped >> bit) & 1u ? L.x : L.w; };
	if (RB) {
		const double om1 = sub(1.0, A.omega);
		for (int k = 0; k < its; ++k) { ccy::rb_half<G>(L, X(), (int)((mask >> k) & 1u), om1, zero_guess && k == 0, R); sync(); }
		return;
	}
	int k = 0;
	if (zero_guess) { ccy::first_sweep<G>(L, X(), R); sync(); k = 1; }
	for (; k < its; ++k) { ccy::sweep<G>(L, X(), W(), R); sync(); *swapped ^= 1u << bit; }
}

// the whole sub-cycle of levels lo .. nlev-1 inside ONE CTA with everything in shared memory (levels of at most
// CC_TINY rows): __syncthreads between phases instead of cluster barriers and L2 round trips.  (Measured, round 2: taking the
// 63-row level in here as well -- 152 KB of shared memory, one SM instead of eight -- is slower by ~11 us per cycle.)
#define CC_TINY 31
#define CC_TINY_DOUBLES 1792              // per vector, all tiny levels together (ghost row above and below each level)
#define CC_TINY_COEF 512                  // coefficient rows of all tiny levels (8 doubles per grid row)
#define CC_SMEM_BYTES ((3 * CC_TINY_DOUBLES + CC_TINY_COEF) * sizeof(double))     // 46 KB of dynamic shared memory
template <int MULTADD, bool RB>
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
	auto bar = [&]() { __syncthreads(); };
	for (int l = 0; l < nt; ++l) {
		cc_smooth<false, RB>(A, T[l], true, T[l].its_down, T[l].mask_down, &swapped, l, R, bar);
		if (l + 1 < nt) { ccy::residual_restrict<false, RB>(T[l], X(l), T[l + 1], A.R3, R); __syncthreads(); }
	}
	for (int l = nt - 2; l >= 0; --l) {
		ccy::prolong_add<MULTADD, false, RB>(T[l], X(l), T[l + 1], X(l + 1), A.P3, R);
		__syncthreads();
		cc_smooth<false, RB>(A, T[l], false, T[l].its_up, T[l].mask_up, &swapped, l, R, bar);
	}
	// the correction of the first tiny level goes back to global memory, into the buffer the host expects after the
	// same number of swaps; the deeper levels' vectors are scratch and need not be written back
	double *gx = (swapped & 1u) ? A.lev[lo].w : A.lev[lo].x;
	double *sxf = X(0);
	for (int k = threadIdx.x; k < T[0].ni * T[0].pitch; k += blockDim.x) gx[k] = sxf[k];
}

template <bool RB>
__device__ __forceinline__ void coarse_cycle_body(const CoarseArgs &A, double *sx, double *sw, double *sb, double *sc)
{
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
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
	auto bar = [&]() { cluster.sync(); };
	// ---- down (levels 0 .. lt-1 by the whole cluster)
	for (int l = 0; l < lt; ++l) {
		const CLevel &L = A.lev[l];
		cc_smooth<true, RB>(A, L, true, L.its_down, L.mask_down, &swapped, l, R, bar);
		if (l + 1 < A.nlev) {
			ccy::residual_restrict<true, RB>(L, X(l), A.lev[l + 1], A.R3, R);
			cluster.sync();
		}
	}
	// ---- the tiny tail in one CTA's shared memory
	if (lt < A.nlev) {
		if (cluster.block_rank() == 0) {
			if (A.multadd) tiny_cycle<1, RB>(A, lt, sx, sw, sb, sc, R1);
			else           tiny_cycle<0, RB>(A, lt, sx, sw, sb, sc, R1);
		}
		// host-side bookkeeping counts the swaps of every level; the first tiny level's result was stored accordingly
		if (!RB) {
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
		if (A.multadd) ccy::prolong_add<1, true, RB>(L, X(l), A.lev[l + 1], X(l + 1), A.P3, R);
		else           ccy::prolong_add<0, true, RB>(L, X(l), A.lev[l + 1], X(l + 1), A.P3, R);
		cluster.sync();
		cc_smooth<true, RB>(A, L, false, L.its_up, L.mask_up, &swapped, l, R, bar);
	}
}

__global__ void __cluster_dims__(CC_CTAS, 1, 1) __launch_bounds__(CC_THREADS)
k_coarse_cycle(CoarseArgs A)
{
	pdl_enter();
	extern __shared__ __align__(16) double cc_smem[];
	double *sx = cc_smem, *sw = sx + CC_TINY_DOUBLES, *sb = sw + CC_TINY_DOUBLES, *sc = sb + CC_TINY_DOUBLES;
	if (A.rb) coarse_cycle_body<true>(A, sx, sw, sb, sc);
	else      coarse_cycle_body<false>(A, sx, sw, sb, sc);
}
