// mgb_stencil.cuh -- matrix-free 5-point kernels: apply, residual, residual norm, weighted-Jacobi sweep,
// red-black SOR half sweep.  One thread owns two adjacent columns (16-byte loads/stores) and streams down
// ry grid rows keeping a three-row register window, so every x value is fetched from L2/HBM once per block
// (+2 halo rows per RY) and the west/east neighbours come from two extra (L1-resident) scalar loads.
//
// Algorithmic HBM bytes per unknown: apply 16, residual 24, residual norm 16, Jacobi sweep 24,
// red-black half sweep 24 (x is read and written in full, b in full).
#pragma once
#include "mgb_common.cuh"

#define MGB_SB_THREADS 128          // threads per block of the streaming kernels
#define MGB_SB_COLS (2 * MGB_SB_THREADS)

enum { ST_APPLY = 0, ST_RESID = 1, ST_RESNORM = 2, ST_JACOBI = 3, ST_APPLYDOT = 4 };   // APPLYDOT: y = A x and partial sums of x . y (CG)

// x: input vector, b: right-hand side (unused for ST_APPLY), y: output (unused for ST_RESNORM)
// partial: one double per block (ST_RESNORM)
template <int MODE>
__global__ void __launch_bounds__(MGB_SB_THREADS)
k_stream5(const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ y,
          LevelDev L, double scale, double *__restrict__ partial, int ry)
{
	pdl_enter();
	const int j0 = (blockIdx.x * MGB_SB_THREADS + threadIdx.x) * 2;
	const int ibeg = blockIdx.y * ry;
	const int iend = min(ibeg + ry, L.ni);
	double acc = 0.0;
	if (j0 < L.pitch) {
		const size_t P = (size_t)L.pitch;
		const double *xp = x + (size_t)ibeg * P + j0;
		const double *cf = L.coef + (size_t)(L.i0 + ibeg) * MGB_COEF_STRIDE;
		double aS = cf[0], aW = cf[1], aC = cf[2], aE = cf[3], aN = cf[4], dinv = cf[5];
		double2 xm = ld2(xp - P);
		double2 xc = ld2(xp);
		double xw = xp[-1], xe = xp[2];
		const bool in0 = j0 < L.nj, in1 = j0 + 1 < L.nj;
#pragma unroll 4
		for (int i = ibeg; i < iend; ++i) {
			const double2 xn = ld2(xp + P);
			const double xnw = xp[P - 1], xne = xp[P + 2];
			double2 bb = make_double2(0.0, 0.0);
			if (MODE != ST_APPLY && MODE != ST_APPLYDOT) bb = ld2(b + (size_t)i * P + j0);
			if (!L.uniform) {
				aS = cf[0]; aW = cf[1]; aC = cf[2]; aE = cf[3]; aN = cf[4]; dinv = cf[5];
				cf += MGB_COEF_STRIDE;
			}
			// column j0 is even: point (i, j0) is red when the global row is even, (i, j0+1) has the other colour
			const int rowpar = (L.i0 + i) & 1;
			const int o0 = L.rb ? 1 + rowpar : 0, o1 = L.rb ? 2 - rowpar : 0;
			const double t0 = stencil5_ord(o0, aS, aW, aC, aE, aN, xm.x, xw, xc.x, xc.y, xn.x);
			const double t1 = stencil5_ord(o1, aS, aW, aC, aE, aN, xm.y, xc.x, xc.y, xe, xn.y);
			double2 out;
			if (MODE == ST_APPLY || MODE == ST_APPLYDOT) {
				out.x = t0; out.y = t1;
			} else {
				// r = b - A x   (KSPBuildResidual / Richardson: MatMult then VecAYPX(r, -1, b))
				const double r0 = sub(bb.x, t0), r1 = sub(bb.y, t1);
				if (MODE == ST_JACOBI) {
					// z = r * (1/diag)  (PCApply_Jacobi) ; x = x + scale * z  (VecAXPY)
					out.x = add(xc.x, mul(scale, mul(r0, dinv)));
					out.y = add(xc.y, mul(scale, mul(r1, dinv)));
				} else {
					out.x = r0; out.y = r1;
				}
			}
			if (!in0) out.x = 0.0;
			if (!in1) out.y = 0.0;
			if (MODE == ST_RESNORM) acc += out.x * out.x + out.y * out.y;
			else st2(y + (size_t)i * P + j0, out);
			if (MODE == ST_APPLYDOT) acc += xc.x * out.x + xc.y * out.y;
			xm = xc; xc = xn; xw = xnw; xe = xne; xp += P;
		}
	}
	if (MODE == ST_RESNORM || MODE == ST_APPLYDOT) {
		const double s = block_sum<MGB_SB_THREADS>(acc);
		if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
	}
}

// CG direction step in one pass (KSPSolve_CG: VecAYPX(p, b, z) ; MatMult(A, p, w) ; VecDot(p, w), plus the VecAXPY(x, a, p) of
// the PREVIOUS iteration, deferred to here because this pass reads the old p anyway):
//     x += a_prev * p_old ;  p_new = z + b * p_old ;  w = A p_new ;  partial[block] = sum p_new . w
// p_new is written OUT OF PLACE (pn: a block recomputes the rows above and below its chunk from z and p_old, which must not
// have been overwritten by its neighbour) and, on a strip, also into the two ghost rows next to the strip: every rank derives
// them from the ghost rows of z and p_old with the same arithmetic as their owner, so p needs no exchange of its own.
// b = *ratio (beta / betaold) and a_prev = *alpha are device scalars left by the reduction tails.  48 B per unknown for
// what took 24 (AYPX) + 16 (apply) + 24 (x += a p) as separate passes; per value the same operations in the same order.
__global__ void __launch_bounds__(MGB_SB_THREADS)
k_cg_pstep(const double *__restrict__ z, const double *__restrict__ p, double *__restrict__ pn, double *__restrict__ w,
           double *__restrict__ x, LevelDev L, const double *__restrict__ ratio, const double *__restrict__ alpha,
           double *__restrict__ partial, int ry)
{
	pdl_enter();
	const int j0 = (blockIdx.x * MGB_SB_THREADS + threadIdx.x) * 2;
	const int ibeg = blockIdx.y * ry;
	const int iend = min(ibeg + ry, L.ni);
	double acc = 0.0;
	if (j0 < L.pitch) {
		const double b = *ratio, a = *alpha;
		const ptrdiff_t P = (ptrdiff_t)L.pitch;
		ptrdiff_t o = (ptrdiff_t)ibeg * P + j0;
		auto pn1 = [&](ptrdiff_t q) { return add(z[q], mul(b, p[q])); };
		auto pn2 = [&](double2 zz, double2 pp) { return make_double2(add(zz.x, mul(b, pp.x)), add(zz.y, mul(b, pp.y))); };
		const double *cf = L.coef + (size_t)(L.i0 + ibeg) * MGB_COEF_STRIDE;
		double aS = cf[0], aW = cf[1], aC = cf[2], aE = cf[3], aN = cf[4];
		double2 pm = pn2(ld2(z + o - P), ld2(p + o - P));          // row ibeg - 1
		double2 poc = ld2(p + o);
		double2 pc = pn2(ld2(z + o), poc);
		double pw = pn1(o - 1), pe = pn1(o + 2);
		const bool in0 = j0 < L.nj, in1 = j0 + 1 < L.nj;
		if (ibeg == 0 && L.i0 > 0) st2(pn + o - P, pm);            // ghost row above the strip
#pragma unroll 2
		for (int i = ibeg; i < iend; ++i) {
			const double2 pon = ld2(p + o + P);
			const double2 pnx = pn2(ld2(z + o + P), pon);
			const double pnw = pn1(o + P - 1), pne = pn1(o + P + 2);
			if (!L.uniform) { aS = cf[0]; aW = cf[1]; aC = cf[2]; aE = cf[3]; aN = cf[4]; cf += MGB_COEF_STRIDE; }
			const int rowpar = (L.i0 + i) & 1;
			const int o0 = L.rb ? 1 + rowpar : 0, o1 = L.rb ? 2 - rowpar : 0;
			double2 out;
			out.x = stencil5_ord(o0, aS, aW, aC, aE, aN, pm.x, pw, pc.x, pc.y, pnx.x);
			out.y = stencil5_ord(o1, aS, aW, aC, aE, aN, pm.y, pc.x, pc.y, pe, pnx.y);
			if (!in0) out.x = 0.0;
			if (!in1) out.y = 0.0;
			st2(w + o, out);
			st2(pn + o, pc);
			acc += pc.x * out.x + pc.y * out.y;
			double2 xv = ld2(x + o);
			xv.x = add(xv.x, mul(a, poc.x)); xv.y = add(xv.y, mul(a, poc.y));
			st2(x + o, xv);
			pm = pc; pc = pnx; pw = pnw; pe = pne; poc = pon; o += P;
		}
		if (iend == L.ni && L.i0 + L.ni < L.gni) st2(pn + o, pc);  // ghost row below the strip (pc is row ni now)
	}
	const double s = block_sum<MGB_SB_THREADS>(acc);
	if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
}

// First Richardson iteration from a zero initial guess: r = b, z = r * dinv, x = 0 + scale * z
// (KSPSolve_Richardson with guess_zero: no MatMult; ref: src/solver.c:1531-1532,1536).  16 B / unknown.
__global__ void __launch_bounds__(256)
k_jacobi_first(const double *__restrict__ b, double *__restrict__ x, LevelDev L, double scale)
{
	pdl_enter();
	const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
	const int i = blockIdx.y;
	if (j0 >= L.pitch) return;
	const double dinv = L.coef[(size_t)(L.i0 + i) * MGB_COEF_STRIDE + 5];
	const size_t o = (size_t)i * L.pitch + j0;
	const double2 bb = ld2(b + o);
	double2 out;
	out.x = (j0 < L.nj) ? mul(scale, mul(bb.x, dinv)) : 0.0;
	out.y = (j0 + 1 < L.nj) ? mul(scale, mul(bb.y, dinv)) : 0.0;
	st2(x + o, out);
}

// Red-black SOR half sweep, in place, colour c: updates the points with (i_global + j) % 2 == c.
// PETSc MatSOR_SeqAIJ on the red-first numbering (oracle -map 3):
//   sum = b ; sum -= a_k x_k for the four neighbours in ascending column order (S, W, E, N) ;
//   x = (1 - omega) * x + sum * idiag,  idiag = omega / diag (1/diag when omega == 1).
// VARIANT 1 (backward-only sweep with a nonzero guess): the row loop of MatSOR runs over the whole row
// including the diagonal and adds it back:  x = (1-omega) x + (sum + diag*x) * idiag.
// The other colour's values are never written by this launch, so the in-place update is race free; the
// thread writes back its whole 16-byte pair (the partner value unchanged).
template <int VARIANT>
__global__ void __launch_bounds__(MGB_SB_THREADS)
k_rb_half(double *__restrict__ x, const double *__restrict__ b, LevelDev L, int colour, double omega, int ry)
{
	pdl_enter();
	const int j0 = (blockIdx.x * MGB_SB_THREADS + threadIdx.x) * 2;
	const int ibeg = blockIdx.y * ry;
	const int iend = min(ibeg + ry, L.ni);
	if (j0 >= L.pitch) return;
	const size_t P = (size_t)L.pitch;
	double *xp = x + (size_t)ibeg * P + j0;
	const double *cf = L.coef + (size_t)(L.i0 + ibeg) * MGB_COEF_STRIDE;
	double aS = cf[0], aW = cf[1], aE = cf[3], aN = cf[4], idiag = cf[6], mdiag = cf[7];
	const double om1 = sub(1.0, omega);
	double2 xm = ld2(xp - P);
	double2 xc = ld2(xp);
	const bool in0 = j0 < L.nj, in1 = j0 + 1 < L.nj;
#pragma unroll 4
	for (int i = ibeg; i < iend; ++i) {
		const double2 xn = ld2(xp + P);
		const double2 bb = ld2(b + (size_t)i * P + j0);
		if (!L.uniform) {
			aS = cf[0]; aW = cf[1]; aE = cf[3]; aN = cf[4]; idiag = cf[6]; mdiag = cf[7];
			cf += MGB_COEF_STRIDE;
		}
		const int first = ((L.i0 + i + colour) & 1) == 0;     // update element .x (column j0, j0 even)
		double xS, xW, xC, xE, xN, bv;
		if (first) { xS = xm.x; xW = xp[-1]; xC = xc.x; xE = xc.y; xN = xn.x; bv = bb.x; }
		else       { xS = xm.y; xW = xc.x;  xC = xc.y; xE = xp[2]; xN = xn.y; bv = bb.y; }
		double sum = bv;
		double v;
		if (VARIANT == 0) {
			sum = sub(sum, mul(aS, xS));
			sum = sub(sum, mul(aW, xW));
			sum = sub(sum, mul(aE, xE));
			sum = sub(sum, mul(aN, xN));
			v = add(mul(om1, xC), mul(sum, idiag));
		} else {
			// whole row in the red-first numbering: a red row stores diag, S, W, E, N ; a black row S, W, E, N, diag
			if (colour == 0) sum = sub(sum, mul(mdiag, xC));
			sum = sub(sum, mul(aS, xS));
			sum = sub(sum, mul(aW, xW));
			sum = sub(sum, mul(aE, xE));
			sum = sub(sum, mul(aN, xN));
			if (colour == 1) sum = sub(sum, mul(mdiag, xC));
			v = add(mul(om1, xC), mul(add(sum, mul(mdiag, xC)), idiag));
		}
		double2 out = xc;
		if (first) { if (in0) out.x = v; } else { if (in1) out.y = v; }
		st2(xp, out);
		xm = xc; xc = xn; xp += P;
	}
}
