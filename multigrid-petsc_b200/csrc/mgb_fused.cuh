// mgb_fused.cuh -- the whole per-level work of one V-cycle leg in ONE pass over HBM (temporal blocking).
//
//   down leg:  D weighted-Jacobi sweeps  ->  r = b - A u  ->  b_coarse = res * r          (POST_RESTRICT)
//   up leg:    u += pro * u_coarse       ->  D sweeps     [ ->  sum (b - A u)^2 ]          (PRE_PROLONG*, POST_NORM)
//
// replacing, per level and leg, D x KSPSolve(Richardson+PCJACOBI) + KSPBuildResidual + MatMult(res) resp.
// MatMult(pro) + VecAXPY + D x KSPSolve (+ KSPBuildResidual + VecNorm) of the reference's loop
// (ref: src/solver.c:1531-1546).  Unfused, a V(3,3) cycle moves 264 B per fine unknown; fused, each leg
// reads u and b once and writes u once: 2 x 26 B.
//
// Every value is produced by exactly the same sequence of IEEE operations as in the one-sweep kernels
// (mgb_stencil.cuh / mgb_transfer.cuh), so the results are bit-identical; only the order in which points are
// visited changes.  A block owns a tile of FJ_VALID columns x `rows` rows and streams down the rows.  Each thread
// owns two adjacent columns and keeps, for every stage s = 0..D (stage 0 = input, stage s = after s sweeps), a
// three-row register window; stage s of row t-s is computed at step t from stage s-1 of rows t-s-1 .. t-s+1.
// West/east neighbours of the centre row come from a double-buffered shared-memory copy of the rows produced in
// the previous step (one __syncthreads per step).  FJ_HALO columns on each side of the tile and D+2 rows above
// and below it are recomputed redundantly; values outside the global grid are forced to zero at every stage,
// exactly like the pad columns / ghost rows of the one-sweep kernels.
//
// Strips: rows in the ghost zone that belong to a neighbour are recomputed from ghost data, which therefore has
// to be valid to depth D+2 (u) / D+1 (b) on entry -- MGB_GHOST_ROWS = 6 covers D <= 4.
#pragma once
#include "mgb_common.cuh"
#include "mgb_transfer.cuh"

#define FJ_THREADS 128
#define FJ_COLS (2 * FJ_THREADS)
#define FJ_HALO 6
#define FJ_VALID (FJ_COLS - 2 * FJ_HALO)
#define FJ_PF 2                                   // rows of u / b requested ahead of their use

enum { PRE_GIVEN = 0, PRE_ZERO = 1, PRE_PROLONG = 2, PRE_PROLONG_MULTADD = 3 };
enum { POST_NONE = 0, POST_RESTRICT = 1, POST_NORM = 2 };

struct FusedArgs {
	const double *u_in;      // stage 0 (unused for PRE_ZERO)
	const double *b;
	double *u_out;
	const double *uc;        // coarse correction, element (0,0) of this strip's coarse rows (PRE_PROLONG*)
	double *bc;              // coarse right-hand side out (POST_RESTRICT)
	double *partial;         // one double per block (POST_NORM)
	LevelDev F, C;
	Stencil3 R3, P3;
	double scale;
	int rows;                // rows per block (even)
	int gni;                 // global number of rows of the fine level
};

struct Coef { double aS, aW, aC, aE, aN, dinv; };
__device__ __forceinline__ Coef load_coef(const LevelDev &L, int gni, int grow)
{
	const int g = grow < 0 ? 0 : (grow >= gni ? gni - 1 : grow);
	const double *cf = L.coef + (size_t)g * MGB_COEF_STRIDE;
	Coef c; c.aS = cf[0]; c.aW = cf[1]; c.aC = cf[2]; c.aE = cf[3]; c.aN = cf[4]; c.dinv = cf[5];
	return c;
}

// u + pro * uc at fine row i (local), columns j0, j0+1 -- the arithmetic of k_prolong_add, natural numbering
template <int MULTADD>
__device__ __forceinline__ double2 prolonged(double2 u, const double *__restrict__ uc, int i, int j0, size_t PC, const Stencil3 &Pw)
{
	const int J0 = j0 >> 1, Jm = J0 - 1;
	double e0, e1;
	if (i & 1) {
		const double *c = uc + (ptrdiff_t)((i - 1) >> 1) * (ptrdiff_t)PC;
		const double cm = mul(Pw.w[3 + 2], c[Jm]), c0 = mul(Pw.w[3 + 0], c[J0]);
		const double s0 = mul(Pw.w[3 + 1], c[J0]);
		if (MULTADD) { e0 = add(add(u.x, cm), c0); e1 = add(u.y, s0); }
		else { e0 = add(u.x, mul(1.0, add(cm, c0))); e1 = add(u.y, mul(1.0, s0)); }
	} else {
		const double *cA = uc + (ptrdiff_t)((i >> 1) - 1) * (ptrdiff_t)PC;
		const double *cB = cA + PC;
		const double am = mul(Pw.w[6 + 2], cA[Jm]), a0 = mul(Pw.w[6 + 0], cA[J0]);
		const double bm = mul(Pw.w[0 + 2], cB[Jm]), b0 = mul(Pw.w[0 + 0], cB[J0]);
		const double sa = mul(Pw.w[6 + 1], cA[J0]), sb = mul(Pw.w[0 + 1], cB[J0]);
		if (MULTADD) { e0 = add(add(add(add(u.x, am), a0), bm), b0); e1 = add(add(u.y, sa), sb); }
		else { e0 = add(u.x, mul(1.0, add(add(add(am, a0), bm), b0))); e1 = add(u.y, mul(1.0, add(sa, sb))); }
	}
	return make_double2(e0, e1);
}

template <int D, int PRE, int POST>
__global__ void __launch_bounds__(FJ_THREADS)
k_jfused(FusedArgs A)
{
	// rows produced in the previous step, per stage (0..D) and the residual row (index D+1); 2 pad doubles per side
	__shared__ __align__(16) double sh[2][D + 2][FJ_COLS + 4];
	const LevelDev &F = A.F;
	const int tid = threadIdx.x;
	const int c0 = blockIdx.x * FJ_VALID;                 // first valid column of the tile
	const int j0 = c0 - FJ_HALO + 2 * tid;                // this thread's columns j0, j0+1 (j0 even)
	const int y0 = blockIdx.y * A.rows;
	const int y1 = min(y0 + A.rows, F.ni);
	const size_t P = (size_t)F.pitch;
	const bool in0 = j0 >= 0 && j0 < F.nj, in1 = j0 + 1 >= 0 && j0 + 1 < F.nj;
	const bool ld_ok = j0 >= 0 && j0 < F.pitch;           // the pair may be loaded (pad columns hold zeros)
	const bool st_ok = j0 >= c0 && j0 < c0 + FJ_VALID && j0 < F.pitch;

	Coef cu = load_coef(F, A.gni, F.i0);                  // uniform operator: one coefficient set
	double2 win[D + 1][3];                                // win[s][k]: stage s, rows (newest-2+k); win[s][2] is the newest
#pragma unroll
	for (int s = 0; s <= D; ++s)
#pragma unroll
		for (int k = 0; k < 3; ++k) win[s][k] = make_double2(0.0, 0.0);
	double2 bq[D + 1];                                    // bq[k] = b of row t-1-k
#pragma unroll
	for (int k = 0; k <= D; ++k) bq[k] = make_double2(0.0, 0.0);
	double rw[3][3];                                      // POST_RESTRICT: residual rows (own .x, own .y, east) of rows rho'-2..rho'
#pragma unroll
	for (int k = 0; k < 3; ++k) { rw[k][0] = 0.0; rw[k][1] = 0.0; rw[k][2] = 0.0; }
	double2 res_own = make_double2(0.0, 0.0);             // residual row produced in the previous step (own columns)
	double acc = 0.0;

	const int tb = y0 - D - 1, te = y1 + D + 2;
	// software prefetch rings: rows t+FJ_PF of u and t-1+FJ_PF of b are requested FJ_PF steps ahead
	double2 upf[FJ_PF], bpf[FJ_PF];
	auto row_ok = [&](int i) { const int g = F.i0 + i; return g >= 0 && g < A.gni && i >= -MGB_GHOST_ROWS && i < F.ni + MGB_GHOST_ROWS; };
	auto load_u = [&](int i) -> double2 {
		if (PRE == PRE_ZERO || !ld_ok || !row_ok(i)) return make_double2(0.0, 0.0);
		return ld2(A.u_in + (ptrdiff_t)i * (ptrdiff_t)P + j0);
	};
	auto load_b = [&](int i) -> double2 {
		if (!ld_ok || !row_ok(i)) return make_double2(0.0, 0.0);
		return ld2(A.b + (ptrdiff_t)i * (ptrdiff_t)P + j0);
	};
#pragma unroll
	for (int k = 0; k < FJ_PF; ++k) { upf[k] = load_u(tb + k); bpf[k] = load_b(tb - 1 + k); }

	for (int t = tb; t <= te; ++t) {
		const int par = t & 1;
		double (*shp)[FJ_COLS + 4] = sh[par ^ 1];        // rows of the previous step
		double (*shn)[FJ_COLS + 4] = sh[par];            // rows of this step
		// ---- A: stage 0 of row t, b of row t-1 (from the prefetch ring), next requests
		double2 u0 = upf[0], bnew = bpf[0];
#pragma unroll
		for (int k = 0; k + 1 < FJ_PF; ++k) { upf[k] = upf[k + 1]; bpf[k] = bpf[k + 1]; }
		upf[FJ_PF - 1] = load_u(t + FJ_PF);
		bpf[FJ_PF - 1] = load_b(t - 1 + FJ_PF);
#pragma unroll
		for (int k = D; k > 0; --k) bq[k] = bq[k - 1];
		bq[0] = bnew;
		if (PRE == PRE_PROLONG || PRE == PRE_PROLONG_MULTADD) {
			if (ld_ok && row_ok(t)) {
				u0 = prolonged<PRE == PRE_PROLONG_MULTADD>(u0, A.uc, t, j0, (size_t)A.C.pitch, A.P3);
			}
		}
		if (!in0 || !row_ok(t)) u0.x = 0.0;
		if (!in1 || !row_ok(t)) u0.y = 0.0;
		win[0][0] = win[0][1]; win[0][1] = win[0][2]; win[0][2] = u0;
		// ---- B: stage s of row t-s
#pragma unroll
		for (int s = 1; s <= D; ++s) {
			const int c = t - s;                           // centre row
			const int g = F.i0 + c;
			Coef cf = cu;
			if (!F.uniform) cf = load_coef(F, A.gni, g);
			const double2 bb = bq[s - 1];                  // b of row t-s
			double2 o;
			if (PRE == PRE_ZERO && s == 1) {
				// first Richardson iteration from a zero guess: r = b, x = 0 + scale * (r * dinv)
				o.x = mul(A.scale, mul(bb.x, cf.dinv));
				o.y = mul(A.scale, mul(bb.y, cf.dinv));
			} else {
				const double2 xm = win[s - 1][0], xc = win[s - 1][1], xn = win[s - 1][2];
				const double2 wl = *reinterpret_cast<const double2 *>(&shp[s - 1][2 * tid]);       // columns j0-2, j0-1
				const double2 er = *reinterpret_cast<const double2 *>(&shp[s - 1][2 * tid + 4]);   // columns j0+2, j0+3
				const double t0 = stencil5(cf.aS, cf.aW, cf.aC, cf.aE, cf.aN, xm.x, wl.y, xc.x, xc.y, xn.x);
				const double t1 = stencil5(cf.aS, cf.aW, cf.aC, cf.aE, cf.aN, xm.y, xc.x, xc.y, er.x, xn.y);
				const double r0 = sub(bb.x, t0), r1 = sub(bb.y, t1);
				o.x = add(xc.x, mul(A.scale, mul(r0, cf.dinv)));
				o.y = add(xc.y, mul(A.scale, mul(r1, cf.dinv)));
			}
			const bool rok = g >= 0 && g < A.gni;
			if (!in0 || !rok) o.x = 0.0;
			if (!in1 || !rok) o.y = 0.0;
			win[s][0] = win[s][1]; win[s][1] = win[s][2]; win[s][2] = o;
		}
		// ---- F: the finished row t-D
		{
			const int c = t - D;
			if (st_ok && c >= y0 && c < y1) st2(A.u_out + (size_t)c * P + j0, win[D][2]);
		}
		// ---- C: residual of row rho = t-D-1 from stage D
		double2 res = make_double2(0.0, 0.0);
		if (POST != POST_NONE) {
			const int rho = t - D - 1;
			const int g = F.i0 + rho;
			Coef cf = cu;
			if (!F.uniform) cf = load_coef(F, A.gni, g);
			const double2 xm = win[D][0], xc = win[D][1], xn = win[D][2];
			const double2 wl = *reinterpret_cast<const double2 *>(&shp[D][2 * tid]);
			const double2 er = *reinterpret_cast<const double2 *>(&shp[D][2 * tid + 4]);
			const double2 bb = bq[D];
			const double t0 = stencil5(cf.aS, cf.aW, cf.aC, cf.aE, cf.aN, xm.x, wl.y, xc.x, xc.y, xn.x);
			const double t1 = stencil5(cf.aS, cf.aW, cf.aC, cf.aE, cf.aN, xm.y, xc.x, xc.y, er.x, xn.y);
			res.x = sub(bb.x, t0); res.y = sub(bb.y, t1);
			const bool rok = g >= 0 && g < A.gni;
			if (!in0 || !rok) res.x = 0.0;
			if (!in1 || !rok) res.y = 0.0;
			if (POST == POST_NORM) {
				if (st_ok && rho >= y0 && rho < y1) acc += res.x * res.x + res.y * res.y;
			}
		}
		// ---- E: restriction of the residual rows completed in the previous step
		if (POST == POST_RESTRICT) {
			const int rp = t - D - 2;                      // residual row whose east neighbour is now visible
			const double east = shp[D + 1][2 * tid + 4];   // column j0+2 of row rp
			rw[0][0] = rw[1][0]; rw[0][1] = rw[1][1]; rw[0][2] = rw[1][2];
			rw[1][0] = rw[2][0]; rw[1][1] = rw[2][1]; rw[1][2] = rw[2][2];
			rw[2][0] = res_own.x; rw[2][1] = res_own.y; rw[2][2] = east;
			if ((rp & 1) == 0) {
				const int I = (rp >> 1) - 1;               // coarse row (local) fed by fine rows rp-2 .. rp
				const int J = j0 >> 1;
				if (st_ok && I >= (y0 >> 1) && I < (y1 >> 1) && I < A.C.ni && J < A.C.pitch) {
					double sum = mul(A.R3.w[0], rw[0][0]);
					sum = add(sum, mul(A.R3.w[1], rw[0][1]));
					sum = add(sum, mul(A.R3.w[2], rw[0][2]));
					sum = add(sum, mul(A.R3.w[3], rw[1][0]));
					sum = add(sum, mul(A.R3.w[4], rw[1][1]));
					sum = add(sum, mul(A.R3.w[5], rw[1][2]));
					sum = add(sum, mul(A.R3.w[6], rw[2][0]));
					sum = add(sum, mul(A.R3.w[7], rw[2][1]));
					sum = add(sum, mul(A.R3.w[8], rw[2][2]));
					A.bc[(size_t)I * A.C.pitch + J] = (J < A.C.nj) ? sum : 0.0;
				}
			}
			res_own = res;
		}
		// ---- D: publish the rows produced in this step
#pragma unroll
		for (int s = 0; s <= D; ++s) *reinterpret_cast<double2 *>(&shn[s][2 * tid + 2]) = win[s][2];
		if (POST != POST_NONE) *reinterpret_cast<double2 *>(&shn[D + 1][2 * tid + 2]) = res;
		__syncthreads();
	}
	if (POST == POST_NORM) {
		const double s = block_sum<FJ_THREADS>(acc);
		if (tid == 0) A.partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
	}
}
