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
// owns two adjacent columns and keeps, for every stage s = 0..D (stage 0 = input, stage s = after s sweeps), the
// last three rows in registers; stage s of row t-s is computed at step t from stage s-1 of rows t-s-1 .. t-s+1.
// West/east neighbours of the centre row come from a double-buffered shared-memory copy of the rows produced in
// the previous step (one __syncthreads per step).  FJ_HALO columns on each side of the tile and D+2 rows above
// and below it are recomputed redundantly; values outside the global grid are forced to zero at every stage,
// exactly like the pad columns / ghost rows of the one-sweep kernels.
//
// Instruction economy (the kernel is issue / fp64-pipe bound, not HBM bound): all row histories are rings of FOUR
// slots indexed by (row & 3) and the row loop is unrolled by four with the step number modulo 4 as a template
// parameter, so ring positions are compile-time register names -- no register-to-register moves; blocks whose
// tile and halo lie strictly inside the grid run a variant without the out-of-grid masks.
//
// Strips: rows in the ghost zone that belong to a neighbour are recomputed from ghost data, which therefore has
// to be valid to depth D+2 (u) / D+1 (b) on entry -- MGB_GHOST_ROWS = 6 covers D <= 3 (+ the unroll round-down).
#pragma once
#include "mgb_common.cuh"
#include "mgb_transfer.cuh"
#include "mgb_halo.cuh"

#define FJ_THREADS 128
#define FJ_COLS (2 * FJ_THREADS)
#define FJ_HALO 6
#define FJ_VALID (FJ_COLS - 2 * FJ_HALO)
#define FJ_MAXD 3
#define FJ_NR 8                                   // rows of the cp.async input rings (u and b)
#define FJ_PF 6                                   // rows requested ahead of their use
#define FJ_PUB (FJ_COLS + 8)                      // doubles per published row: X[-1..128] then Y[-1..128] (split so that
                                                  // neighbour reads are conflict-free 8-byte accesses)
#define FJ_X(tid) ((tid) + 1)                     // left column (j0) of thread tid
#define FJ_Y(tid) ((tid) + 1 + FJ_THREADS + 4)    // right column (j0+1) of thread tid
template <int D> constexpr size_t jf_smem_bytes() { return sizeof(double) * ((size_t)2 * (D + 2) * FJ_PUB + (size_t)2 * FJ_NR * FJ_COLS + FJ_NR); }

// 16-byte asynchronous copy global -> shared (LDGSTS), bypassing L1: the input rows are consumed exactly once
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
	const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TMA bulk copy global -> shared (UBLKCP), completion counted in bytes on an mbarrier.  One thread moves a whole 2 KB
// input row: no per-thread LDGSTS, and the data is written to shared memory line by line instead of sector by sector
// (ncu: the misaligned per-thread LDGSTS cost 14 shared-memory wavefronts per instruction instead of 4).
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"W_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@!p bra W_%=;\n"
		"}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned smem, const void *gmem, unsigned bytes, unsigned bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem), "l"(gmem), "r"(bytes), "r"(bar) : "memory");
}

enum { PRE_GIVEN = 0, PRE_ZERO = 1, PRE_PROLONG = 2, PRE_PROLONG_MULTADD = 3 };
enum { POST_NONE = 0, POST_RESTRICT = 1, POST_NORM = 2, POST_DOT = 3 };   // POST_DOT: sum u_out . b over the finished rows (the z'r of CG)

// Strip-to-strip traffic done BY the fused kernel (no separate exchange launch; protocol notes in mgb_halo.cuh):
//   push   finished rows of u_out / of the coarse right-hand side that lie in a push range are also stored into a peer's
//          HBM (its ghost rows, or rank 0's whole level for the gather / every rank's staging rows for the broadcast);
//   signal the last of the pushing blocks (ticket) fences and raises the version flags in the destinations' memory;
//   wait   blocks that read ghost rows (the first / last row chunks) or a gathered / broadcast level (all blocks) spin on
//          their own flag words until the peers' pushes of the current version have landed; other blocks never wait.
#define FJ_MAXPUSH 8
#define FJ_MAXWAIT 8
struct PushEnt { double *dst; int lo, hi; };   // rows [lo, hi) of the output -> dst[row * pitch + col] (dst may be null: no neighbour)
struct FusedComm {
	int npu, npb;                               // push ranges of u_out (fine rows) / bc (coarse rows)
	PushEnt pu[FJ_MAXPUSH], pb[2];
	int bc_remote;                              // 1: A.bc itself points into a peer's HBM (gather): every block that writes bc signals
	int nch;                                    // channels signalled by this kernel (0..2)
	unsigned long long *ver[2];                 // my version counter of each channel
	int nsig[2];
	unsigned long long *sig[2][FJ_MAXPUSH];     // flag words in the destinations' memory
	unsigned int *ticket;                       // zero between launches
	int npushblocks;                            // blocks of this launch that take a ticket
	int nw_top, nw_bot, nw_all;                 // waits of the blocks reading ghost rows above / below, and of every block
	const unsigned long long *w_flag[3][FJ_MAXWAIT];
	const unsigned long long *w_ver[3][FJ_MAXWAIT];
	int *status, *status_host;                  // != 0 after a timed-out wait (device word; mapped host mirror)
	long long spin_limit;
};

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
	int rbmask;              // red-black stages (SMK = 1): bit s-1 = colour updated by stage s (0 red = (i+j) even, 1 black); scale = omega
	FusedComm X;             // all zero on a single strip
};

struct Coef { double aS, aW, aC, aE, aN, dinv, nS; };   // nS = -aS (power-of-two path)
template <int SMK = 0>
__device__ __forceinline__ Coef load_coef(const LevelDev &L, int gni, int grow)
{
	const int g = grow < 0 ? 0 : (grow >= gni ? gni - 1 : grow);
	const double *cf = L.coef + (size_t)g * MGB_COEF_STRIDE;
	Coef c; c.aS = cf[0]; c.aW = cf[1]; c.aC = cf[2]; c.aE = cf[3]; c.aN = cf[4]; c.dinv = cf[SMK ? 6 : 5]; c.nS = -cf[0];   // red-black: idiag = omega / d in the place of 1 / d
	return c;
}

// u + pro * uc at fine columns j0, j0+1 -- the arithmetic of k_prolong_add, natural numbering.
// odd fine row: cm / c0 = coarse row (i-1)/2, columns J0-1 / J0.  even fine row: (am, a0) = coarse row i/2-1, (bm, b0) = row i/2.
// TP2: every weight is a power of two (the reference's stencils: 1/4, 1/2, 1): the products are exact, so
// add(x, mul(w, c)) == fma(w, c, x) bit for bit, and mul(1.0, x) == x.
template <int MULTADD, bool TP2>
__device__ __forceinline__ double2 prolonged_odd(double2 u, double cmv, double c0v, const Stencil3 &Pw)
{
	double e0, e1;
	if (TP2) {
		if (MULTADD) { e0 = fma_rn(Pw.w[3 + 0], c0v, fma_rn(Pw.w[3 + 2], cmv, u.x)); e1 = fma_rn(Pw.w[3 + 1], c0v, u.y); }
		else { e0 = add(u.x, fma_rn(Pw.w[3 + 0], c0v, mul(Pw.w[3 + 2], cmv))); e1 = fma_rn(Pw.w[3 + 1], c0v, u.y); }
		return make_double2(e0, e1);
	}
	const double cm = mul(Pw.w[3 + 2], cmv), c0 = mul(Pw.w[3 + 0], c0v);
	const double s0 = mul(Pw.w[3 + 1], c0v);
	if (MULTADD) { e0 = add(add(u.x, cm), c0); e1 = add(u.y, s0); }
	else { e0 = add(u.x, mul(1.0, add(cm, c0))); e1 = add(u.y, mul(1.0, s0)); }
	return make_double2(e0, e1);
}
template <int MULTADD, bool TP2>
__device__ __forceinline__ double2 prolonged_even(double2 u, double amv, double a0v, double bmv, double b0v, const Stencil3 &Pw)
{
	double e0, e1;
	if (TP2) {
		if (MULTADD) {
			e0 = fma_rn(Pw.w[0 + 0], b0v, fma_rn(Pw.w[0 + 2], bmv, fma_rn(Pw.w[6 + 0], a0v, fma_rn(Pw.w[6 + 2], amv, u.x))));
			e1 = fma_rn(Pw.w[0 + 1], b0v, fma_rn(Pw.w[6 + 1], a0v, u.y));
		} else {
			e0 = add(u.x, fma_rn(Pw.w[0 + 0], b0v, fma_rn(Pw.w[0 + 2], bmv, fma_rn(Pw.w[6 + 0], a0v, mul(Pw.w[6 + 2], amv)))));
			e1 = add(u.y, fma_rn(Pw.w[0 + 1], b0v, mul(Pw.w[6 + 1], a0v)));
		}
		return make_double2(e0, e1);
	}
	const double am = mul(Pw.w[6 + 2], amv), a0 = mul(Pw.w[6 + 0], a0v);
	const double bm = mul(Pw.w[0 + 2], bmv), b0 = mul(Pw.w[0 + 0], b0v);
	const double sa = mul(Pw.w[6 + 1], a0v), sb = mul(Pw.w[0 + 1], b0v);
	if (MULTADD) { e0 = add(add(add(add(u.x, am), a0), bm), b0); e1 = add(add(u.y, sa), sb); }
	else { e0 = add(u.x, mul(1.0, add(add(add(am, a0), bm), b0))); e1 = add(u.y, mul(1.0, add(sa, sb))); }
	return make_double2(e0, e1);
}

// ---- red-black numbering (-map 3): the same transfers with the row sums in ascending RED-FIRST column order, i.e. the
// arithmetic of k_prolong_add / k_restrict / stencil5_ord with L.rb set.  pm = colour parity of coarse point (I, J0-1):
// 0 red.  (General arithmetic, no fma: the red-black path is not the headline.)
template <int MULTADD>
__device__ __forceinline__ double2 prolonged_odd_rb(double2 u, double cmv, double c0v, const Stencil3 &Pw, int pm)
{
	const double cm = mul(Pw.w[3 + 2], cmv), c0 = mul(Pw.w[3 + 0], c0v);
	const double s0 = mul(Pw.w[3 + 1], c0v);
	const double t0 = pm ? c0 : cm, t1 = pm ? cm : c0;              // (I, J0-1) black: the red (I, J0) comes first
	double e0, e1;
	if (MULTADD) { e0 = add(add(u.x, t0), t1); e1 = add(u.y, s0); }
	else { e0 = add(u.x, mul(1.0, add(t0, t1))); e1 = add(u.y, mul(1.0, s0)); }
	return make_double2(e0, e1);
}
// pa = colour parity of coarse point (IA, J0-1)
template <int MULTADD>
__device__ __forceinline__ double2 prolonged_even_rb(double2 u, double amv, double a0v, double bmv, double b0v, const Stencil3 &Pw, int pa)
{
	const double am = mul(Pw.w[6 + 2], amv), a0 = mul(Pw.w[6 + 0], a0v);
	const double bm = mul(Pw.w[0 + 2], bmv), b0 = mul(Pw.w[0 + 0], b0v);
	const double sa = mul(Pw.w[6 + 1], a0v), sb = mul(Pw.w[0 + 1], b0v);
	double t0, t1, t2, t3, s0, s1;
	if (pa == 0) { t0 = am; t1 = b0; t2 = a0; t3 = bm; s0 = sb; s1 = sa; }   // (IA,J0-1) and (IB,J0) red
	else         { t0 = a0; t1 = bm; t2 = am; t3 = b0; s0 = sa; s1 = sb; }
	double e0, e1;
	if (MULTADD) { e0 = add(add(add(add(u.x, t0), t1), t2), t3); e1 = add(add(u.y, s0), s1); }
	else { e0 = add(u.x, mul(1.0, add(add(add(t0, t1), t2), t3))); e1 = add(u.y, mul(1.0, add(s0, s1))); }
	return make_double2(e0, e1);
}

// per-thread state: rings of four slots, slot = row & 3
template <int D>
struct JfState {
	double2 win[D + 1][4];   // win[s][row & 3]: stage s of that row (rows t-s-2 .. t-s live)
	double2 bq[4];           // b of rows t-4 .. t-1
	double rw[4][3];         // POST_RESTRICT: residual row (own .x, own .y, east neighbour)
	double wer[2];           // POST_*: west / east neighbours (stage D) of the row whose residual is formed at the NEXT step
	double2 cq[3];           // PRE_PROLONG*: coarse rows T-1, T, T+1 (columns J0-1, J0) of the current group of four fine rows
	double2 cn[2];           //               coarse rows T+2, T+3 requested for the next group
	double acc;
};

struct JfBlock {
	int tid, c0, j0, y0, y1;
	ptrdiff_t P;
	bool in0, in1, ld_ok, st_ok;
	bool tma;                // input rows arrive by TMA bulk copies (pitch >= FJ_COLS) instead of per-thread cp.async
	int rbase;               // first row of the input rings: row i lives in slot (i - rbase) & (FJ_NR - 1)
	unsigned long long *bar; // FJ_NR mbarriers, one per ring slot (TMA path)
	unsigned s_bar, s_u, s_b; // 32-bit shared-window addresses of the mbarriers and of the two input rings
	const double *g_u, *g_b; // column c0 - FJ_HALO of row 0 of u_in / b
	Coef cu;                 // coefficients of a uniform operator, held in ordinary (per-thread) registers
	double scale;
	double sd;               // scale * dinv (power-of-two operator: exact, see jf_point)
};

// One Jacobi update / residual at a point.  OP = 0: per-row coefficients, 1: one coefficient set, 2: one set with
// aS = aW = aE = aN = c = 2^m and aC = -4c (every uniform n = 2^k - 1 grid of the reference: c = 1/h^2 = 4^k).
// For OP = 2 the products c * x and r * dinv are exact (power-of-two scalings commute with rounding), so
//     fl(fl(fl(fl(c xS + c xW) - 4c xC) + c xE) + c xN)  ==  c * fl(fl(fl(fl(xS + xW) - 4 xC) + xE) + xN)
// bit for bit (barring overflow / underflow, which these magnitudes never reach): 9 fp64 instructions per point
// instead of 13, same result.
template <int OP>
__device__ __forceinline__ double jf_residual(const Coef &cf, double b, double xS, double xW, double xC, double xE, double xN)
{
	if (OP == 2) {
		// -4 xC and c t are exact products: add(s, mul(-4, xC)) == fma(-4, xC, s) and sub(b, mul(c, t)) == fma(-c, t, b), same bits
		const double t = add(add(fma_rn(-4.0, xC, add(xS, xW)), xE), xN);
		return fma_rn(cf.nS, t, b);
	}
	return sub(b, stencil5(cf.aS, cf.aW, cf.aC, cf.aE, cf.aN, xS, xW, xC, xE, xN));
}
template <int OP>
__device__ __forceinline__ double jf_update(const Coef &cf, double scale, double sd, double xC, double r)
{
	if (OP == 2) return add(xC, mul(sd, r));              // scale * (r * dinv) == (scale * dinv) * r: r * dinv is exact
	return add(xC, mul(scale, mul(r, cf.dinv)));
}

// b - A x with the row sum in red-first order (stencil5_ord): order 1 = red row (diagonal first), 2 = black row (diagonal last)
template <int OP>
__device__ __forceinline__ double jf_residual_rb(const Coef &cf, int order, double b, double xS, double xW, double xC, double xE, double xN)
{
	if (OP == 2) {
		// c * fl(...) with the same association as stencil5_ord; -4 xC and c t are exact products (see jf_residual)
		double t;
		if (order == 1) t = add(add(add(fma_rn(-4.0, xC, xS), xW), xE), xN);
		else            t = fma_rn(-4.0, xC, add(add(add(xS, xW), xE), xN));
		return fma_rn(cf.nS, t, b);
	}
	return sub(b, stencil5_ord(order, cf.aS, cf.aW, cf.aC, cf.aE, cf.aN, xS, xW, xC, xE, xN));
}
// one point of a red-black half sweep (MatSOR on the red-first numbering, k_rb_half variant 0): cf.dinv holds idiag = omega / d
template <int OP>
__device__ __forceinline__ double jf_rb_update(const Coef &cf, double om1, double b, double xS, double xW, double xC, double xE, double xN)
{
	double sum = b;
	if (OP == 2) {
		sum = fma_rn(cf.nS, xS, sum); sum = fma_rn(cf.nS, xW, sum); sum = fma_rn(cf.nS, xE, sum); sum = fma_rn(cf.nS, xN, sum);
	} else {
		sum = sub(sum, mul(cf.aS, xS)); sum = sub(sum, mul(cf.aW, xW)); sum = sub(sum, mul(cf.aE, xE)); sum = sub(sum, mul(cf.aN, xN));
	}
	return add(mul(om1, xC), mul(sum, cf.dinv));
}

// The stencil coefficients are warp-uniform; left to itself the compiler parks them in uniform registers and copies
// them into vector registers in front of every DMUL/DADD (fp64 instructions take no uniform operands): ~45 extra
// instructions per row step.  Passing them through an opaque asm makes them ordinary per-thread values.
__device__ __forceinline__ double vreg(double x) { double y; asm volatile("mov.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

// one row step; K = t & 3 (compile time), MASK = the block touches the outside of the grid
// request rows `i` of u and b into the input rings (one commit group per call, possibly empty)
template <int PRE, bool MASK, int KW>
__device__ __forceinline__ void jf_request(const FusedArgs &A, const JfBlock &B, double (*in_u)[FJ_COLS], double (*in_b)[FJ_COLS], int i)
{
	const LevelDev &F = A.F;
	const int slot = (i - B.rbase) & (FJ_NR - 1);
	if (!MASK || B.tma) {
		// one thread, one 2 KB bulk copy per array.  Rows outside the arrays are clamped (their values are masked to zero at
		// stage 0 / at every stage of an out-of-grid row); columns outside the grid read the neighbouring rows' memory, which
		// lies inside the allocation whenever pitch >= FJ_COLS, and are masked the same way.  The issuing warp rotates with
		// the step (KW = t & 3): the four warps of a block sit on the four schedulers of the SM, so no scheduler carries
		// the request instructions of every resident block.
		if (B.tid == KW * 32) {
			int ic = i;
			if (MASK) ic = ic < -MGB_GHOST_ROWS ? -MGB_GHOST_ROWS : (ic > F.ni + MGB_GHOST_ROWS - 1 ? F.ni + MGB_GHOST_ROWS - 1 : ic);
			const ptrdiff_t o = (ptrdiff_t)ic * B.P;
			const unsigned bar = B.s_bar + slot * 8u, so = (unsigned)slot * (unsigned)(FJ_COLS * sizeof(double));
			mbar_expect_tx(bar, (PRE != PRE_ZERO ? 2u : 1u) * (unsigned)(FJ_COLS * sizeof(double)));
			if (PRE != PRE_ZERO) bulk_g2s(B.s_u + so, B.g_u + o, FJ_COLS * sizeof(double), bar);
			bulk_g2s(B.s_b + so, B.g_b + o, FJ_COLS * sizeof(double), bar);
		}
		return;
	}
	bool ok = true;
	{
		const int g = F.i0 + i;
		ok = B.ld_ok && g >= 0 && g < A.gni && i >= -MGB_GHOST_ROWS && i < F.ni + MGB_GHOST_ROWS;
	}
	if (ok) {
		if (PRE != PRE_ZERO) cp_async16(&in_u[slot][2 * B.tid], A.u_in + (ptrdiff_t)i * B.P + B.j0);
		cp_async16(&in_b[slot][2 * B.tid], A.b + (ptrdiff_t)i * B.P + B.j0);
	} else {
		if (PRE != PRE_ZERO) *reinterpret_cast<double2 *>(&in_u[slot][2 * B.tid]) = make_double2(0.0, 0.0);
		*reinterpret_cast<double2 *>(&in_b[slot][2 * B.tid]) = make_double2(0.0, 0.0);
	}
	cp_async_commit();
}
// row i of the input rings has arrived (TMA: phase parity of its slot's mbarrier; cp.async: at most FJ_PF younger groups pending)
template <bool MASK>
__device__ __forceinline__ void jf_arrived(const JfBlock &B, int i)
{
	if (!MASK || B.tma) {
		const int k = i - B.rbase;
		mbar_wait(B.s_bar + 8u * (unsigned)(k & (FJ_NR - 1)), (unsigned)(k >> 3) & 1u);
	} else cp_async_wait<FJ_PF>();
}

template <int D, int PRE, int POST, bool MASK, int OP, int K, int SMK>
__device__ __forceinline__ void jf_step(const FusedArgs &A, const JfBlock &B, JfState<D> &S, double (*sh)[D + 2][FJ_PUB],
                                        double (*in_u)[FJ_COLS], double (*in_b)[FJ_COLS], int t)
{
	const LevelDev &F = A.F;
	constexpr int par = K & 1;
	double (*shp)[FJ_PUB] = sh[par ^ 1];                 // rows published in the previous step
	double (*shn)[FJ_PUB] = sh[par];
	const int tid = B.tid;
	auto row_ok = [&](int i) {
		if (!MASK) return true;
		const int g = F.i0 + i;
		return g >= 0 && g < A.gni && i >= -MGB_GHOST_ROWS && i < F.ni + MGB_GHOST_ROWS;
	};
	// ---- rows t+FJ_PF are requested; stage 0 of row t and b of row t-1 have arrived in the input rings
	jf_request<PRE, MASK, K>(A, B, in_u, in_b, t + FJ_PF);
	jf_arrived<MASK>(B, t);
	double2 u0 = make_double2(0.0, 0.0);
	if (PRE != PRE_ZERO) u0 = *reinterpret_cast<const double2 *>(&in_u[(t - B.rbase) & (FJ_NR - 1)][2 * tid]);
	const double2 bold = S.bq[(K + 3) & 3];               // b of row t-5, evicted now (the residual of a D = 3 leg still needs it)
	S.bq[(K + 3) & 3] = *reinterpret_cast<const double2 *>(&in_b[(t - 1 - B.rbase) & (FJ_NR - 1)][2 * tid]);   // slot of row t-1
	if (PRE == PRE_PROLONG || PRE == PRE_PROLONG_MULTADD) {
		// t = 4m + K: coarse rows T-1, T, T+1 with T = 2m are in cq[0..2] (fine rows t..t+3 of the group need exactly these)
		if (SMK == 0) {
			if (K == 0) u0 = prolonged_even<PRE == PRE_PROLONG_MULTADD, OP == 2>(u0, S.cq[0].x, S.cq[0].y, S.cq[1].x, S.cq[1].y, A.P3);
			if (K == 1) u0 = prolonged_odd<PRE == PRE_PROLONG_MULTADD, OP == 2>(u0, S.cq[1].x, S.cq[1].y, A.P3);
			if (K == 2) u0 = prolonged_even<PRE == PRE_PROLONG_MULTADD, OP == 2>(u0, S.cq[1].x, S.cq[1].y, S.cq[2].x, S.cq[2].y, A.P3);
			if (K == 3) u0 = prolonged_odd<PRE == PRE_PROLONG_MULTADD, OP == 2>(u0, S.cq[2].x, S.cq[2].y, A.P3);
		} else {
			// colour parity of coarse point (I, J0-1) for the first coarse row the fine row touches: I = T-1, T, T, T+1 for K = 0..3
			const int T = (t - K) >> 1;
			const int Jm = (B.j0 >> 1) - 1;
			const int pc = (A.C.i0 + T + (K == 0 ? -1 : (K == 3 ? 1 : 0)) + Jm) & 1;
			if (K == 0) u0 = prolonged_even_rb<PRE == PRE_PROLONG_MULTADD>(u0, S.cq[0].x, S.cq[0].y, S.cq[1].x, S.cq[1].y, A.P3, pc);
			if (K == 1) u0 = prolonged_odd_rb<PRE == PRE_PROLONG_MULTADD>(u0, S.cq[1].x, S.cq[1].y, A.P3, pc);
			if (K == 2) u0 = prolonged_even_rb<PRE == PRE_PROLONG_MULTADD>(u0, S.cq[1].x, S.cq[1].y, S.cq[2].x, S.cq[2].y, A.P3, pc);
			if (K == 3) u0 = prolonged_odd_rb<PRE == PRE_PROLONG_MULTADD>(u0, S.cq[2].x, S.cq[2].y, A.P3, pc);
		}
	}
	if (MASK) {
		const bool rk = row_ok(t);
		if (!B.in0 || !rk) u0.x = 0.0;
		if (!B.in1 || !rk) u0.y = 0.0;
	}
	if (PRE == PRE_ZERO) u0 = make_double2(0.0, 0.0);
	S.win[0][K] = u0;
	// ---- stage s of row t-s
#pragma unroll
	for (int s = 1; s <= D; ++s) {
		const int c = t - s;
		const int g = F.i0 + c;
		Coef cf = B.cu;
		if (OP == 0) cf = load_coef<SMK>(F, A.gni, g);
		const double2 bb = S.bq[(K - s) & 3];             // b of row t-s
		double2 o;
		if (SMK == 1) {
			// half sweep of one colour: the points of that colour take the MatSOR update from their four neighbours (all of
			// the other colour, i.e. unchanged by this stage), the others pass through.  .x is column j0 (even): red iff g even
			const double2 xm = S.win[s - 1][(K - s - 1) & 3], xc = S.win[s - 1][(K - s) & 3], xn = S.win[s - 1][(K - s + 1) & 3];
			const int colour = (A.rbmask >> (s - 1)) & 1;
			o = xc;
			if (((g + colour) & 1) == 0) {
				const double xw = shp[s - 1][FJ_Y(tid - 1)];
				o.x = jf_rb_update<OP>(cf, B.scale, bb.x, xm.x, xw, xc.x, xc.y, xn.x);
			} else {
				const double xe = shp[s - 1][FJ_X(tid + 1)];
				o.y = jf_rb_update<OP>(cf, B.scale, bb.y, xm.y, xc.x, xc.y, xe, xn.y);
			}
		} else if (PRE == PRE_ZERO && s == 1) {
			// first Richardson iteration from a zero guess: r = b, x = 0 + scale * (r * dinv)
			o.x = (OP == 2) ? mul(B.sd, bb.x) : mul(B.scale, mul(bb.x, cf.dinv));
			o.y = (OP == 2) ? mul(B.sd, bb.y) : mul(B.scale, mul(bb.y, cf.dinv));
		} else {
			const double2 xm = S.win[s - 1][(K - s - 1) & 3], xc = S.win[s - 1][(K - s) & 3], xn = S.win[s - 1][(K - s + 1) & 3];
			const double xw = shp[s - 1][FJ_Y(tid - 1)];  // column j0-1
			const double xe = shp[s - 1][FJ_X(tid + 1)];  // column j0+2
			const double r0 = jf_residual<OP>(cf, bb.x, xm.x, xw, xc.x, xc.y, xn.x);
			const double r1 = jf_residual<OP>(cf, bb.y, xm.y, xc.x, xc.y, xe, xn.y);
			o.x = jf_update<OP>(cf, B.scale, B.sd, xc.x, r0);
			o.y = jf_update<OP>(cf, B.scale, B.sd, xc.y, r1);
		}
		if (MASK) {
			const bool rok = g >= 0 && g < A.gni;
			if (!B.in0 || !rok) o.x = 0.0;
			if (!B.in1 || !rok) o.y = 0.0;
		}
		S.win[s][(K - s) & 3] = o;
	}
	// ---- the finished row t-D
	{
		const int c = t - D;
		if (D > 0 && B.st_ok && c >= B.y0 && c < B.y1) {                    // D = 0: u is unchanged
			st2(A.u_out + (ptrdiff_t)c * B.P + B.j0, S.win[D][(K - D) & 3]);
			if (POST == POST_DOT) {                                              // b of row t-D is still in the ring (D <= 3)
				const double2 o = S.win[D][(K - D) & 3], bb = S.bq[(K - D) & 3];
				S.acc = fma_rn(o.y, bb.y, fma_rn(o.x, bb.x, S.acc));
			}
		}
	}
	// ---- residual of row rho = t-D-2 from stage D.  It lags the stages by one more row so that all its inputs (rows
	// rho-1 .. rho+1 of stage D and the neighbours of row rho, fetched during the previous step) predate this step:
	// the residual is off the dependent chain stage 1 -> ... -> stage D of the step.
	double2 res = make_double2(0.0, 0.0);
	if (POST == POST_RESTRICT || POST == POST_NORM) {
		const int rho = t - D - 2;
		const int g = F.i0 + rho;
		Coef cf = B.cu;
		if (OP == 0) cf = load_coef<SMK>(F, A.gni, g);
		const double2 xm = S.win[D][(K - D - 3) & 3], xc = S.win[D][(K - D - 2) & 3], xn = S.win[D][(K - D - 1) & 3];
		const double2 bb = (D + 2 <= 4) ? S.bq[(K - D - 2) & 3] : bold;
		if (SMK == 1) {
			const int ox = (g & 1) ? 2 : 1;                // .x (even column) is red iff the global row is even
			res.x = jf_residual_rb<OP>(cf, ox, bb.x, xm.x, S.wer[0], xc.x, xc.y, xn.x);
			res.y = jf_residual_rb<OP>(cf, 3 - ox, bb.y, xm.y, xc.x, xc.y, S.wer[1], xn.y);
		} else {
			res.x = jf_residual<OP>(cf, bb.x, xm.x, S.wer[0], xc.x, xc.y, xn.x);
			res.y = jf_residual<OP>(cf, bb.y, xm.y, xc.x, xc.y, S.wer[1], xn.y);
		}
		if (MASK) {
			const bool rok = g >= 0 && g < A.gni;
			if (!B.in0 || !rok) res.x = 0.0;
			if (!B.in1 || !rok) res.y = 0.0;
		}
		if (POST == POST_NORM) {
			if (B.st_ok && rho >= B.y0 && rho < B.y1) S.acc = fma_rn(res.y, res.y, fma_rn(res.x, res.x, S.acc));   // the norm is compared to 1e-10, not bitwise
		}
		// neighbours of stage D, row t-D-1 (published in the previous step): the centre row of the next step's residual
		S.wer[0] = shp[D][FJ_Y(tid - 1)];
		S.wer[1] = shp[D][FJ_X(tid + 1)];
	}
	// ---- restriction of the residual rows completed in the previous step (their east neighbours are visible now)
	if (POST == POST_RESTRICT) {
		constexpr int RP = (K - D - 3) & 3;               // slot of row rp = t-D-3
		S.rw[RP][2] = shp[D + 1][FJ_X(tid + 1)];          // column j0+2 of row rp
		if ((((K - D - 3) & 1) == 0)) {
			const int rp = t - D - 3;
			const int I = (rp >> 1) - 1;                   // coarse row (local) fed by fine rows rp-2 .. rp
			const int J = B.j0 >> 1;
			if (B.st_ok && I >= (B.y0 >> 1) && I < (B.y1 >> 1) && I < A.C.ni && J < A.C.pitch) {
				constexpr int R0 = (RP + 2) & 3, R1 = (RP + 3) & 3;    // rows rp-2, rp-1
				double sum = mul(A.R3.w[0], S.rw[R0][0]);
				if (SMK == 1) {
					// red-first numbering of the fine grid: the five red fine points (a + b even), then the four black ones
					sum = add(sum, mul(A.R3.w[2], S.rw[R0][2]));
					sum = add(sum, mul(A.R3.w[4], S.rw[R1][1]));
					sum = add(sum, mul(A.R3.w[6], S.rw[RP][0]));
					sum = add(sum, mul(A.R3.w[8], S.rw[RP][2]));
					sum = add(sum, mul(A.R3.w[1], S.rw[R0][1]));
					sum = add(sum, mul(A.R3.w[3], S.rw[R1][0]));
					sum = add(sum, mul(A.R3.w[5], S.rw[R1][2]));
					sum = add(sum, mul(A.R3.w[7], S.rw[RP][1]));
				} else if (OP == 2) {
					// power-of-two weights (1/16, 1/8, 1/4): exact products, add(sum, mul(w, r)) == fma(w, r, sum)
					sum = fma_rn(A.R3.w[1], S.rw[R0][1], sum);
					sum = fma_rn(A.R3.w[2], S.rw[R0][2], sum);
					sum = fma_rn(A.R3.w[3], S.rw[R1][0], sum);
					sum = fma_rn(A.R3.w[4], S.rw[R1][1], sum);
					sum = fma_rn(A.R3.w[5], S.rw[R1][2], sum);
					sum = fma_rn(A.R3.w[6], S.rw[RP][0], sum);
					sum = fma_rn(A.R3.w[7], S.rw[RP][1], sum);
					sum = fma_rn(A.R3.w[8], S.rw[RP][2], sum);
				} else {
					sum = add(sum, mul(A.R3.w[1], S.rw[R0][1]));
					sum = add(sum, mul(A.R3.w[2], S.rw[R0][2]));
					sum = add(sum, mul(A.R3.w[3], S.rw[R1][0]));
					sum = add(sum, mul(A.R3.w[4], S.rw[R1][1]));
					sum = add(sum, mul(A.R3.w[5], S.rw[R1][2]));
					sum = add(sum, mul(A.R3.w[6], S.rw[RP][0]));
					sum = add(sum, mul(A.R3.w[7], S.rw[RP][1]));
					sum = add(sum, mul(A.R3.w[8], S.rw[RP][2]));
				}
				A.bc[(size_t)I * A.C.pitch + J] = (J < A.C.nj) ? sum : 0.0;
			}
		}
		S.rw[(K - D - 2) & 3][0] = res.x; S.rw[(K - D - 2) & 3][1] = res.y;
	}
	// ---- publish the rows produced in this step
#pragma unroll
	for (int s = 0; s <= D; ++s) { shn[s][FJ_X(tid)] = S.win[s][(K - s) & 3].x; shn[s][FJ_Y(tid)] = S.win[s][(K - s) & 3].y; }
	if (POST == POST_RESTRICT) shn[D + 1][FJ_X(tid)] = res.x;     // only the east neighbour's left column is ever read
	__syncthreads();
}

template <int D, int PRE, int POST, bool MASK, int OP, int SMK>
__device__ __forceinline__ void jf_run(const FusedArgs &A, const JfBlock &B, double (*sh)[D + 2][FJ_PUB],
                                       double (*in_u)[FJ_COLS], double (*in_b)[FJ_COLS], int t0, int t1)
{
	JfState<D> S;
#pragma unroll
	for (int s = 0; s <= D; ++s)
#pragma unroll
		for (int k = 0; k < 4; ++k) S.win[s][k] = make_double2(0.0, 0.0);
#pragma unroll
	for (int k = 0; k < 4; ++k) { S.bq[k] = make_double2(0.0, 0.0); S.rw[k][0] = 0.0; S.rw[k][1] = 0.0; S.rw[k][2] = 0.0; }
	S.acc = 0.0; S.wer[0] = 0.0; S.wer[1] = 0.0;
	const LevelDev &F = A.F;
	auto row_ok = [&](int i) {
		if (!MASK) return true;
		const int g = F.i0 + i;
		return g >= 0 && g < A.gni && i >= -MGB_GHOST_ROWS && i < F.ni + MGB_GHOST_ROWS;
	};
	// rows t0-1 .. t0+FJ_PF-1 are requested up front, one commit group each (row t0-1 only feeds the b ring)
	for (int i = t0 - 1; i < t0 + FJ_PF; ++i) jf_request<PRE, MASK, 0>(A, B, in_u, in_b, i);
	if (!MASK || B.tma) jf_arrived<MASK>(B, t0 - 1);     // b of row t0-1 is read at step t0 (each step waits for its own row only)
	// coarse values of one coarse row (columns J0-1, J0); rows outside the coarse arrays are clamped (their fine rows are masked)
	auto load_c = [&](int I) -> double2 {
		if (MASK) {
			if (!B.ld_ok) return make_double2(0.0, 0.0);
			I = I < -MGB_GHOST_ROWS ? -MGB_GHOST_ROWS : (I > A.C.ni + MGB_GHOST_ROWS - 1 ? A.C.ni + MGB_GHOST_ROWS - 1 : I);
		}
		const double *c = A.uc + (ptrdiff_t)I * (ptrdiff_t)A.C.pitch + (B.j0 >> 1);
		return make_double2(c[-1], c[0]);
	};
	if (PRE == PRE_PROLONG || PRE == PRE_PROLONG_MULTADD) {
		const int T = t0 >> 1;
		S.cq[0] = load_c(T - 1); S.cq[1] = load_c(T); S.cq[2] = load_c(T + 1);
	}
	for (int t = t0; t <= t1; t += 4) {
		if (PRE == PRE_PROLONG || PRE == PRE_PROLONG_MULTADD) {
			const int T = (t >> 1) + 2;                    // the next group of four fine rows needs coarse rows T-1 (held), T, T+1
			S.cn[0] = load_c(T); S.cn[1] = load_c(T + 1);
		}
		jf_step<D, PRE, POST, MASK, OP, 0, SMK>(A, B, S, sh, in_u, in_b, t);
		jf_step<D, PRE, POST, MASK, OP, 1, SMK>(A, B, S, sh, in_u, in_b, t + 1);
		jf_step<D, PRE, POST, MASK, OP, 2, SMK>(A, B, S, sh, in_u, in_b, t + 2);
		jf_step<D, PRE, POST, MASK, OP, 3, SMK>(A, B, S, sh, in_u, in_b, t + 3);
		if (PRE == PRE_PROLONG || PRE == PRE_PROLONG_MULTADD) { S.cq[0] = S.cq[2]; S.cq[1] = S.cn[0]; S.cq[2] = S.cn[1]; }
	}
	if (!MASK || B.tma) {
		// the FJ_PF rows requested beyond the last step: a block must not retire with bulk copies into its shared memory in flight
		const int tl = t0 + ((t1 - t0) & ~3) + 3;
		for (int i = tl + 1; i <= tl + FJ_PF; ++i) jf_arrived<MASK>(B, i);
	}
	if (POST == POST_NORM || POST == POST_DOT) {
		const double s = block_sum<FJ_THREADS>(S.acc);
		if (B.tid == 0) A.partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
	}
}

// thread 0 of a block: spin until every listed flag has reached my version of its channel (bounded; see mgb_halo.cuh)
__device__ __forceinline__ void jf_wait_list(const FusedComm &X, int which, int n)
{
	if (*(volatile int *)X.status != 0) return;                // a wait already timed out: do not spin again, the host aborts
	const long long t0 = clock64();
	for (int w = 0; w < n; ++w) {
		const unsigned long long v = *(volatile const unsigned long long *)X.w_ver[which][w];
		while (ld_acquire_sys(X.w_flag[which][w]) < v) {
			if (clock64() - t0 > X.spin_limit) { atomicExch(X.status, 1); *(volatile int *)X.status_host = 1; return; }
		}
	}
}

template <int D, int PRE, int POST, int SMK>
__global__ void __launch_bounds__(FJ_THREADS, 4)
k_jfused(FusedArgs A)
{
	// dynamic shared memory: rows produced in the previous step, per stage (0..D) and the residual row (index D+1),
	// double-buffered; then the cp.async input rings of u and b (FJ_NR rows each)
	extern __shared__ __align__(16) unsigned char jf_smem[];
	double (*sh)[D + 2][FJ_PUB] = reinterpret_cast<double (*)[D + 2][FJ_PUB]>(jf_smem);
	double (*in_u)[FJ_COLS] = reinterpret_cast<double (*)[FJ_COLS]>(jf_smem + sizeof(double) * 2 * (D + 2) * FJ_PUB);
	double (*in_b)[FJ_COLS] = in_u + FJ_NR;
	const LevelDev &F = A.F;
	JfBlock B;
	B.bar = reinterpret_cast<unsigned long long *>(in_b + FJ_NR);
	B.s_bar = (unsigned)__cvta_generic_to_shared(B.bar);
	B.s_u = (unsigned)__cvta_generic_to_shared(in_u); B.s_b = (unsigned)__cvta_generic_to_shared(in_b);
	B.tma = F.pitch >= FJ_COLS;
	if (threadIdx.x == 0) {
		for (int k = 0; k < FJ_NR; ++k) mbar_init(B.bar + k, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	pdl_enter();                                          // nothing above touches global memory (programmatic dependent launch)
	B.tid = threadIdx.x;
	B.c0 = blockIdx.x * FJ_VALID;                         // first valid column of the tile
	B.j0 = B.c0 - FJ_HALO + 2 * B.tid;                    // this thread's columns j0, j0+1 (j0 even)
	B.y0 = blockIdx.y * A.rows;
	B.y1 = min(B.y0 + A.rows, F.ni);
	B.P = (ptrdiff_t)F.pitch;
	B.in0 = B.j0 >= 0 && B.j0 < F.nj; B.in1 = B.j0 + 1 >= 0 && B.j0 + 1 < F.nj;
	B.ld_ok = B.j0 >= 0 && B.j0 < F.pitch;                // the pair may be loaded (pad columns hold zeros)
	B.st_ok = B.j0 >= B.c0 && B.j0 < B.c0 + FJ_VALID && B.j0 < F.pitch;
	{
		const Coef c = load_coef<SMK>(F, A.gni, F.i0);    // uniform operator: one coefficient set
		B.cu.aS = vreg(c.aS); B.cu.aW = vreg(c.aW); B.cu.aC = vreg(c.aC); B.cu.aE = vreg(c.aE); B.cu.aN = vreg(c.aN);
		B.cu.dinv = vreg(c.dinv); B.cu.nS = vreg(c.nS);
		B.scale = vreg(SMK ? sub(1.0, A.scale) : A.scale);    // red-black: 1 - omega (A.scale carries omega)
		B.sd = vreg(A.scale * c.dinv);
	}
	// steps: stage 0 of row y0-D-1 is the first needed, the restriction of row y1 completes at step y1+D+3;
	// the first step is rounded down to a multiple of four (ring slots are compile-time functions of t & 3)
	const int tb = (B.y0 - D - 1) & ~3, te = B.y1 + D + 3;
	B.rbase = tb - 1;
	if (A.X.nw_top | A.X.nw_bot | A.X.nw_all) {
		// ghost rows above the strip are read by the blocks whose first step lies above row 0, ghost rows below it by those
		// whose last step reaches row ni (the coarse rows of a prolongation follow the same fine rows)
		if (threadIdx.x == 0) {
			if (A.X.nw_all) jf_wait_list(A.X, 2, A.X.nw_all);
			if (A.X.nw_top && tb - 1 < 0) jf_wait_list(A.X, 0, A.X.nw_top);
			if (A.X.nw_bot && te + 1 >= F.ni) jf_wait_list(A.X, 1, A.X.nw_bot);
		}
		__syncthreads();
	}
	B.g_u = A.u_in + (B.c0 - FJ_HALO); B.g_b = A.b + (B.c0 - FJ_HALO);
	// interior blocks: every row and column this block touches lies inside the grid and inside this strip's arrays
	const bool interior = (B.c0 - FJ_HALO >= 0) && (B.c0 - FJ_HALO + FJ_COLS <= F.nj) &&
	                      (F.i0 + tb - 1 >= 0) && (F.i0 + te + 4 + FJ_PF < A.gni) && (tb - 1 >= -MGB_GHOST_ROWS) &&
	                      (te + 4 + FJ_PF < F.ni + MGB_GHOST_ROWS);
	if (F.uniform == 2) {
		if (interior) jf_run<D, PRE, POST, false, 2, SMK>(A, B, sh, in_u, in_b, tb, te);
		else          jf_run<D, PRE, POST, true, 2, SMK>(A, B, sh, in_u, in_b, tb, te);
	} else if (F.uniform == 1 && SMK == 0) {
		if (interior) jf_run<D, PRE, POST, false, 1, SMK>(A, B, sh, in_u, in_b, tb, te);
		else          jf_run<D, PRE, POST, true, 1, SMK>(A, B, sh, in_u, in_b, tb, te);
	} else            jf_run<D, PRE, POST, true, 0, SMK>(A, B, sh, in_u, in_b, tb, te);    // (red-black: the general path also for uniform, non-power-of-two operators)
	// ---- push: the rows of this block that lie in a push range travel to the peers AFTER the row loop (nothing of this is
	// in the hot loop): every thread re-reads exactly the elements it stored itself, so no fence is needed in between
	if (A.X.nch) {
		// (computed here, not carried through the row loop: the kernel runs at the 128-register cap)
		bool push_u = false, push_b = false;
		for (int k = 0; k < A.X.npu; ++k) push_u |= (D > 0 && B.y0 < A.X.pu[k].hi && B.y1 > A.X.pu[k].lo);
		for (int k = 0; k < A.X.npb; ++k) push_b |= (POST == POST_RESTRICT && (B.y0 >> 1) < A.X.pb[k].hi && (B.y1 >> 1) > A.X.pb[k].lo);
		const bool takes = push_u || push_b || (A.X.bc_remote && POST == POST_RESTRICT);
		if (push_u && B.st_ok) {
			for (int k = 0; k < A.X.npu; ++k) {
				double *dst = A.X.pu[k].dst;
				if (!dst) continue;
				const int c0 = max(B.y0, A.X.pu[k].lo), c1 = min(B.y1, A.X.pu[k].hi);
				for (int c = c0; c < c1; ++c) st2(dst + (ptrdiff_t)c * B.P + B.j0, ld2(A.u_out + (ptrdiff_t)c * B.P + B.j0));
			}
		}
		if (push_b && B.st_ok) {
			const int J = B.j0 >> 1;
			if (J < A.C.pitch) {
				for (int k = 0; k < A.X.npb; ++k) {
					double *dst = A.X.pb[k].dst;
					if (!dst) continue;
					const int I0 = max(B.y0 >> 1, A.X.pb[k].lo), I1 = min(min(B.y1 >> 1, A.C.ni), A.X.pb[k].hi);
					for (int I = I0; I < I1; ++I) dst[(size_t)I * A.C.pitch + J] = A.bc[(size_t)I * A.C.pitch + J];
				}
			}
		}
		// ---- signal: the last pushing block makes every pushed row visible system-wide, then raises the flags
		if (takes) {
			__syncthreads();                              // the block's stores are ordered before thread 0's fence (cumulativity)
			if (threadIdx.x == 0) {
				__threadfence_system();
				const unsigned int t = atomicAdd(A.X.ticket, 1u);
				if (t == (unsigned)A.X.npushblocks - 1u) {
					*A.X.ticket = 0u;
					__threadfence_system();
					for (int c = 0; c < A.X.nch; ++c) {
						const unsigned long long newv = *A.X.ver[c] + 1ull;
						for (int k = 0; k < A.X.nsig[c]; ++k) st_relaxed_sys(A.X.sig[c][k], newv);
						*A.X.ver[c] = newv;
					}
				}
			}
		}
	}
}
