// mgb_common.cuh -- shared device-side definitions of the B200 multigrid engine.
//
// HBM layout of one level vector (DESIGN.md "Data layout"):
//   * dense row-major ni x nj interior unknowns (the reference's natural numbering i*nj + j,
//     ref: src/matbuild.c:296-303), stored with a row pitch that is a multiple of 16 doubles (128 B)
//     and >= nj + 1;
//   * columns nj .. pitch-1 of every row are ZERO: they are the right Dirichlet ghost of row i and, at
//     address (i+1)*pitch - 1, the left ghost (j = -1) of row i+1;
//   * MGB_GHOST_ROWS ghost rows above row 0 and below row ni-1: zero at the physical boundary (Dirichlet),
//     copies of the neighbour strips' boundary rows otherwise (pushed by mgb_halo.cuh after every write);
//   * element (0,0) is 128-byte aligned.
// With that layout the 5-point stencil needs no boundary branches and every row start is aligned for
// 16-byte vector loads (for the reference's n = 2^k - 1 grids pitch == n + 1 exactly: no wasted bytes).
//
// Arithmetic convention: IEEE binary64, round-to-nearest, NO fused multiply-add on any value that is
// compared against the reference (the reference is plain C over PETSc's unfused loops).  All products and
// sums are written with __dmul_rn/__dadd_rn/__dsub_rn so the order of operations is PETSc's, bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MGB_GHOST_ROWS 6     // ghost rows above and below every strip (room for fused multi-sweep kernels)
#define MGB_COEF_STRIDE 8     // doubles per grid row in the coefficient table: S W C E N dinv idiag mdiag

struct LevelDev {
	int ni, nj;          // local interior rows, columns
	int pitch;           // doubles per row
	int i0;              // global grid-row index of local row 0 (0 on a single GPU)
	int gni;             // global number of grid rows of the level
	int uniform;         // 1: every grid row has the same coefficients (mesh 0); 2: ... and S = W = E = N = 2^m, C = -4 * 2^m
	int rb;              // 1: red-black numbering (-map 3): row sums run in ascending RED-FIRST column order
	const double *coef;  // MGB_COEF_STRIDE doubles per GLOBAL grid row
};

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
// fused multiply-add: used ONLY where the product is exact (one factor a power of two), so that the single rounding of
// the sum gives the same bits as the reference's separate multiply and add
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become
// resident while its predecessor in the stream is still draining; it must not touch memory before pdl_wait() (which
// returns once the predecessor grid has completed and its writes are visible -- a no-op in an ordinary launch).
// pdl_go() lets the successor's blocks be scheduled as soon as every block of this grid has started: launch latency
// and block start-up of the ~26 dependent launches of a V-cycle overlap the tail of the previous kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_go() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_go(); }

// (A x)_ij in MatMult_SeqAIJ order: ascending columns row-n, row-1, row, row+1, row+n
// (ref call sites: src/solver.c:1516,1534,1545; the ghost zeros stand in for the entries fillJacobians drops)
__device__ __forceinline__ double stencil5(double aS, double aW, double aC, double aE, double aN,
                                           double xS, double xW, double xC, double xE, double xN)
{
	double s = mul(aS, xS);
	s = add(s, mul(aW, xW));
	s = add(s, mul(aC, xC));
	s = add(s, mul(aE, xE));
	s = add(s, mul(aN, xN));
	return s;
}

// The same row sum when the unknowns are numbered red-first (-map 3; red = (i+j) even): a red row stores its
// diagonal first (every black neighbour has a higher number), a black row stores it last; within one colour the
// natural order S < W < E < N is kept.  order: 0 natural, 1 red row, 2 black row.
__device__ __forceinline__ double stencil5_ord(int order, double aS, double aW, double aC, double aE, double aN,
                                               double xS, double xW, double xC, double xE, double xN)
{
	if (order == 0) return stencil5(aS, aW, aC, aE, aN, xS, xW, xC, xE, xN);
	double s;
	if (order == 1) { s = mul(aC, xC); s = add(s, mul(aS, xS)); } else s = mul(aS, xS);
	s = add(s, mul(aW, xW));
	s = add(s, mul(aE, xE));
	s = add(s, mul(aN, xN));
	if (order == 2) s = add(s, mul(aC, xC));
	return s;
}

// block-wide sum in a fixed order (deterministic): warp shuffles, then warp 0 over the warp sums
template <int THREADS>
__device__ __forceinline__ double block_sum(double v)
{
	__shared__ double red[THREADS / 32];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (lane == 0) red[w] = v;
	__syncthreads();
	v = 0.0;
	if (w == 0) {
		v = (lane < THREADS / 32) ? red[lane] : 0.0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	}
	__syncthreads();
	return v;   // valid in thread 0
}
