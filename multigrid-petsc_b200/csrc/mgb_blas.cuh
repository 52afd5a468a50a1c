// mgb_blas.cuh -- vector kernels of the Krylov wrapper and the stopping tests: nrm2, dot, axpy family,
// separable RHS / error evaluation.  Vectors are processed over their whole padded extent (rows 0..ni-1,
// all `pitch` columns): the pad columns hold zeros, contribute nothing to sums and stay zero under
// y <- y + a x.  Reductions are two-pass and deterministic: fixed grid, per-block partial sums in a fixed
// order, then one block sums the partials in a fixed order (run-to-run reproducible).
#pragma once
#include "mgb_common.cuh"

#define MGB_RED_THREADS 256
#define MGB_RED_MAXBLOCKS 2368      // 148 SMs x 16 : partial-sum buffer size

// op 0: sum x*x   op 1: sum x*y
template <int OP>
__global__ void __launch_bounds__(MGB_RED_THREADS)
k_reduce1(const double *__restrict__ x, const double *__restrict__ y, size_t n2, double *__restrict__ partial)
{
	pdl_enter();
	// n2 = number of double2 elements
	double acc = 0.0;
	const size_t stride = (size_t)gridDim.x * MGB_RED_THREADS;
	for (size_t k = (size_t)blockIdx.x * MGB_RED_THREADS + threadIdx.x; k < n2; k += stride) {
		const double2 a = ld2(x + 2 * k);
		if (OP == 0) acc += a.x * a.x + a.y * a.y;
		else { const double2 c = ld2(y + 2 * k); acc += a.x * c.x + a.y * c.y; }
	}
	const double s = block_sum<MGB_RED_THREADS>(acc);
	if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// final pass: out[slot] = f(sum partial[0..n)) ; f = sqrt when take_sqrt
__global__ void __launch_bounds__(1024)
k_reduce2(const double *__restrict__ partial, int n, double *__restrict__ out, int slot, int take_sqrt)
{
	double acc = 0.0;
	for (int k = threadIdx.x; k < n; k += 1024) acc += partial[k];
	const double s = block_sum<1024>(acc);
	if (threadIdx.x == 0) out[slot] = take_sqrt ? sqrt(s) : s;
}

// kind 0: y = y + alpha*x  (VecAXPY)   kind 1: y = x + alpha*y  (VecAYPX)   kind 2: y = x (copy)   kind 3: y = 0
// alpha is read from a device scalar when alpha_dev != nullptr (value alpha_sign * alpha_dev[0]) so that the
// CG recurrences need no host round trip.
template <int KIND>
__global__ void __launch_bounds__(256)
k_axpy(double *__restrict__ y, const double *__restrict__ x, size_t n2, double alpha,
       const double *__restrict__ alpha_dev, double alpha_sign)
{
	pdl_enter();
	if (alpha_dev) alpha = alpha_sign * alpha_dev[0];
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
		double2 o;
		if (KIND == 3) { o.x = 0.0; o.y = 0.0; }
		else {
			const double2 a = ld2(x + 2 * k);
			if (KIND == 2) o = a;
			else {
				const double2 c = ld2(y + 2 * k);
				if (KIND == 0) { o.x = add(c.x, mul(alpha, a.x)); o.y = add(c.y, mul(alpha, a.y)); }
				else           { o.x = add(a.x, mul(alpha, c.x)); o.y = add(a.y, mul(alpha, c.y)); }
			}
		}
		st2(y + 2 * k, o);
	}
}

// b[i][j] = gx[j] * gy[i]   (separable right-hand side, one rounding: identical to the host product)
__global__ void __launch_bounds__(256)
k_outer(double *__restrict__ v, const double *__restrict__ gx, const double *__restrict__ gy, LevelDev L)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int i = blockIdx.y;
	if (j >= L.pitch) return;
	v[(size_t)i * L.pitch + j] = (j < L.nj) ? mul(gx[j], gy[L.i0 + i]) : 0.0;
}

// error triple against s[i][j] = sx[j]*sy[i]: partial[3*blk + {0,1,2}] = {max, sum, sum of squares} of |u - s|
__global__ void __launch_bounds__(MGB_RED_THREADS)
k_error(const double *__restrict__ u, const double *__restrict__ sx, const double *__restrict__ sy, LevelDev L,
        double *__restrict__ partial)
{
	__shared__ double smax[MGB_RED_THREADS / 32];
	double mx = 0.0, s1 = 0.0, s2 = 0.0;
	const size_t total = (size_t)L.ni * L.pitch;
	const size_t stride = (size_t)gridDim.x * MGB_RED_THREADS;
	for (size_t k = (size_t)blockIdx.x * MGB_RED_THREADS + threadIdx.x; k < total; k += stride) {
		const int i = (int)(k / L.pitch), j = (int)(k - (size_t)i * L.pitch);
		if (j < L.nj) {
			const double d = fabs(sub(u[k], mul(sx[j], sy[L.i0 + i])));
			mx = fmax(mx, d); s1 += d; s2 += d * d;
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
	if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
	const double t1 = block_sum<MGB_RED_THREADS>(s1);
	const double t2 = block_sum<MGB_RED_THREADS>(s2);
	if (threadIdx.x == 0) {
		double m = 0.0;
		for (int w = 0; w < MGB_RED_THREADS / 32; ++w) m = fmax(m, smax[w]);
		partial[3 * blockIdx.x + 0] = m; partial[3 * blockIdx.x + 1] = t1; partial[3 * blockIdx.x + 2] = t2;
	}
}
// finalize = 0 leaves the sum of squares unrooted (several strips are combined by the caller)
__global__ void k_error2(const double *__restrict__ partial, int n, double *__restrict__ out, int finalize)
{
	if (threadIdx.x != 0) return;
	double m = 0.0, s1 = 0.0, s2 = 0.0;
	for (int k = 0; k < n; ++k) { m = fmax(m, partial[3 * k]); s1 += partial[3 * k + 1]; s2 += partial[3 * k + 2]; }
	out[0] = m; out[1] = s1; out[2] = finalize ? sqrt(s2) : s2;
}

// Host arrays are dense (row length nj), level vectors are pitched: PCIe copies go through a dense staging buffer
// (a pitched cudaMemcpy2D from host memory runs at half the link rate) and these kernels repack at HBM speed.
__global__ void __launch_bounds__(256)
k_unpack_rows(double *__restrict__ v, const double *__restrict__ dense, int ni, int nj, int pitch)
{
	const size_t total = (size_t)ni * nj;
	for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
		const size_t i = k / nj, j = k - i * nj;
		v[i * pitch + j] = dense[k];
	}
}
__global__ void __launch_bounds__(256)
k_pack_rows(double *__restrict__ dense, const double *__restrict__ v, int ni, int nj, int pitch)
{
	const size_t total = (size_t)ni * nj;
	for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
		const size_t i = k / nj, j = k - i * nj;
		dense[k] = v[i * pitch + j];
	}
}

// device scalars -> mapped pinned host memory by a kernel store: the per-cycle read-back of the residual norm must
// not queue on a copy engine behind a half-gigabyte upload or download that overlaps the solve
__global__ void k_publish(double *__restrict__ host_mapped, const double *__restrict__ dev, int first, int count)
{
	if ((int)threadIdx.x < count) host_mapped[first + threadIdx.x] = dev[first + threadIdx.x];
	__threadfence_system();
}

// The CG vector update in one pass (KSPSolve_CG: VecAXPY(x, a, p); VecAXPY(r, -a, w); VecNorm(r)):
// x += a p ; r -= a w ; partial[block] = sum r^2.  40 B per unknown instead of 24 + 24 + 8.
// x == nullptr: the x update is deferred into the next direction step (k_cg_pstep): r -= a w ; ||r|| only, 24 B.
__global__ void __launch_bounds__(MGB_RED_THREADS)
k_cg_update(double *__restrict__ x, const double *__restrict__ p, double *__restrict__ r, const double *__restrict__ w,
            size_t n2, double a, double *__restrict__ partial, const double *__restrict__ a_dev)
{
	pdl_enter();
	if (a_dev) a = a_dev[0];                      // alpha = beta / p'w as left by the reduction tail (TAIL_DPI)
	double acc = 0.0;
	const double ma = -a;
	const size_t stride = (size_t)gridDim.x * MGB_RED_THREADS;
	for (size_t k = (size_t)blockIdx.x * MGB_RED_THREADS + threadIdx.x; k < n2; k += stride) {
		const double2 rv = ld2(r + 2 * k), wv = ld2(w + 2 * k);
		double2 ro;
		if (x) {
			const double2 xv = ld2(x + 2 * k), pv = ld2(p + 2 * k);
			double2 xo;
			xo.x = add(xv.x, mul(a, pv.x)); xo.y = add(xv.y, mul(a, pv.y));
			st2(x + 2 * k, xo);
		}
		ro.x = add(rv.x, mul(ma, wv.x)); ro.y = add(rv.y, mul(ma, wv.y));
		st2(r + 2 * k, ro);
		acc += ro.x * ro.x + ro.y * ro.y;
	}
	const double s = block_sum<MGB_RED_THREADS>(acc);
	if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
