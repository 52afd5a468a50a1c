// mgb_halo.cuh -- strip-to-strip data movement over NVLink peer memory: ghost-row exchange, gather of the first
// agglomerated level onto rank 0, broadcast of its correction, all-reduce of norm/dot partials.
//
// Replaces what PETSc's MPIAIJ VecScatter (ghost values of every MatMult), MPI_Allreduce (VecNorm/VecDot) and
// the hand-written MPI_Send/MPI_Recv gather (ref: src/solver.c:1273-1299) do for the reference.  There is no
// NCCL call on this path: one generic kernel stores rows straight into the peer's HBM (addresses obtained with
// cudaIpcOpenMemHandle, or plain device pointers when several strips live in one process) and then raises a
// flag in the peer's memory; the receiving side spins on its own flag word.
//
// Protocol.  Every transfer site is a CHANNEL c with a device-resident version counter ver[c] that all ranks
// advance in lock step (every rank launches the kernel for every use of a channel, even with nothing to send).
//   push:  newv = ver[c] + 1; all blocks copy their share; the last block to finish (atomic ticket) issues
//          __threadfence_system(), stores newv into flag[c][my_rank] of every destination (st.relaxed.sys)
//          and sets ver[c] = newv.
//   wait:  v = ver[c]; spin (ld.acquire.sys) until flag[c][src] >= v for every expected source.
// In a multi-process run (one strip per process, one GPU each) push and wait are ONE launch: the last block
// signals and then waits, so a rank can never run more than one exchange ahead of its neighbours (this is what
// makes overwriting the neighbour's ghost rows safe, see DESIGN.md "Halo protocol").  When several strips are
// emulated on one GPU (tests) the engine launches all pushes first and the waits afterwards, so that no kernel
// ever waits for a kernel queued behind it.  Spins are bounded (~4 s): on timeout status[0] is set and the
// host reports MGB_ECUDA instead of hanging the GPU.
#pragma once
#include "mgb_common.cuh"

#define MGB_MAX_RANKS 8
#define MGB_XFER_MAX 24                             // (source, destination) pairs per launch: several vectors travel together
#define MGB_XFER_THREADS 256

struct XferArgs {
	int ndst;                                       // copies of this launch (0..MGB_XFER_MAX)
	const double *src[MGB_XFER_MAX];
	double *dst[MGB_XFER_MAX];
	unsigned long long cnt2[MGB_XFER_MAX];          // double2 elements per copy
	unsigned long long *peer_flag[MGB_XFER_MAX];    // flag word in the destination's memory (its flag[c][my_rank])
	int nwait;
	const unsigned long long *wait_flag[MGB_MAX_RANKS]; // my own flag words flag[c][src_rank]
	unsigned long long *ver;                        // my version counter of the channel
	unsigned int *ticket;                           // my block ticket counter of the channel (zero between launches)
	int *status;                                    // status[0] != 0 after a timeout
	int *status_host;                               // mapped host mirror of status[0] (the host tests it after every synchronisation)
	int do_push, do_wait;
	long long spin_limit;                           // clock64 ticks
	unsigned long long parity_stride;               // doubles added to every dst when the new version is odd (all-reduce slots)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(MGB_XFER_THREADS)
k_xfer(XferArgs a)
{
	__shared__ int is_last;
	if (a.do_push) {
		// safe to read: ver is only advanced by the last block, after every block has taken its ticket
		const unsigned long long par = (a.parity_stride && ((*a.ver + 1ull) & 1ull)) ? a.parity_stride : 0ull;
		for (int d = 0; d < a.ndst; ++d) {
			const double2 *s = reinterpret_cast<const double2 *>(a.src[d]);
			double2 *t = reinterpret_cast<double2 *>(a.dst[d] + par);
			for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < a.cnt2[d];
			     k += (unsigned long long)gridDim.x * blockDim.x)
				t[k] = s[k];
		}
		__syncthreads();                              // the block's stores are ordered before thread 0's fence (cumulativity)
		if (threadIdx.x == 0) {
			__threadfence_system();
			const unsigned int t = atomicAdd(a.ticket, 1u);
			is_last = (t == gridDim.x - 1);
		}
		__syncthreads();
		if (!is_last) return;
		if (threadIdx.x == 0) {
			*a.ticket = 0u;
			const unsigned long long newv = *a.ver + 1ull;
			__threadfence_system();                   // one fence, then the flag stores go out back to back (fence + relaxed
			for (int d = 0; d < a.ndst; ++d) st_relaxed_sys(a.peer_flag[d], newv);   // store = release pattern)
			*a.ver = newv;
		}
	} else if (blockIdx.x != 0) return;
	if (a.do_wait && threadIdx.x == 0) {
		const unsigned long long v = *(volatile unsigned long long *)a.ver;
		const long long t0 = clock64();
		// after one timeout the ranks are out of step for good: later exchanges do not spin again (the host sees the
		// status word after its next synchronisation and aborts the solve)
		bool dead = *(volatile int *)a.status != 0;
		for (int w = 0; w < a.nwait && !dead; ++w) {
			while (ld_acquire_sys(a.wait_flag[w]) < v) {
				if (clock64() - t0 > a.spin_limit) {
					atomicExch(a.status, 1);
					if (a.status_host) *(volatile int *)a.status_host = 1;
					dead = true; break;
				}
			}
		}
		__threadfence_system();
	}
}

// The tail of every norm / dot product in ONE launch of one block: sum of the per-block partials in a fixed order, then
//   single strip:  out[slot] = f(sum), mirrored into mapped host memory (no separate publish launch);
//   row strips:    the local sum goes into slot [my rank] of every rank (parity sets as in k_reduce_ranks), the flags are
//                  raised, the peers' flags are awaited and the slots are summed in rank order -- identical values and
//                  identical stopping decisions on every rank (replaces k_reduce2 + k_xfer + k_reduce_ranks + k_publish).
// do_push / do_wait split the two halves for strips emulated on one GPU (all pushes first, then the waits).
struct TailArgs {
	const double *partial; int n;
	double *scal; int slot, take_sqrt;
	double *host;                                   // mapped host mirror of scal[] (or null)
	int nranks;
	double *slot_dst[MGB_MAX_RANKS];                // slot [my rank] in every rank's slot array (parity offset added here)
	unsigned long long *peer_flag[MGB_MAX_RANKS];
	const unsigned long long *wait_flag[MGB_MAX_RANKS];
	const double *my_slots;
	unsigned long long *ver;
	unsigned long long parity_stride;
	int *status, *status_host;
	long long spin_limit;
	int do_push, do_wait;
	int post_op;                                    // TAIL_*: CG scalars derived on the device from the reduced value
};
// scal[] slots of the CG recurrences (KSPSolve_CG: beta = z'r, b = beta / betaold, dpi = p'w, a = beta / dpi); they are
// derived where the dot products finish, so that the vector kernels read them from HBM and the host reads only ||r||
#define SC_BETA 16
#define SC_BETAOLD 17
#define SC_RATIO 18
#define SC_DPI 19
#define SC_ALPHA 20
enum { TAIL_NONE = 0, TAIL_BETA = 1, TAIL_DPI = 2 };

__device__ __forceinline__ void tail_finish(const TailArgs &a, double v)
{
	a.scal[a.slot] = v;
	if (a.host) a.host[a.slot] = v;
	if (a.post_op == TAIL_BETA) {
		const double old = a.scal[SC_BETA];           // +inf before the first iteration: ratio 0, p = z + 0 p
		const double ratio = v / old;
		a.scal[SC_BETAOLD] = old; a.scal[SC_BETA] = v; a.scal[SC_RATIO] = ratio;
		if (a.host) { a.host[SC_BETAOLD] = old; a.host[SC_BETA] = v; a.host[SC_RATIO] = ratio; }
	} else if (a.post_op == TAIL_DPI) {
		const double al = a.scal[SC_BETA] / v;
		a.scal[SC_DPI] = v; a.scal[SC_ALPHA] = al;
		if (a.host) { a.host[SC_DPI] = v; a.host[SC_ALPHA] = al; }
	}
	if (a.host) __threadfence_system();
}

__global__ void __launch_bounds__(1024)
k_reduce_tail(TailArgs a)
{
	pdl_enter();
	double s = 0.0;
	if (a.do_push) {
		double acc = 0.0;
		for (int k = threadIdx.x; k < a.n; k += 1024) acc += a.partial[k];
		s = block_sum<1024>(acc);
	}
	if (threadIdx.x != 0) return;
	if (a.nranks <= 1) {
		tail_finish(a, a.take_sqrt ? sqrt(s) : s);
		return;
	}
	if (a.do_push) {
		const unsigned long long newv = *a.ver + 1ull;
		const unsigned long long par = (newv & 1ull) ? a.parity_stride : 0ull;
		for (int q = 0; q < a.nranks; ++q) a.slot_dst[q][par] = s;
		__threadfence_system();
		for (int q = 0; q < a.nranks; ++q) st_relaxed_sys(a.peer_flag[q], newv);
		*a.ver = newv;
	}
	if (a.do_wait) {
		const unsigned long long v = *(volatile unsigned long long *)a.ver;
		const long long t0 = clock64();
		bool dead = *(volatile int *)a.status != 0;
		for (int q = 0; q < a.nranks && !dead; ++q) {
			while (ld_acquire_sys(a.wait_flag[q]) < v) {
				if (clock64() - t0 > a.spin_limit) {
					atomicExch(a.status, 1);
					if (a.status_host) *(volatile int *)a.status_host = 1;
					dead = true; break;
				}
			}
		}
		__threadfence_system();
		const double *sl = a.my_slots + ((v & 1ull) ? a.parity_stride : 0ull);
		double t = 0.0;
		for (int r = 0; r < a.nranks; ++r) t += *(volatile const double *)(sl + r * 4);
		tail_finish(a, a.take_sqrt ? sqrt(t) : t);
	}
}

__global__ void k_set_scalar(double *p, double v) { pdl_enter(); *p = v; }

// keeps a channel's version counter in step on a rank that takes no part in a transfer done inside a compute kernel
__global__ void k_bump(unsigned long long *ver) { pdl_enter(); *ver = *ver + 1ull; }

// out[out_slot + k] = f(sum over ranks of slot[r][k]) in rank order (identical on every rank): the all-reduce tail.
// slots: two parity sets of MGB_MAX_RANKS x 4 doubles (the set of the version just completed is read, so a fast
// rank's next all-reduce cannot overwrite values a slow rank has not summed yet); take_sqrt applies to value 0.
__global__ void k_reduce_ranks(const double *__restrict__ slots, unsigned long long parity_stride,
                               const unsigned long long *__restrict__ ver, int nranks, int nvals,
                               double *__restrict__ out, int out_slot, int take_sqrt)
{
	if (threadIdx.x >= nvals) return;
	const double *sl = slots + ((*ver & 1ull) ? parity_stride : 0ull);
	double s = 0.0;
	for (int r = 0; r < nranks; ++r) s += sl[r * 4 + threadIdx.x];
	out[out_slot + threadIdx.x] = (take_sqrt && threadIdx.x == 0) ? sqrt(s) : s;
}
