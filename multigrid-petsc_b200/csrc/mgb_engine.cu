// mgb_engine.cu -- the B200 multigrid engine behind include/mgb200.h.
//
// Host-side orchestration (level hierarchy, row-strip decomposition, HBM arena, kernel sequencing of the V-cycle
// and of the MG-preconditioned Krylov wrapper, CUDA-graph replay) plus the extern "C" entry points.  The kernels
// are in mgb_stencil.cuh / mgb_transfer.cuh / mgb_blas.cuh / mgb_csr.cuh / mgb_coarse.cuh / mgb_halo.cuh.
// There is NO CPU fallback: every entry point fails with MGB_ECUDA when no device is usable.
//
// Reference sequencing that this file reproduces ("ref:" = /root/reference):
//   cycle 0  MultigridVcycle      ref: src/solver.c:1414-1575 (hot loop :1530-1550)
//   cycle 8  MultigridPetscPCMG   ref: src/solver.c:1884-1989, with PETSc's KSPCG / KSPRICHARDSON / PCMG
//            semantics as restated in oracle/minipetsc/minipetsc.c ([PETSc-upstream], unpinned).
//   row ranges per rank           ref: src/matbuild.c:120-144 (GetRanges) -- re-designed as whole-row strips
//
// Strips.  With nranks > 1 the levels whose row count exceeds cfg.agglomerate_below are split into contiguous
// row strips, one per rank (= one GPU, one process); the coarser levels live whole on rank 0.  The partition is
// defined on the first agglomerated level (rows c_r = floor(r * n / P)) and doubled upwards (rank r owns fine
// rows [2 c_r, 2 c_{r+1}), the last rank also the final row), so that restriction needs one foreign residual
// row and prolongation one foreign coarse row.  Every kernel that writes a distributed vector is
// followed by a ghost-row push into the neighbours' HBM (mgb_halo.cuh).  A process holds ONE strip in
// production (peers reached through CUDA IPC) or ALL strips on one GPU when cfg.emulate is set (tests: the
// same kernels and the same protocol, launched in lock step on one stream).
#include "../../include/mgb200.h"
#include "mgb_common.cuh"
#include "mgb_stencil.cuh"
#include "mgb_transfer.cuh"
#include "mgb_blas.cuh"
#include "mgb_csr.cuh"
#include "mgb_coarse.cuh"
#include "mgb_halo.cuh"
#include "mgb_fused.cuh"
#include "mgb_coarse_cycle.cuh"
#include "mgb_wave.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
	return code;
}
extern "C" void mgb__set_error(const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg ? msg : ""); }   // for mgb_sparse.cu
#define CU(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) \
	return fail(MGB_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); } while (0)
#define TRY(call) do { int _r = (call); if (_r != MGB_OK) return _r; } while (0)
#define KCHECK() CU(cudaGetLastError())
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Launch with programmatic stream serialisation (PDL): the kernel may become resident while its predecessor drains; every
// kernel launched through here starts with pdl_enter() (mgb_common.cuh).  Inside a stream capture the launch becomes a
// programmatic edge of the graph.  Off by default: measured on B200 (gpurun_out/r2q_*), the replayed cycle graph gains
// nothing at 1025^2 and above (the legs are bound by their own row pipeline, not by the launch gap) and 4-15 % only on
// 129^2-sized problems; MGB_PDL=1 turns it on.  A launch error is left for KCHECK().
static bool g_pdl = false;
template <class... KP, class... A>
static void klaunch(void (*kern)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A &&...args)
{
	cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
	(void)cudaLaunchKernelEx(&cfg, kern, static_cast<KP>(args)...);
}

// ------------------------------------------------------------------------------------------------ engine state
#define MGB_MAXL 32
#define NOOFF ((size_t)-1)
// channels: halo of (level, physical buffer), then gather / broadcast per level, then the all-reduce
#define CH_HALO(l, k) ((l) * MGB_NVEC + (k))
#define CH_GATHER(l)  (MGB_MAXL * MGB_NVEC + (l))
#define CH_BCAST(l)   (MGB_MAXL * MGB_NVEC + MGB_MAXL + (l))
#define CH_REDUCE     (MGB_MAXL * MGB_NVEC + 2 * MGB_MAXL)
#define MGB_NCHAN     (CH_REDUCE + 1)
#define RED_VALS 4                          // doubles per rank slot of the all-reduce
#define SC_LOCAL 32                         // scal[] index where a strip parks its local partial before the all-reduce

struct Csr {
	int m = 0, n = 0; long long nnz = 0;
	int *rowptr = nullptr, *col = nullptr; double *val = nullptr;
};

// geometry of one level: global part, identical on every rank
struct LevelGeom {
	int gni = 0, nj = 0, pitch = 0;
	bool dist = false;                       // split into row strips (else whole, on rank 0)
	int rows[MGB_MAX_RANKS + 1] = {};        // partition: rank r handles global rows [rows[r], rows[r+1])
	std::vector<double> coef_host;           // gni * MGB_COEF_STRIDE
	bool coef_set = false;
	int uniform = 1;
};

// arena layout of one rank (a pure function of the configuration, so every rank can address its peers)
struct Layout {
	size_t vec_off[MGB_MAXL][MGB_NVEC];      // byte offset of each vector allocation, NOOFF if absent
	size_t origin[MGB_MAXL];                 // doubles from allocation start to element (0,0)
	size_t alloc[MGB_MAXL];                  // doubles per vector allocation
	int ni[MGB_MAXL], r0[MGB_MAXL];          // local rows / first global row on this rank
	size_t flags_off, slots_off, ver_off, ticket_off, status_off, total;
};

struct SLevel {                              // one level on one strip
	int ni = 0, r0 = 0;
	bool present = false;                    // arrays exist on this strip
	bool active = false;                     // this strip computes on the level (dist, or whole on rank 0)
	double *base[MGB_NVEC] = {};
	double *v[MGB_NVEC] = {};                // v[k] = pointer to element (0,0); the Jacobi ping-pong swaps two of them
	int phys[MGB_NVEC] = {};                 // which physical allocation v[k] currently points into
	double *coef = nullptr;                  // device copy of LevelGeom::coef_host (all global rows)
	Csr A, R, P;
	BandLU lu;
	double *ilu = nullptr;                    // ILU(0) of A[l]: inverted pivots, S multipliers, W multipliers (3 x ni x pitch), lazily
	bool ilu_valid = false;
};

struct Strip {
	int rank = 0, device = 0;
	char *arena = nullptr;
	std::vector<SLevel> lev;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	double *partial = nullptr; size_t partial_cap = 0;
	double *scal = nullptr, *scal_host = nullptr, *scal_host_dev = nullptr;   // device scalars, mapped pinned mirror (+ its device alias)
	double *tab_x = nullptr, *tab_y = nullptr;
	int *wave_progress = nullptr;            // per-panel row counters of the wavefront sweeps (mgb_wave.cuh)
	int *status_host = nullptr, *status_host_dev = nullptr;   // mapped pinned: a timed-out wait is visible to the host without a copy
	double *stage_in = nullptr, *stage_out = nullptr;     // dense staging buffers for PCIe copies of large vectors (lazy)
	size_t stage_cap = 0;
	cudaStream_t copy_in = nullptr, copy_out = nullptr;   // host <-> device copies overlapped with a solve (mgb_solve_vcycle_many)
	cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_sol = nullptr;
};

struct PtrState { double *v[MGB_MAXL][MGB_NVEC]; int phys[MGB_MAXL][MGB_NVEC]; };
struct GraphEntry { std::vector<char> key; cudaGraphExec_t exec; long long launches; std::vector<PtrState> after; };
enum { REQ_HALO = 0, REQ_GATHER = 1, REQ_BCAST = 2 };
struct XferReq { int type, level, phys, depth; };   // a deferred strip-to-strip transfer (see flush_levels)

struct mgb_engine {
	mgb_config cfg;
	int P = 1, L = 1, La = 0;                // ranks, levels, first agglomerated level (== L: none)
	std::vector<LevelGeom> geo;
	std::vector<Layout> lay;                 // per rank
	std::vector<Strip> strips;               // local strips: 1 (production) or P (emulation)
	char *arena_of[MGB_MAX_RANKS] = {};      // arena base of every rank as addressable from this process
	bool peer_opened[MGB_MAX_RANKS] = {};
	bool connected = false;
	Stencil3 R3, P3; bool transfer_set = false;
	bool transfer_pow2 = false;              // every res / pro weight is a power of two (exact products: the fused kernels may use fma)
	double sor_omega = -1.0;
	long long launches = 0;
	double last_solve_ms = 0.0;
	bool csr_built = false;
	std::vector<struct GraphEntry> gcache;   // instantiated V-cycle graphs, keyed by parameters + pointer state
	long long spin_limit = 8000000000LL;     // ~4 s at 2 GHz
	int coarse_threshold = 63;               // levels with at most this many rows run in the persistent bottom kernel
	int rb_coarse_threshold = 63;            // ... with red-black SOR (MGB_RB_BOTTOM_ROWS, experiment knob)
	std::vector<XferReq> pending;            // deferred transfer requests (see flush_levels)
	int bcast_done = -1;                     // level whose result a fused leg has just broadcast from inside the kernel
	bool dead = false;                       // a ghost-row wait timed out: the ranks' version counters are out of step, the engine is unusable
	int rb_fuse_min_rows = 2047;             // red-black SOR: levels with fewer rows take the one-sweep kernels (fuse_level; MGB_RB_FUSE_MIN_ROWS)
	bool cg_fuse = true;                     // CG: direction update + operator apply + deferred x update in one pass (MGB_CG_FUSE=0: separate passes)
	bool inkernel = true;                    // fused legs push / wait for their strip-to-strip rows themselves (MGB_INKERNEL_HALO=0: separate k_xfer launches)
};
#define LAUNCHED(e) do { (e)->launches++; } while (0)
static void drop_graphs(mgb_engine *e);

// ------------------------------------------------------------------------------------------------ partition + layout
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int build_geometry(const mgb_config *cfg, std::vector<LevelGeom> &geo, int *La_out)
{
	const int L = cfg->levels, P = cfg->nranks < 1 ? 1 : cfg->nranks;
	geo.assign(L, LevelGeom());
	for (int l = 0; l < L; ++l) {
		// level sizes: n_l = (N-1)/2^l - 1 with N-1 = n_0 + 1   (ref: src/matbuild.c:64-66)
		geo[l].gni = (cfg->ni + 1) / (1 << l) - 1;
		geo[l].nj = (cfg->nj + 1) / (1 << l) - 1;
		if (geo[l].gni < 1 || geo[l].nj < 1) return fail(MGB_EINVAL, "level %d has no interior points", l);
		geo[l].pitch = ((geo[l].nj + 1 + 15) / 16) * 16;
		geo[l].coef_host.assign((size_t)geo[l].gni * MGB_COEF_STRIDE, 0.0);
	}
	int La = L;
	if (P > 1) {
		const int thr = cfg->agglomerate_below > 0 ? cfg->agglomerate_below : 511;
		for (int l = 0; l < L; ++l) if (geo[l].gni <= thr) { La = l; break; }
		if (La == 0) return fail(MGB_EINVAL, "the finest level (%d rows) is not above agglomerate_below = %d: nothing to distribute over %d ranks",
		                         geo[0].gni, thr, P);
		// base partition on the first agglomerated level (or on the coarsest level when nothing is agglomerated)
		const int lb = La < L ? La : L - 1;
		for (int r = 0; r <= P; ++r) geo[lb].rows[r] = (int)((long long)r * geo[lb].gni / P);
		for (int l = lb - 1; l >= 0; --l) {
			// the coarse grid row I sits on fine row 2I+1 (ref: src/solver.c:231-232): fine rows [2 c_r, 2 c_{r+1})
			for (int r = 0; r < P; ++r) geo[l].rows[r] = 2 * geo[l + 1].rows[r];
			geo[l].rows[P] = geo[l].gni;
		}
		for (int l = 0; l < La; ++l) {
			geo[l].dist = true;
			for (int r = 0; r < P; ++r)
				if (geo[l].rows[r + 1] - geo[l].rows[r] < 2 * MGB_GHOST_ROWS)
					return fail(MGB_EINVAL, "level %d: rank %d would hold %d rows (< %d): too many ranks for this grid, raise "
					            "agglomerate_below", l, r, geo[l].rows[r + 1] - geo[l].rows[r], 2 * MGB_GHOST_ROWS);
		}
	}
	for (int l = 0; l < L; ++l)
		if (!geo[l].dist && !(P > 1 && l == La)) { geo[l].rows[0] = 0; for (int r = 1; r <= P; ++r) geo[l].rows[r] = geo[l].gni; }
	*La_out = La;
	return MGB_OK;
}

static void build_layout(const mgb_config *cfg, const std::vector<LevelGeom> &geo, int La, int rank, Layout &y)
{
	const int L = cfg->levels;
	size_t off = 0;
	for (int l = 0; l < MGB_MAXL; ++l) for (int k = 0; k < MGB_NVEC; ++k) y.vec_off[l][k] = NOOFF;
	for (int l = 0; l < L; ++l) {
		const LevelGeom &g = geo[l];
		const bool present = g.dist || rank == 0 || l == La;       // rank != 0 keeps staging arrays of level La only
		y.ni[l] = g.dist ? g.rows[rank + 1] - g.rows[rank] : g.gni;
		y.r0[l] = g.dist ? g.rows[rank] : 0;
		y.origin[l] = (size_t)MGB_GHOST_ROWS * g.pitch + 16;
		y.alloc[l] = (size_t)(y.ni[l] + 2 * MGB_GHOST_ROWS + 1) * g.pitch + 32;
		if (!present) continue;
		const int nvec = (l == 0) ? MGB_NVEC : MGB_VEC_W + 1;      // Krylov work vectors exist on the finest level only
		for (int k = 0; k < nvec; ++k) { y.vec_off[l][k] = off; off = align_up(off + y.alloc[l] * sizeof(double), 256); }
	}
	y.flags_off = off;  off = align_up(off + sizeof(unsigned long long) * MGB_NCHAN * MGB_MAX_RANKS, 256);
	y.slots_off = off;  off = align_up(off + sizeof(double) * 2 * MGB_MAX_RANKS * RED_VALS, 256);
	y.ver_off = off;    off = align_up(off + sizeof(unsigned long long) * MGB_NCHAN, 256);
	y.ticket_off = off; off = align_up(off + sizeof(unsigned int) * MGB_NCHAN, 256);
	y.status_off = off; off = align_up(off + 256, 256);
	y.total = off;
}

// the row partition (host arithmetic only, usable without a GPU)
extern "C" int mgb_strip_rows(const mgb_config *cfg, int level, int rank, int *row0, int *row1, int *distributed)
{
	if (!cfg || level < 0 || level >= cfg->levels || rank < 0 || rank >= (cfg->nranks < 1 ? 1 : cfg->nranks))
		return fail(MGB_EINVAL, "bad argument");
	if (cfg->levels > MGB_MAXL) return fail(MGB_EINVAL, "too many levels");
	std::vector<LevelGeom> geo; int La;
	TRY(build_geometry(cfg, geo, &La));
	const bool part = geo[level].dist || ((cfg->nranks > 1) && level == La);
	if (row0) *row0 = part ? geo[level].rows[rank] : 0;
	if (row1) *row1 = part ? geo[level].rows[rank + 1] : geo[level].gni;
	if (distributed) *distributed = geo[level].dist ? 1 : 0;
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ small accessors
static LevelDev ldev(const mgb_engine *e, const Strip &s, int l)
{
	LevelDev d; d.ni = s.lev[l].ni; d.nj = e->geo[l].nj; d.pitch = e->geo[l].pitch; d.i0 = s.lev[l].r0;
	d.uniform = e->geo[l].uniform; d.rb = e->cfg.red_black_numbering ? 1 : 0; d.coef = s.lev[l].coef; d.gni = e->geo[l].gni;
	return d;
}
// view of rows [c0, c1) of a whole (agglomerated) level held in full on this strip
static LevelDev ldev_rows(const mgb_engine *e, const Strip &s, int l, int c0, int c1)
{
	LevelDev d = ldev(e, s, l); d.ni = c1 - c0; d.i0 = c0; return d;
}
static int check_vec(const mgb_engine *e, int which, int level)
{
	if (level < 0 || level >= e->L) return fail(MGB_EINVAL, "level %d out of range", level);
	if (which < 0 || which >= MGB_NVEC) return fail(MGB_EINVAL, "vector id %d out of range", which);
	if (level > 0 && which > MGB_VEC_W) return fail(MGB_EINVAL, "Krylov work vectors (ids > %d) exist on the finest level only", MGB_VEC_W);
	return MGB_OK;
}
static unsigned long long *flags_of(mgb_engine *e, int rank) { return (unsigned long long *)(e->arena_of[rank] + e->lay[rank].flags_off); }
static double *slots_of(mgb_engine *e, int rank) { return (double *)(e->arena_of[rank] + e->lay[rank].slots_off); }
static unsigned long long *ver_of(mgb_engine *e, const Strip &s) { return (unsigned long long *)(s.arena + e->lay[s.rank].ver_off); }
static unsigned int *ticket_of(mgb_engine *e, const Strip &s) { return (unsigned int *)(s.arena + e->lay[s.rank].ticket_off); }
static int *status_of(mgb_engine *e, const Strip &s) { return (int *)(s.arena + e->lay[s.rank].status_off); }
// element (0,0) of physical buffer k of level l on rank r, as addressable from this process
static double *peer_vec(mgb_engine *e, int r, int l, int k)
{
	return (double *)(e->arena_of[r] + e->lay[r].vec_off[l][k]) + e->lay[r].origin[l];
}
static int need_peers(mgb_engine *e)
{
	if (e->P > 1 && !e->connected) return fail(MGB_ESTATE, "multi-rank engine: call mgb_ipc_export / mgb_ipc_connect on every rank first");
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ lifetime
extern "C" int mgb_version(void) { return 200; }
extern "C" const char *mgb_last_error(void) { return g_err; }

static void free_csr(Csr &c) { cudaFree(c.rowptr); cudaFree(c.col); cudaFree(c.val); c = Csr(); }

extern "C" int mgb_destroy(mgb_engine *e)
{
	if (!e) return MGB_OK;
	for (auto &s : e->strips) if (s.stream) cudaStreamSynchronize(s.stream);
	for (auto &g : e->gcache) cudaGraphExecDestroy(g.exec);
	for (int r = 0; r < MGB_MAX_RANKS; ++r) if (e->peer_opened[r]) cudaIpcCloseMemHandle(e->arena_of[r]);
	for (size_t i = 0; i < e->strips.size(); ++i) {
		Strip &s = e->strips[i];
		for (auto &L : s.lev) { cudaFree(L.coef); free_csr(L.A); free_csr(L.R); free_csr(L.P); bandlu_free(L.lu); cudaFree(L.ilu); }
		cudaFree(s.wave_progress);
		cudaFree(s.stage_in); cudaFree(s.stage_out);
		cudaFree(s.arena); cudaFree(s.partial); cudaFree(s.scal); cudaFreeHost(s.scal_host); cudaFreeHost(s.status_host);
		cudaFree(s.tab_x); cudaFree(s.tab_y);
		if (s.ev0) cudaEventDestroy(s.ev0);
		if (s.ev1) cudaEventDestroy(s.ev1);
		if (s.ev_in) cudaEventDestroy(s.ev_in);
		if (s.ev_out) cudaEventDestroy(s.ev_out);
		if (s.ev_sol) cudaEventDestroy(s.ev_sol);
		if (s.copy_in) cudaStreamDestroy(s.copy_in);
		if (s.copy_out) cudaStreamDestroy(s.copy_out);
		if (s.stream && i == 0) cudaStreamDestroy(s.stream);       // emulated strips share the stream of strip 0
	}
	delete e;
	return MGB_OK;
}

// everything of mgb_create that can fail after the engine object exists (the caller destroys it on failure)
static int create_body(mgb_engine *e, const mgb_config *cfg, int P, int dev)
{
	e->cfg = *cfg; e->cfg.nranks = P;
	e->P = P; e->L = cfg->levels;
	TRY(build_geometry(&e->cfg, e->geo, &e->La));
	e->lay.resize(P);
	for (int r = 0; r < P; ++r) build_layout(&e->cfg, e->geo, e->La, r, e->lay[r]);
	const int nlocal = (P > 1 && cfg->emulate) ? P : 1;
	e->strips.resize(nlocal);
	for (int i = 0; i < nlocal; ++i) {
		Strip &s = e->strips[i];
		s.rank = nlocal > 1 ? i : cfg->rank;
		s.device = dev;
		const Layout &y = e->lay[s.rank];
		if (i == 0) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
		else s.stream = e->strips[0].stream;                       // emulation: one stream, lock step
		CU(cudaEventCreate(&s.ev0)); CU(cudaEventCreate(&s.ev1));
		CU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&s.ev_sol, cudaEventDisableTiming));
		CU(cudaStreamCreateWithFlags(&s.copy_in, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&s.copy_out, cudaStreamNonBlocking));
		CU(cudaMalloc(&s.arena, y.total));
		CU(cudaMemsetAsync(s.arena, 0, y.total, s.stream));
		e->arena_of[s.rank] = s.arena;
		s.lev.resize(e->L);
		size_t nb = 3 * (size_t)MGB_RED_MAXBLOCKS;                 // partial sums: >= the largest streaming grid of the strip
		for (int l = 0; l < e->L; ++l) {
			SLevel &S = s.lev[l];
			const LevelGeom &g = e->geo[l];
			S.ni = y.ni[l]; S.r0 = y.r0[l];
			S.present = y.vec_off[l][0] != NOOFF;
			S.active = g.dist || s.rank == 0;
			for (int k = 0; k < MGB_NVEC; ++k) {
				S.phys[k] = k;
				if (y.vec_off[l][k] == NOOFF) continue;
				S.base[k] = (double *)(s.arena + y.vec_off[l][k]);
				S.v[k] = S.base[k] + y.origin[l];
			}
			if (S.present) CU(cudaMalloc(&S.coef, sizeof(double) * g.coef_host.size()));
			const size_t blocks = (size_t)cdiv(g.pitch, MGB_SB_COLS) * (size_t)(S.ni / 2 + 1);
			if (blocks > nb) nb = blocks;
			const size_t fblocks = (size_t)(cdiv(g.pitch, FJ_VALID) + 1) * (size_t)(S.ni / 2 + 1);
			if (fblocks > nb) nb = fblocks;
		}
		s.partial_cap = nb;
		CU(cudaMalloc(&s.partial, sizeof(double) * nb));
		CU(cudaMalloc(&s.scal, sizeof(double) * 64));
		CU(cudaMemsetAsync(s.scal, 0, sizeof(double) * 64, s.stream));
		CU(cudaHostAlloc(&s.scal_host, sizeof(double) * 64, cudaHostAllocMapped));
		CU(cudaHostGetDevicePointer(&s.scal_host_dev, s.scal_host, 0));
		CU(cudaHostAlloc(&s.status_host, sizeof(int) * 4, cudaHostAllocMapped));
		s.status_host[0] = 0;
		CU(cudaHostGetDevicePointer(&s.status_host_dev, s.status_host, 0));
		CU(cudaMalloc(&s.wave_progress, sizeof(int) * 160));
		CU(cudaMalloc(&s.tab_x, sizeof(double) * (size_t)(cfg->nj + 16)));
		CU(cudaMalloc(&s.tab_y, sizeof(double) * (size_t)(cfg->ni + 16)));
	}
	CU(cudaStreamSynchronize(e->strips[0].stream));
	e->connected = (P == 1) || nlocal > 1;
	{ const char *v = getenv("MGB_INKERNEL_HALO"); if (v && v[0] == '0') e->inkernel = false; }
	{ const char *v = getenv("MGB_PDL"); if (v && v[0]) g_pdl = v[0] != '0'; }
	{ const char *v = getenv("MGB_CG_FUSE"); if (v && v[0] == '0') e->cg_fuse = false; }
	{ const char *v = getenv("MGB_RB_BOTTOM_ROWS"); if (v && v[0]) e->rb_coarse_threshold = atoi(v); }
	{ const char *v = getenv("MGB_BOTTOM_ROWS"); if (v && v[0]) e->coarse_threshold = atoi(v); }
	{ const char *v = getenv("MGB_RB_FUSE_MIN_ROWS"); if (v && v[0]) e->rb_fuse_min_rows = atoi(v); }
	return MGB_OK;
}


extern "C" int mgb_create(const mgb_config *cfg, mgb_engine **out)
{
	if (!cfg || !out) return fail(MGB_EINVAL, "null argument");
	if (cfg->levels < 1 || cfg->ni < 1 || cfg->nj < 1) return fail(MGB_EINVAL, "levels, ni, nj must be positive");
	if (cfg->levels > MGB_MAXL) return fail(MGB_EINVAL, "at most %d levels", MGB_MAXL);
	const int P = cfg->nranks < 1 ? 1 : cfg->nranks;
	if (P > MGB_MAX_RANKS) return fail(MGB_EINVAL, "at most %d ranks (one NVSwitch domain)", MGB_MAX_RANKS);
	if (cfg->rank < 0 || cfg->rank >= P) return fail(MGB_EINVAL, "rank %d out of range", cfg->rank);
	int ndev = 0;
	cudaError_t ce = cudaGetDeviceCount(&ndev);
	if (ce != cudaSuccess || ndev < 1)
		return fail(MGB_ECUDA, "no CUDA device available (%s): the B200 engine has no CPU fallback", cudaGetErrorString(ce));
	if (cfg->device >= 0) CU(cudaSetDevice(cfg->device));
	int dev = 0; CU(cudaGetDevice(&dev));
	mgb_engine *e = new mgb_engine();
	const int rc = create_body(e, cfg, P, dev);
	if (rc != MGB_OK) { mgb_destroy(e); return rc; }     // g_err keeps the message of the failing call
	*out = e;
	return MGB_OK;
}

extern "C" int mgb_ipc_export(mgb_engine *e, void *handle)
{
	if (!e || !handle) return fail(MGB_EINVAL, "null argument");
	if (e->strips.size() != 1) return fail(MGB_ESTATE, "mgb_ipc_export is for one-strip-per-process engines");
	cudaIpcMemHandle_t h;
	CU(cudaIpcGetMemHandle(&h, e->strips[0].arena));
	static_assert(sizeof(cudaIpcMemHandle_t) == MGB_IPC_HANDLE_BYTES, "IPC handle size");
	memcpy(handle, &h, sizeof h);
	return MGB_OK;
}

extern "C" int mgb_ipc_connect(mgb_engine *e, const void *handles)
{
	if (!e || !handles) return fail(MGB_EINVAL, "null argument");
	if (e->P == 1 || e->strips.size() != 1) return fail(MGB_ESTATE, "mgb_ipc_connect is for one-strip-per-process engines with nranks > 1");
	if (e->connected) return MGB_OK;
	const int me = e->strips[0].rank;
	for (int r = 0; r < e->P; ++r) {
		if (r == me) continue;
		cudaIpcMemHandle_t h;
		memcpy(&h, (const char *)handles + (size_t)r * MGB_IPC_HANDLE_BYTES, sizeof h);
		void *p = nullptr;
		CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		e->arena_of[r] = (char *)p;
		e->peer_opened[r] = true;
	}
	e->connected = true;
	return MGB_OK;
}

extern "C" int mgb_level_dims(const mgb_engine *e, int level, int *ni, int *nj)
{
	if (!e || level < 0 || level >= e->L) return fail(MGB_EINVAL, "level out of range");
	if (ni) *ni = e->geo[level].gni;
	if (nj) *nj = e->geo[level].nj;
	return MGB_OK;
}
// rows of level `level` that this process holds and computes ([0, n) on a single rank or in emulation)
extern "C" int mgb_local_rows(const mgb_engine *e, int level, int *row0, int *row1)
{
	if (!e || level < 0 || level >= e->L) return fail(MGB_EINVAL, "level out of range");
	int a = 0, b = e->geo[level].gni;
	if (e->strips.size() == 1 && e->P > 1) {
		const SLevel &S = e->strips[0].lev[level];
		if (e->geo[level].dist) { a = S.r0; b = S.r0 + S.ni; }
		else if (e->strips[0].rank != 0) { a = 0; b = 0; }
	}
	if (row0) *row0 = a;
	if (row1) *row1 = b;
	return MGB_OK;
}
extern "C" long long mgb_launch_count(const mgb_engine *e) { return e ? e->launches : 0; }
extern "C" double mgb_last_solve_ms(const mgb_engine *e) { return e ? e->last_solve_ms : 0.0; }

static int sync_all(mgb_engine *e)
{
	for (auto &s : e->strips) CU(cudaStreamSynchronize(s.stream));
	return MGB_OK;
}
// after a synchronisation: did a halo wait time out?
static int check_status(mgb_engine *e)
{
	if (e->P == 1) return MGB_OK;
	for (auto &s : e->strips) {
		CU(cudaMemcpyAsync(s.status_host, status_of(e, s), sizeof(int), cudaMemcpyDeviceToHost, s.stream));
		CU(cudaStreamSynchronize(s.stream));
		if (s.status_host[0] != 0) {
			e->dead = true;
			return fail(MGB_ECUDA, "rank %d: timed out waiting for a neighbour's ghost rows (a peer stopped or the ranks diverged)", s.rank);
		}
	}
	return MGB_OK;
}
// after any synchronisation, without a copy: the mapped host mirror of the status word (set by a timed-out wait)
static int quick_status(mgb_engine *e)
{
	if (e->P == 1) return MGB_OK;
	for (auto &s : e->strips)
		if (s.status_host && s.status_host[0] != 0) {
			e->dead = true;
			return fail(MGB_ECUDA, "rank %d: timed out waiting for a neighbour's rows (a peer stopped or the ranks diverged); the engine must be "
			            "destroyed on every rank", s.rank);
		}
	return MGB_OK;
}
static int flush_all(mgb_engine *e);
static int sync(mgb_engine *e) { TRY(flush_all(e)); TRY(sync_all(e)); return check_status(e); }

// ------------------------------------------------------------------------------------------------ operators
extern "C" int mgb_set_level_operator(mgb_engine *e, int level, const double *row_coeff)
{
	if (!e || !row_coeff) return fail(MGB_EINVAL, "null argument");
	if (level < 0 || level >= e->L) return fail(MGB_EINVAL, "level %d out of range", level);
	LevelGeom &g = e->geo[level];
	g.uniform = 1;
	for (int i = 0; i < g.gni; ++i) {
		double *c = &g.coef_host[(size_t)i * MGB_COEF_STRIDE];
		for (int k = 0; k < 5; ++k) c[k] = row_coeff[i * 5 + k];
		if (c[2] == 0.0) return fail(MGB_EINVAL, "zero diagonal on level %d grid row %d", level, i);
		c[5] = 1.0 / c[2];          // PCSetUp_Jacobi: reciprocal of the diagonal
		c[6] = 1.0 / c[2];          // MatInvertDiagonal_SeqAIJ with omega == 1
		c[7] = c[2];                // mdiag
		if (memcmp(c, &g.coef_host[0], 5 * sizeof(double)) != 0) g.uniform = 0;
	}
	if (g.uniform) {
		// power-of-two operator (every uniform 2^k - 1 grid): lets the fused kernel factor the coefficient out exactly
		const double *c = &g.coef_host[0];
		int ex = 0;
		if (c[0] == c[1] && c[0] == c[3] && c[0] == c[4] && c[2] == -4.0 * c[0] && c[0] > 0.0 && frexp(c[0], &ex) == 0.5)
			g.uniform = 2;
	}
	g.coef_set = true;
	e->sor_omega = -1.0;          // idiag of this level is now 1/diag whatever omega the others hold: recompute all on the next SOR use
	e->csr_built = false;
	drop_graphs(e);
	for (auto &s : e->strips) {
		SLevel &S = s.lev[level];
		bandlu_free(S.lu);
		S.ilu_valid = false;
		if (!S.present) continue;
		CU(cudaMemcpyAsync(S.coef, g.coef_host.data(), sizeof(double) * g.coef_host.size(), cudaMemcpyHostToDevice, s.stream));
		CU(cudaStreamSynchronize(s.stream));
	}
	return MGB_OK;
}

// idiag = omega / diag for the SOR kernels (MatInvertDiagonal_SeqAIJ: 1/d when omega == 1 and fshift == 0)
static int set_sor_omega(mgb_engine *e, double omega)
{
	if (e->sor_omega == omega) return MGB_OK;
	for (int l = 0; l < e->L; ++l) {
		LevelGeom &g = e->geo[l];
		for (int i = 0; i < g.gni; ++i) {
			double *c = &g.coef_host[(size_t)i * MGB_COEF_STRIDE];
			c[6] = (omega == 1.0) ? 1.0 / c[2] : omega / (0.0 + c[2]);
		}
		for (auto &s : e->strips) {
			if (!s.lev[l].present) continue;
			CU(cudaMemcpyAsync(s.lev[l].coef, g.coef_host.data(), sizeof(double) * g.coef_host.size(), cudaMemcpyHostToDevice, s.stream));
			CU(cudaStreamSynchronize(s.stream));
		}
	}
	e->sor_omega = omega;
	return MGB_OK;
}

extern "C" int mgb_set_transfer(mgb_engine *e, const double res3[9], const double pro3[9])
{
	if (!e || !res3 || !pro3) return fail(MGB_EINVAL, "null argument");
	for (int k = 0; k < 9; ++k) {
		if (res3[k] == 0.0 || pro3[k] == 0.0)
			return fail(MGB_EINVAL, "zero transfer weight: the reference drops such entries from res/pro (src/solver.c:1086), not supported");
		e->R3.w[k] = res3[k]; e->P3.w[k] = pro3[k];
	}
	e->transfer_pow2 = true;
	for (int k = 0; k < 9; ++k) {
		int ex = 0;
		if (!(res3[k] > 0.0 && frexp(res3[k], &ex) == 0.5 && pro3[k] > 0.0 && frexp(pro3[k], &ex) == 0.5)) e->transfer_pow2 = false;
	}
	e->transfer_set = true;
	e->csr_built = false;
	drop_graphs(e);
	return MGB_OK;
}

static int require_ops(mgb_engine *e, bool transfer)
{
	if (!e) return fail(MGB_EINVAL, "null engine");
	for (int l = 0; l < e->L; ++l)
		if (!e->geo[l].coef_set) return fail(MGB_ESTATE, "mgb_set_level_operator was not called for level %d", l);
	if (transfer && e->L > 1 && !e->transfer_set) return fail(MGB_ESTATE, "mgb_set_transfer was not called");
	return need_peers(e);
}

// ------------------------------------------------------------------------------------------------ strip-to-strip transfers
static int xfer_blocks(unsigned long long total2)
{
	long long b = (long long)((total2 + MGB_XFER_THREADS * 4 - 1) / (MGB_XFER_THREADS * 4));
	return (int)(b < 1 ? 1 : (b > 64 ? 64 : b));
}
// one process per strip: push + wait in one launch.  Emulation (all strips in this process, one GPU): the pushes of
// every strip first, then the waits, so that no kernel ever waits for a kernel queued behind it.
static int xfer_run(mgb_engine *e, std::vector<XferArgs> &args, const std::vector<unsigned long long> &tot, int chan)
{
	for (size_t i = 0; i < e->strips.size(); ++i) {
		Strip &s = e->strips[i]; XferArgs &a = args[i];
		a.ver = ver_of(e, s) + chan; a.ticket = ticket_of(e, s) + chan; a.status = status_of(e, s); a.status_host = s.status_host_dev; a.spin_limit = e->spin_limit;
		a.do_push = 1; a.do_wait = e->strips.size() == 1 ? 1 : 0;
		k_xfer<<<xfer_blocks(tot[i]), MGB_XFER_THREADS, 0, s.stream>>>(a);
		LAUNCHED(e); KCHECK();
	}
	if (e->strips.size() == 1) return MGB_OK;
	for (size_t i = 0; i < e->strips.size(); ++i) {
		Strip &s = e->strips[i]; XferArgs a = args[i];
		a.do_push = 0; a.do_wait = 1;
		k_xfer<<<1, 32, 0, s.stream>>>(a);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}

// Transfer requests are DEFERRED: halo / gather_rows / bcast_rows only record what has to move; flush_levels() turns
// every pending request that concerns the given levels into ONE k_xfer launch per strip (rows of several vectors and
// levels travel together, one flag, one wait).  Compute helpers flush the levels they read before launching, so e.g.
// the ghost rows of u[l] written by the down leg travel together with those of u[l+1] right before the up leg of
// level l, and a restricted right-hand side travels alone right before the next level's down leg.
static void add_dst(XferArgs &a, const double *src, double *dst, unsigned long long cnt2, unsigned long long *peer_flag_base)
{
	const int d = a.ndst++;
	a.src[d] = src; a.dst[d] = dst; a.cnt2[d] = cnt2; a.peer_flag[d] = peer_flag_base;   // flag base of the destination rank; channel offset added at flush
}

// ghost rows of vector `which` on distributed level l: `depth` boundary rows go to each neighbour
static int halo(mgb_engine *e, int l, int which, int depth)
{
	if (e->P == 1 || !e->geo[l].dist) return MGB_OK;
	e->pending.push_back({REQ_HALO, l, e->strips[0].lev[l].phys[which], depth});
	return MGB_OK;
}
// rows [rows[r], rows[r+1]) of vector `which` on the first agglomerated level: every rank -> rank 0
static int gather_rows(mgb_engine *e, int l, int which)
{
	if (e->P == 1) return MGB_OK;
	e->pending.push_back({REQ_GATHER, l, e->strips[0].lev[l].phys[which], 0});
	return MGB_OK;
}
// rows [rows[r]-3, rows[r+1]+3) of vector `which` on the first agglomerated level: rank 0 -> every rank
static int bcast_rows(mgb_engine *e, int l, int which)
{
	if (e->P == 1) return MGB_OK;
	e->pending.push_back({REQ_BCAST, l, e->strips[0].lev[l].phys[which], 0});
	return MGB_OK;
}

// launch the pending requests with lo <= level <= hi as one merged exchange
static int flush_levels(mgb_engine *e, int lo, int hi)
{
	if (e->pending.empty()) return MGB_OK;
	std::vector<XferReq> take, keep;
	for (auto &q : e->pending) (q.level >= lo && q.level <= hi ? take : keep).push_back(q);
	if (take.empty()) return MGB_OK;
	e->pending.swap(keep);
	int chan = MGB_NCHAN;
	for (auto &q : take) {
		const int c = q.type == REQ_HALO ? CH_HALO(q.level, q.phys) : (q.type == REQ_GATHER ? CH_GATHER(q.level) : CH_BCAST(q.level));
		if (c < chan) chan = c;
	}
	std::vector<XferArgs> args(e->strips.size());
	std::vector<unsigned long long> tot(e->strips.size(), 0);
	for (size_t i = 0; i < e->strips.size(); ++i) {
		Strip &s = e->strips[i];
		XferArgs &a = args[i]; memset(&a, 0, sizeof a);
		const int r = s.rank;
		unsigned wait_mask = 0;
		for (auto &q : take) {
			const int l = q.level, k = q.phys;
			const size_t pitch = e->geo[l].pitch;
			double *mine = (double *)(s.arena + e->lay[r].vec_off[l][k]) + e->lay[r].origin[l];
			if (a.ndst + 8 > MGB_XFER_MAX) return fail(MGB_EINVAL, "too many transfers merged into one exchange");
			if (q.type == REQ_HALO) {
				const unsigned long long cnt2 = (unsigned long long)q.depth * pitch / 2;
				const int ni = e->lay[r].ni[l];
				if (r > 0) {                               // my first rows -> the lower ghost rows of rank r-1
					add_dst(a, mine, peer_vec(e, r - 1, l, k) + (size_t)e->lay[r - 1].ni[l] * pitch, cnt2, flags_of(e, r - 1));
					wait_mask |= 1u << (r - 1);
				}
				if (r < e->P - 1) {                        // my last rows -> the upper ghost rows of rank r+1
					add_dst(a, mine + (size_t)(ni - q.depth) * pitch, peer_vec(e, r + 1, l, k) - (size_t)q.depth * pitch, cnt2, flags_of(e, r + 1));
					wait_mask |= 1u << (r + 1);
				}
			} else if (q.type == REQ_GATHER) {
				if (r != 0) {
					const int c0 = e->geo[l].rows[r], c1 = e->geo[l].rows[r + 1];
					add_dst(a, mine + (size_t)c0 * pitch, peer_vec(e, 0, l, k) + (size_t)c0 * pitch, (unsigned long long)(c1 - c0) * pitch / 2, flags_of(e, 0));
				} else wait_mask |= ((1u << e->P) - 1u) & ~1u;
			} else {
				if (r == 0) {
					for (int t = 1; t < e->P; ++t) {
						int c0 = e->geo[l].rows[t] - 3, c1 = e->geo[l].rows[t + 1] + 3;   // the fused up leg reads 3 coarse ghost rows
						if (c0 < 0) c0 = 0;
						if (c1 > e->geo[l].gni) c1 = e->geo[l].gni;
						add_dst(a, mine + (size_t)c0 * pitch, peer_vec(e, t, l, k) + (size_t)c0 * pitch, (unsigned long long)(c1 - c0) * pitch / 2, flags_of(e, t));
					}
				} else wait_mask |= 1u;
			}
		}
		// every destination rank gets the flag of this channel raised once per entry (idempotent); waits on the union of sources
		for (int d = 0; d < a.ndst; ++d) { a.peer_flag[d] += (size_t)chan * MGB_MAX_RANKS + r; tot[i] += a.cnt2[d]; }
		for (int q = 0; q < e->P; ++q)
			if (wait_mask & (1u << q)) a.wait_flag[a.nwait++] = flags_of(e, r) + (size_t)chan * MGB_MAX_RANKS + q;
	}
	return xfer_run(e, args, tot, chan);
}
static int flush_all(mgb_engine *e) { return flush_levels(e, 0, MGB_MAXL); }

// scal[SC_LOCAL .. SC_LOCAL+nvals) of every rank, summed in rank order into scal[slot ..) of every rank
static int allreduce(mgb_engine *e, int nvals, int slot, int take_sqrt)
{
	const int chan = CH_REDUCE;
	std::vector<XferArgs> args(e->strips.size());
	std::vector<unsigned long long> tot(e->strips.size(), 1);
	for (size_t i = 0; i < e->strips.size(); ++i) {
		Strip &s = e->strips[i];
		XferArgs &a = args[i]; memset(&a, 0, sizeof a);
		const int r = s.rank;
		for (int q = 0; q < e->P; ++q) {
			const int d = a.ndst++;
			a.src[d] = s.scal + SC_LOCAL;
			a.dst[d] = slots_of(e, q) + (size_t)r * RED_VALS;            // + parity offset, added in the kernel
			a.cnt2[d] = RED_VALS / 2;
			a.peer_flag[d] = flags_of(e, q) + (size_t)chan * MGB_MAX_RANKS + r;
			a.wait_flag[a.nwait++] = flags_of(e, r) + (size_t)chan * MGB_MAX_RANKS + q;
		}
		a.parity_stride = (unsigned long long)MGB_MAX_RANKS * RED_VALS;
	}
	TRY(xfer_run(e, args, tot, chan));
	for (auto &s : e->strips) {
		k_reduce_ranks<<<1, 32, 0, s.stream>>>(slots_of(e, s.rank), (unsigned long long)MGB_MAX_RANKS * RED_VALS, ver_of(e, s) + chan,
		                                        e->P, nvals, s.scal, slot, take_sqrt);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ CSR (single rank)
static int alloc_csr(Csr &c, int m, int n, long long nnz)
{
	if (c.rowptr && c.m == m && c.n == n && c.nnz == nnz) return MGB_OK;   // re-assembly into the same arrays
	free_csr(c);
	if (nnz > 2147483647LL) return fail(MGB_EINVAL, "nnz %lld overflows the 32-bit PetscInt of the reference", nnz);
	c.m = m; c.n = n; c.nnz = nnz;
	CU(cudaMalloc(&c.rowptr, sizeof(int) * ((size_t)m + 1)));
	CU(cudaMalloc(&c.col, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
	CU(cudaMalloc(&c.val, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
	return MGB_OK;
}

static long long host_touch_prefix(int t, int nc)
{
	long long s = 0;
	for (int K = 0; K < nc; ++K) { int d = t - 2 * K; s += d < 0 ? 0 : (d > 3 ? 3 : d); }
	return s;
}

// A[l], res[l], pro[l] in CSR.  With row strips every rank assembles the rows it owns -- the grid rows of its strip for
// A and pro, its coarse rows for res -- with local row pointers and GLOBAL column indices, i.e. the local rows of the
// reference's MPIAIJ matrices (ref: src/solver.c:218 "for row in [ranges[rank], ranges[rank+1])", :502); agglomerated
// levels are assembled whole on rank 0.
extern "C" int mgb_assemble_csr(mgb_engine *e)
{
	TRY(require_ops(e, true));
	if (e->cfg.red_black_numbering)
		return fail(MGB_EINVAL, "CSR assembly is offered for the reference's natural numbering only (-map 0,1,2), not the -map 3 extension");
	for (auto &s : e->strips) {
		const int r = s.rank;
		for (int l = 0; l < e->L; ++l) {
			SLevel &S = s.lev[l]; const LevelGeom &g = e->geo[l];
			if (!(S.present && S.active)) continue;
			const long long NG = (long long)g.gni * g.nj;
			if (NG > 2147483647LL) return fail(MGB_EINVAL, "level %d has more rows than a 32-bit PetscInt holds", l);
			const int i0 = S.r0, nloc = S.ni;
			const long long N = (long long)nloc * g.nj;
			// entries of the local rows: 5 per row minus the dropped neighbours (first / last grid row, first / last column)
			const long long nnz = 5 * N - 2LL * nloc - (i0 == 0 ? g.nj : 0) - (i0 + nloc == g.gni ? g.nj : 0);
			TRY(alloc_csr(S.A, (int)N, (int)NG, nnz));
			dim3 gr(cdiv(g.nj, 256), nloc);
			k_csr_A<<<gr, 256, 0, s.stream>>>(S.A.rowptr, S.A.col, S.A.val, g.gni, g.nj, S.coef, i0, nloc);
			LAUNCHED(e); KCHECK();
			if (l + 1 < e->L) {
				const LevelGeom &c = e->geo[l + 1];
				// coarse rows this rank produces in the restriction from level l
				int I0 = 0, ncl = c.gni;
				if (g.dist) { I0 = c.rows[r]; ncl = c.rows[r + 1] - c.rows[r]; }
				const long long NC = (long long)ncl * c.nj;
				TRY(alloc_csr(S.R, (int)NC, (int)NG, 9 * NC));
				if (ncl > 0) {
					dim3 grr(cdiv(c.nj, 256), ncl);
					k_csr_R<<<grr, 256, 0, s.stream>>>(S.R.rowptr, S.R.col, S.R.val, ncl, c.nj, g.nj, e->R3, I0);
					LAUNCHED(e); KCHECK();
				}
				const long long pnnz = (host_touch_prefix(i0 + nloc, c.gni) - host_touch_prefix(i0, c.gni)) * host_touch_prefix(g.nj, c.nj);
				TRY(alloc_csr(S.P, (int)N, (int)((long long)c.gni * c.nj), pnnz));
				k_csr_P<<<gr, 256, 0, s.stream>>>(S.P.rowptr, S.P.col, S.P.val, g.gni, g.nj, c.gni, c.nj, e->P3, i0, nloc);
				LAUNCHED(e); KCHECK();
			}
		}
		CU(cudaStreamSynchronize(s.stream));
	}
	e->csr_built = true;
	return MGB_OK;
}

// matrix `which` of level `level` as held by rank `rank` (-1: the first local strip)
static int pick_csr(const mgb_engine *e, int rank, int which, int level, const Csr **out, int *row0)
{
	if (!e) return fail(MGB_EINVAL, "null engine");
	if (!e->csr_built) return fail(MGB_ESTATE, "mgb_assemble_csr was not called");
	if (level < 0 || level >= e->L) return fail(MGB_EINVAL, "level %d out of range", level);
	const Strip *sp = &e->strips[0];
	if (rank >= 0) {
		sp = nullptr;
		for (auto &s : e->strips) if (s.rank == rank) sp = &s;
		if (!sp) return fail(MGB_EINVAL, "rank %d is not held by this process", rank);
	}
	const SLevel &S = sp->lev[level];
	if (!(S.present && S.active)) return fail(MGB_EINVAL, "rank %d holds no rows of level %d", sp->rank, level);
	int r0 = S.r0 * e->geo[level].nj;
	if (which == MGB_MAT_A) *out = &S.A;
	else if (level + 1 >= e->L) return fail(MGB_EINVAL, "no transfer operator below the coarsest level");
	else if (which == MGB_MAT_RES) { *out = &S.R; r0 = (e->geo[level].dist ? e->geo[level + 1].rows[sp->rank] : 0) * e->geo[level + 1].nj; }
	else if (which == MGB_MAT_PRO) *out = &S.P;
	else return fail(MGB_EINVAL, "matrix id %d out of range", which);
	if (row0) *row0 = r0;
	return MGB_OK;
}

extern "C" int mgb_csr_dims_rank(const mgb_engine *e, int rank, int which, int level, int *m, int *n, long long *nnz, int *row0)
{
	const Csr *c = nullptr; TRY(pick_csr(e, rank, which, level, &c, row0));
	if (m) *m = c->m;
	if (n) *n = c->n;
	if (nnz) *nnz = c->nnz;
	return MGB_OK;
}
extern "C" int mgb_csr_get_rank(const mgb_engine *e, int rank, int which, int level, int *rowptr, int *col, double *val)
{
	const Csr *c = nullptr; TRY(pick_csr(e, rank, which, level, &c, nullptr));
	if (rowptr) CU(cudaMemcpy(rowptr, c->rowptr, sizeof(int) * ((size_t)c->m + 1), cudaMemcpyDeviceToHost));
	if (col && c->nnz) CU(cudaMemcpy(col, c->col, sizeof(int) * (size_t)c->nnz, cudaMemcpyDeviceToHost));
	if (val && c->nnz) CU(cudaMemcpy(val, c->val, sizeof(double) * (size_t)c->nnz, cudaMemcpyDeviceToHost));
	return MGB_OK;
}
extern "C" int mgb_csr_dims(const mgb_engine *e, int which, int level, int *m, int *n, long long *nnz)
{
	return mgb_csr_dims_rank(e, -1, which, level, m, n, nnz, nullptr);
}
extern "C" int mgb_csr_get(const mgb_engine *e, int which, int level, int *rowptr, int *col, double *val)
{
	return mgb_csr_get_rank(e, -1, which, level, rowptr, col, val);
}

// ------------------------------------------------------------------------------------------------ vectors
// rows of a level vector <- dense host rows (host points at this strip's first row).  Large vectors go through a
// dense staging buffer + repack kernel; `st` is the stream everything is enqueued on.
#define STAGE_MIN_BYTES (4u << 20)
static int stage_alloc(mgb_engine *e, Strip &s)
{
	const size_t need = (size_t)s.lev[0].ni * e->geo[0].nj;
	if (s.stage_cap >= need) return MGB_OK;
	cudaFree(s.stage_in); cudaFree(s.stage_out); s.stage_in = s.stage_out = nullptr; s.stage_cap = 0;
	CU(cudaMalloc(&s.stage_in, need * sizeof(double)));
	CU(cudaMalloc(&s.stage_out, need * sizeof(double)));
	s.stage_cap = need;
	return MGB_OK;
}
static int put_rows(mgb_engine *e, Strip &s, int level, int which, const double *host, cudaStream_t st)
{
	const LevelGeom &g = e->geo[level]; SLevel &S = s.lev[level];
	const size_t bytes = (size_t)S.ni * g.nj * sizeof(double);
	if (level != 0 || bytes < STAGE_MIN_BYTES) {
		CU(cudaMemcpy2DAsync(S.v[which], sizeof(double) * g.pitch, host, sizeof(double) * g.nj, sizeof(double) * g.nj, S.ni,
		                     cudaMemcpyHostToDevice, st));
		return MGB_OK;
	}
	TRY(stage_alloc(e, s));
	CU(cudaMemcpyAsync(s.stage_in, host, bytes, cudaMemcpyHostToDevice, st));
	k_unpack_rows<<<148 * 8, 256, 0, st>>>(S.v[which], s.stage_in, S.ni, g.nj, g.pitch);
	LAUNCHED(e); KCHECK();
	return MGB_OK;
}
// dense host rows <- rows of a level vector; the pack kernel runs on `kst`, the PCIe copy on `cst` (after `ev`)
static int get_rows(mgb_engine *e, Strip &s, int level, int which, double *host, cudaStream_t kst, cudaStream_t cst, cudaEvent_t ev)
{
	const LevelGeom &g = e->geo[level]; SLevel &S = s.lev[level];
	const size_t bytes = (size_t)S.ni * g.nj * sizeof(double);
	if (level != 0 || bytes < STAGE_MIN_BYTES) {
		if (cst != kst) { CU(cudaEventRecord(ev, kst)); CU(cudaStreamWaitEvent(cst, ev, 0)); }
		CU(cudaMemcpy2DAsync(host, sizeof(double) * g.nj, S.v[which], sizeof(double) * g.pitch, sizeof(double) * g.nj, S.ni,
		                     cudaMemcpyDeviceToHost, cst));
		return MGB_OK;
	}
	TRY(stage_alloc(e, s));
	k_pack_rows<<<148 * 8, 256, 0, kst>>>(s.stage_out, S.v[which], S.ni, g.nj, g.pitch);
	LAUNCHED(e); KCHECK();
	if (cst != kst) { CU(cudaEventRecord(ev, kst)); CU(cudaStreamWaitEvent(cst, ev, 0)); }
	CU(cudaMemcpyAsync(host, s.stage_out, bytes, cudaMemcpyDeviceToHost, cst));
	return MGB_OK;
}

// Host vectors are whole-grid arrays (gni x nj, natural order).  Every local strip moves its own rows; in a
// one-strip-per-process run the other rows of the host array are neither read nor written.
extern "C" int mgb_vec_set(mgb_engine *e, int which, int level, const double *host)
{
	if (!e || !host) return fail(MGB_EINVAL, "null argument");
	TRY(check_vec(e, which, level));
	const LevelGeom &g = e->geo[level];
	for (auto &s : e->strips) {
		SLevel &S = s.lev[level];
		if (!S.present || !S.active) continue;
		TRY(put_rows(e, s, level, which, host + (size_t)S.r0 * g.nj, s.stream));
	}
	return sync_all(e);
}
extern "C" int mgb_vec_get(mgb_engine *e, int which, int level, double *host)
{
	if (!e || !host) return fail(MGB_EINVAL, "null argument");
	TRY(check_vec(e, which, level));
	const LevelGeom &g = e->geo[level];
	for (auto &s : e->strips) {
		SLevel &S = s.lev[level];
		if (!S.present || !S.active) continue;
		TRY(get_rows(e, s, level, which, host + (size_t)S.r0 * g.nj, s.stream, s.stream, s.ev_sol));
	}
	return sync_all(e);
}
// zero the own rows AND the ghost rows (the neighbours zero theirs in the same step, so no exchange is needed)
static int vec_zero(mgb_engine *e, int which, int level)
{
	TRY(check_vec(e, which, level));
	const LevelGeom &g = e->geo[level];
	for (auto &s : e->strips) {
		SLevel &S = s.lev[level];
		if (!S.present || !S.active) continue;
		const size_t n2 = (size_t)(S.ni + 2 * MGB_GHOST_ROWS) * g.pitch / 2;
		const int blocks = (int)((n2 + 255) / 256 < 2368 ? (n2 + 255) / 256 : 2368);
		klaunch(k_axpy<3>, dim3(blocks), dim3(256), 0, s.stream, S.v[which] - (size_t)MGB_GHOST_ROWS * g.pitch, nullptr, n2, 0.0, nullptr, 0.0);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}
extern "C" int mgb_vec_zero(mgb_engine *e, int which, int level)
{
	if (!e) return fail(MGB_EINVAL, "null engine");
	TRY(vec_zero(e, which, level));
	return sync(e);
}
extern "C" int mgb_set_rhs(mgb_engine *e, const double *b0) { return mgb_vec_set(e, MGB_VEC_B, 0, b0); }
extern "C" int mgb_get_solution(mgb_engine *e, double *u0) { return mgb_vec_get(e, MGB_VEC_U, 0, u0); }

extern "C" int mgb_set_rhs_separable(mgb_engine *e, const double *gx, const double *gy)
{
	if (!e || !gx || !gy) return fail(MGB_EINVAL, "null argument");
	const LevelGeom &g = e->geo[0];
	for (auto &s : e->strips) {
		SLevel &S = s.lev[0];
		CU(cudaMemcpyAsync(s.tab_x, gx, sizeof(double) * g.nj, cudaMemcpyHostToDevice, s.stream));
		CU(cudaMemcpyAsync(s.tab_y, gy, sizeof(double) * g.gni, cudaMemcpyHostToDevice, s.stream));
		dim3 gr(cdiv(g.pitch, 256), S.ni);
		k_outer<<<gr, 256, 0, s.stream>>>(S.v[MGB_VEC_B], s.tab_x, s.tab_y, ldev(e, s, 0));
		LAUNCHED(e); KCHECK();
	}
	return sync_all(e);
}

// ------------------------------------------------------------------------------------------------ launch helpers
// All helpers enqueue on the strips' streams and do not synchronise.  They loop over the local strips that compute
// on the level (every strip of a distributed level; rank 0 only on an agglomerated one).
static inline bool computes(const Strip &s, int l) { return s.lev[l].present && s.lev[l].active; }

// rows per block of the streaming kernels: 32 on large levels, fewer on small ones so that the grid still fills
// the 148 SMs (>= ~1184 blocks when the level allows it)
static int pick_ry(const LevelGeom &g, int ni)
{
	const long long cb = cdiv(g.pitch, MGB_SB_COLS);
	long long ry = (long long)ni * cb / 1184;
	if (ry > 32) ry = 32;
	if (ry < 2) ry = 2;
	return (int)ry;
}
static dim3 stream_grid(const LevelGeom &g, int ni, int ry) { return dim3(cdiv(g.pitch, MGB_SB_COLS), cdiv(ni, ry)); }

static int k_apply(mgb_engine *e, int l, int xv, int yv)
{
	TRY(flush_levels(e, l, l));
	const LevelGeom &g = e->geo[l];
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const int ry = pick_ry(g, S.ni);
		klaunch(k_stream5<ST_APPLY>, dim3(stream_grid(g, S.ni, ry)), dim3(MGB_SB_THREADS), 0, s.stream, S.v[xv], nullptr, S.v[yv], ldev(e, s, l), 0.0, nullptr, ry);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}
static int k_residual(mgb_engine *e, int l, int xv, int bv, int rv)
{
	TRY(flush_levels(e, l, l));
	const LevelGeom &g = e->geo[l];
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const int ry = pick_ry(g, S.ni);
		klaunch(k_stream5<ST_RESID>, dim3(stream_grid(g, S.ni, ry)), dim3(MGB_SB_THREADS), 0, s.stream, S.v[xv], S.v[bv], S.v[rv], ldev(e, s, l), 0.0, nullptr, ry);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}
// final stage of a reduction in one launch (k_reduce_tail): local partials -> scal[slot] (+ its mapped host mirror), over all
// ranks of a distributed level
static int reduce_tail(mgb_engine *e, int l, const std::vector<int> &nblocks, int slot, int take_sqrt, int post_op = TAIL_NONE)
{
	const bool global = e->P > 1 && e->geo[l].dist;
	const int chan = CH_REDUCE;
	std::vector<TailArgs> args;
	size_t i = 0;
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		TailArgs a; memset(&a, 0, sizeof a);
		a.partial = s.partial; a.n = nblocks[i++]; a.scal = s.scal; a.slot = slot; a.take_sqrt = take_sqrt; a.host = s.scal_host_dev;
		a.nranks = global ? e->P : 1; a.do_push = 1; a.do_wait = 1; a.post_op = post_op;
		if (global) {
			const int r = s.rank;
			for (int q = 0; q < e->P; ++q) {
				a.slot_dst[q] = slots_of(e, q) + (size_t)r * RED_VALS;
				a.peer_flag[q] = flags_of(e, q) + (size_t)chan * MGB_MAX_RANKS + r;
				a.wait_flag[q] = flags_of(e, r) + (size_t)chan * MGB_MAX_RANKS + q;
			}
			a.my_slots = slots_of(e, r); a.ver = ver_of(e, s) + chan; a.parity_stride = (unsigned long long)MGB_MAX_RANKS * RED_VALS;
			a.status = status_of(e, s); a.status_host = s.status_host_dev; a.spin_limit = e->spin_limit;
			if (e->strips.size() > 1) a.do_wait = 0;               // emulation: every strip pushes first ...
		}
		klaunch(k_reduce_tail, dim3(1), dim3(1024), 0, s.stream, a);
		LAUNCHED(e); KCHECK();
		args.push_back(a);
	}
	if (global && e->strips.size() > 1)
		for (size_t k = 0; k < args.size(); ++k) {                 // ... then every strip waits and sums
			TailArgs a = args[k]; a.do_push = 0; a.do_wait = 1;
			klaunch(k_reduce_tail, dim3(1), dim3(32), 0, e->strips[k].stream, a);
			LAUNCHED(e); KCHECK();
		}
	return MGB_OK;
}
// scal[slot] = || b - A x ||_2
static int k_resnorm(mgb_engine *e, int l, int xv, int bv, int slot)
{
	TRY(flush_levels(e, l, l));
	const LevelGeom &g = e->geo[l];
	std::vector<int> nb;
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const int ry = pick_ry(g, S.ni);
		const dim3 gr = stream_grid(g, S.ni, ry);
		if ((size_t)gr.x * gr.y > s.partial_cap) return fail(MGB_EINVAL, "grid too large for the partial-sum buffer");
		klaunch(k_stream5<ST_RESNORM>, dim3(gr), dim3(MGB_SB_THREADS), 0, s.stream, S.v[xv], S.v[bv], nullptr, ldev(e, s, l), 0.0, s.partial, ry);
		LAUNCHED(e); KCHECK();
		nb.push_back((int)(gr.x * gr.y));
	}
	return reduce_tail(e, l, nb, slot, 1);
}
// scal[slot] = sqrt?(sum x*y) (yv < 0: sum x*x)
static int k_reduce(mgb_engine *e, int l, int xv, int yv, int slot, int take_sqrt, int post_op = TAIL_NONE)
{
	const LevelGeom &g = e->geo[l];
	std::vector<int> nb;
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const size_t n2 = (size_t)S.ni * g.pitch / 2;
		size_t want = (n2 + MGB_RED_THREADS * 4 - 1) / (MGB_RED_THREADS * 4);
		const int blocks = (int)(want < 1 ? 1 : (want > MGB_RED_MAXBLOCKS ? MGB_RED_MAXBLOCKS : want));
		if (yv >= 0) klaunch(k_reduce1<1>, dim3(blocks), dim3(MGB_RED_THREADS), 0, s.stream, S.v[xv], S.v[yv], n2, s.partial);
		else         klaunch(k_reduce1<0>, dim3(blocks), dim3(MGB_RED_THREADS), 0, s.stream, S.v[xv], nullptr, n2, s.partial);
		LAUNCHED(e); KCHECK();
		nb.push_back(blocks);
	}
	return reduce_tail(e, l, nb, slot, take_sqrt, post_op);
}
template <int KIND>
static int k_vecop(mgb_engine *e, int l, int yv, int xv, double alpha, int alpha_slot = -1)
{
	const LevelGeom &g = e->geo[l];
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const size_t n2 = (size_t)S.ni * g.pitch / 2;
		size_t want = (n2 + 255) / 256;
		const int blocks = (int)(want < 1 ? 1 : (want > 148 * 32 ? 148 * 32 : want));
		klaunch(k_axpy<KIND>, dim3(blocks), dim3(256), 0, s.stream, S.v[yv], S.v[xv], n2, alpha, alpha_slot >= 0 ? s.scal + alpha_slot : nullptr, 1.0);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}
// y = A x ; scal[slot] = x . y in one pass (the w = A p, p'w pair of CG)
static int k_apply_dot(mgb_engine *e, int l, int xv, int yv, int slot, int post_op = TAIL_NONE)
{
	TRY(flush_levels(e, l, l));
	const LevelGeom &g = e->geo[l];
	std::vector<int> nb;
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const int ry = pick_ry(g, S.ni);
		const dim3 gr = stream_grid(g, S.ni, ry);
		if ((size_t)gr.x * gr.y > s.partial_cap) return fail(MGB_EINVAL, "grid too large for the partial-sum buffer");
		klaunch(k_stream5<ST_APPLYDOT>, dim3(gr), dim3(MGB_SB_THREADS), 0, s.stream, S.v[xv], nullptr, S.v[yv], ldev(e, s, l), 0.0, s.partial, ry);
		LAUNCHED(e); KCHECK();
		nb.push_back((int)(gr.x * gr.y));
	}
	return reduce_tail(e, l, nb, slot, 0, post_op);
}
static void swap_vec(mgb_engine *e, int l, int a, int b);
// the CG direction step in one pass (mgb_stencil.cuh: k_cg_pstep): x += a_prev p ; p_new = z + b p (into the buffer of `nv`, then
// p and nv swap) ; w = A p_new ; scal[slot] = p_new . w [+ TAIL_DPI].  Needs ghost rows of z and p to depth 1; leaves those of
// the new p valid to depth 1 without an exchange.
static int k_cg_dir(mgb_engine *e, int l, int zv, int pv, int nv, int wv, int xv, int slot, int post_op)
{
	TRY(flush_levels(e, l, l));
	const LevelGeom &g = e->geo[l];
	std::vector<int> nb;
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const int ry = pick_ry(g, S.ni);
		const dim3 gr = stream_grid(g, S.ni, ry);
		if ((size_t)gr.x * gr.y > s.partial_cap) return fail(MGB_EINVAL, "grid too large for the partial-sum buffer");
		klaunch(k_cg_pstep, dim3(gr), dim3(MGB_SB_THREADS), 0, s.stream, S.v[zv], S.v[pv], S.v[nv], S.v[wv], S.v[xv], ldev(e, s, l),
		        s.scal + SC_RATIO, s.scal + SC_ALPHA, s.partial, ry);
		LAUNCHED(e); KCHECK();
		nb.push_back((int)(gr.x * gr.y));
	}
	swap_vec(e, l, pv, nv);
	return reduce_tail(e, l, nb, slot, 0, post_op);
}
// x += a p ; r -= a w ; scal[slot] = ||r||_2 in one pass (the CG update, mgb_blas.cuh: k_cg_update); xv < 0: r and ||r|| only
static int k_cg_step(mgb_engine *e, int l, int xv, int pv, int rv, int wv, double a, int slot, int a_slot = -1)
{
	const LevelGeom &g = e->geo[l];
	std::vector<int> nb;
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const size_t n2 = (size_t)S.ni * g.pitch / 2;
		size_t want = (n2 + MGB_RED_THREADS * 4 - 1) / (MGB_RED_THREADS * 4);
		const int blocks = (int)(want < 1 ? 1 : (want > MGB_RED_MAXBLOCKS ? MGB_RED_MAXBLOCKS : want));
		klaunch(k_cg_update, dim3(blocks), dim3(MGB_RED_THREADS), 0, s.stream, xv >= 0 ? S.v[xv] : nullptr, S.v[pv], S.v[rv], S.v[wv], n2, a, s.partial, a_slot >= 0 ? s.scal + a_slot : nullptr);
		LAUNCHED(e); KCHECK();
		nb.push_back(blocks);
	}
	return reduce_tail(e, l, nb, slot, 1);
}
// scal[first .. first+count) of the first local strip -> its pinned mirror (every rank holds identical values)
static int read_scalars(mgb_engine *e, int first, int count)
{
	// every reduction ends in k_reduce_tail, which mirrors its result into the mapped host copy of scal[]: only a stream
	// synchronisation is needed here (no publish launch, no copy)
	(void)first; (void)count;
	TRY(flush_all(e));
	Strip &s = e->strips[0];
	CU(cudaStreamSynchronize(s.stream));
	return quick_status(e);
}
static double *host_scal(mgb_engine *e) { return e->strips[0].scal_host; }

// one red-black half sweep of colour c (0 = red: (i+j) even), in place; ghost rows valid on return
static int k_rb(mgb_engine *e, int l, int xv, int bv, int colour, double omega, int variant)
{
	TRY(flush_levels(e, l, l));
	const LevelGeom &g = e->geo[l];
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		const int ry = pick_ry(g, S.ni);
		const dim3 gr = stream_grid(g, S.ni, ry);
		if (variant == 0) klaunch(k_rb_half<0>, dim3(gr), dim3(MGB_SB_THREADS), 0, s.stream, S.v[xv], S.v[bv], ldev(e, s, l), colour, omega, ry);
		else              klaunch(k_rb_half<1>, dim3(gr), dim3(MGB_SB_THREADS), 0, s.stream, S.v[xv], S.v[bv], ldev(e, s, l), colour, omega, ry);
		LAUNCHED(e); KCHECK();
	}
	return halo(e, l, xv, 2);
}


// ------------------------------------------------------------------------------------------------ lexicographic sweeps (wavefronts)
// PETSc's natural-order smoothers (-pc_type sor on -map 0,1,2; the default ILU(0)): single GPU, whole level (mgb_wave.cuh)
static int wave(mgb_engine *e, int l, int op, bool reverse, double *x, const double *b, double *t, double omega)
{
	Strip &s = e->strips[0];
	SLevel &S = s.lev[l];
	const int panels = cdiv(e->geo[l].nj, WV_T);
	if (panels > 148) return fail(MGB_EINVAL, "wavefront sweeps need all column panels resident: at most %d columns", 148 * WV_T);
	CU(cudaMemsetAsync(s.wave_progress, 0, sizeof(int) * panels, s.stream));
	WaveArgs a; memset(&a, 0, sizeof a);
	a.op = op; a.reverse = reverse ? 1 : 0; a.L = ldev(e, s, l); a.x = x; a.b = b; a.t = t; a.omega = omega;
	if (S.ilu) { const size_t n = (size_t)S.ni * e->geo[l].pitch; a.invd = S.ilu; a.mS = S.ilu + n; a.mW = S.ilu + 2 * n; }
	a.progress = s.wave_progress; a.spin_limit = e->spin_limit; a.status = status_of(e, s);
	k_wave<<<panels, WV_T, 0, s.stream>>>(a);
	LAUNCHED(e); KCHECK();
	return MGB_OK;
}
static int ilu_factor(mgb_engine *e, int l)
{
	SLevel &S = e->strips[0].lev[l];
	if (S.ilu_valid) return MGB_OK;
	const size_t n = (size_t)S.ni * e->geo[l].pitch;
	if (!S.ilu) { CU(cudaMalloc(&S.ilu, sizeof(double) * 3 * n)); CU(cudaMemsetAsync(S.ilu, 0, sizeof(double) * 3 * n, e->strips[0].stream)); }
	TRY(wave(e, l, WV_ILU_FACTOR, false, nullptr, nullptr, nullptr, 1.0));
	S.ilu_valid = true;
	return MGB_OK;
}

static void swap_vec(mgb_engine *e, int l, int a, int b)
{
	for (auto &s : e->strips) {
		SLevel &S = s.lev[l];
		double *t = S.v[a]; S.v[a] = S.v[b]; S.v[b] = t;
		int p = S.phys[a]; S.phys[a] = S.phys[b]; S.phys[b] = p;
	}
}

// The level smoother: KSPSolve(KSPRICHARDSON, KSP_NORM_NONE, max_it = its) on (b, x) -- exactly `its`
// iterations, no convergence test (ref: src/solver.c:1463-1510; semantics in SURVEY.md appendix A).
// xv / sv: vector ids of the iterate and of the Jacobi ping-pong scratch; on return the iterate is in v[xv]
// (the two device pointers are swapped after every out-of-place sweep) and its ghost rows are valid (depth 2).
static int smooth(mgb_engine *e, int l, const mgb_smoother *sm, int its, bool guess_zero, int bv, int xv, int sv)
{
	const LevelGeom &g = e->geo[l];
	if (sm->type == MGB_SMOOTH_JACOBI) {
		int k = 0;
		if (guess_zero) {
			if (its <= 0) return vec_zero(e, xv, l);
			TRY(flush_levels(e, l, l));
			for (auto &s : e->strips) {
				if (!computes(s, l)) continue;
				SLevel &S = s.lev[l];
				dim3 gr(cdiv(g.pitch, 512), S.ni);
				klaunch(k_jacobi_first, dim3(gr), dim3(256), 0, s.stream, S.v[bv], S.v[xv], ldev(e, s, l), sm->scale);
				LAUNCHED(e); KCHECK();
			}
			TRY(halo(e, l, xv, 2));
			k = 1;
		}
		for (; k < its; ++k) {
			TRY(flush_levels(e, l, l));
			for (auto &s : e->strips) {
				if (!computes(s, l)) continue;
				SLevel &S = s.lev[l];
				const int ry = pick_ry(g, S.ni);
				klaunch(k_stream5<ST_JACOBI>, dim3(stream_grid(g, S.ni, ry)), dim3(MGB_SB_THREADS), 0, s.stream, S.v[xv], S.v[bv], S.v[sv], ldev(e, s, l), sm->scale, nullptr, ry);
				LAUNCHED(e); KCHECK();
			}
			swap_vec(e, l, xv, sv);
			TRY(halo(e, l, xv, 2));
		}
		return MGB_OK;
	}
	if (sm->type == MGB_SMOOTH_RBSOR) {
		// KSPSolve_Richardson hands the whole loop to PCApplyRichardson_SOR only when scale == 1:
		// MatSOR(its * pc_its * lits sweeps).  Other scales go through PCApply_SOR per iteration: not offered.
		if (sm->scale != 1.0) return fail(MGB_EINVAL, "red-black SOR needs -ksp_richardson_scale 1 (PCApplyRichardson_SOR path)");
		if (guess_zero) TRY(vec_zero(e, xv, l));
		const int total = its * (sm->sor_its > 0 ? sm->sor_its : 1);
		const double om = sm->omega;
		for (int k = 0; k < total; ++k) {
			if (sm->sor_sweep == MGB_SOR_SYMMETRIC) {
				// forward: red, black ; backward: black (no-op when omega == 1: x = t * idiag again), red.
				// A red half sweep directly after a red half sweep recomputes the same values when omega == 1.
				const bool prev_red = (k > 0);
				if (!(prev_red && om == 1.0)) TRY(k_rb(e, l, xv, bv, 0, om, 0));
				TRY(k_rb(e, l, xv, bv, 1, om, 0));
				if (om != 1.0) TRY(k_rb(e, l, xv, bv, 1, om, 0));
				TRY(k_rb(e, l, xv, bv, 0, om, 0));
			} else if (sm->sor_sweep == MGB_SOR_FORWARD) {
				TRY(k_rb(e, l, xv, bv, 0, om, 0));
				TRY(k_rb(e, l, xv, bv, 1, om, 0));
			} else if (sm->sor_sweep == MGB_SOR_BACKWARD) {
				const int variant = (guess_zero && k == 0) ? 0 : 1;
				TRY(k_rb(e, l, xv, bv, 1, om, variant));
				TRY(k_rb(e, l, xv, bv, 0, om, variant));
			} else return fail(MGB_EINVAL, "unknown sor_sweep %d", sm->sor_sweep);
		}
		return MGB_OK;
	}
	if (sm->type == MGB_SMOOTH_LEXSOR) {
		// PCApplyRichardson_SOR (scale == 1): MatSOR with its * pc_its * lits sweeps in the natural numbering; a zero initial
		// guess is the general sweep on x = 0 (the terms MatSOR skips are products with zero)
		if (sm->scale != 1.0) return fail(MGB_EINVAL, "lexicographic SOR needs -ksp_richardson_scale 1 (PCApplyRichardson_SOR path)");
		SLevel &S = e->strips[0].lev[l];
		if (guess_zero) TRY(vec_zero(e, xv, l));
		const int total = its * (sm->sor_its > 0 ? sm->sor_its : 1);
		for (int k = 0; k < total; ++k) {
			if (sm->sor_sweep == MGB_SOR_SYMMETRIC) {
				TRY(wave(e, l, WV_SOR_FWD, false, S.v[xv], S.v[bv], S.v[sv], sm->omega));      // t = b - L x in the scratch vector
				TRY(wave(e, l, WV_SOR_BWD_T, true, S.v[xv], S.v[bv], S.v[sv], sm->omega));
			} else if (sm->sor_sweep == MGB_SOR_FORWARD) {
				TRY(wave(e, l, WV_SOR_FWD, false, S.v[xv], S.v[bv], nullptr, sm->omega));
			} else if (sm->sor_sweep == MGB_SOR_BACKWARD) {
				TRY(wave(e, l, WV_SOR_BWD_B, true, S.v[xv], S.v[bv], nullptr, sm->omega));
			} else return fail(MGB_EINVAL, "unknown sor_sweep %d", sm->sor_sweep);
		}
		return MGB_OK;
	}
	if (sm->type == MGB_SMOOTH_ILU0) {
		// KSPSolve_Richardson, general path (PCILU has no PCApplyRichardson), KSP_NORM_NONE:
		// r = b - A x (r = b from a zero guess); repeat: z = (LU)^-1 r ; x += scale z ; r = b - A x (skipped after the last)
		SLevel &S = e->strips[0].lev[l];
		// work vector for r: the level's residual vector, or -- when the PCMG preconditioner smooths on (r, z) of the outer
		// Krylov method on the finest level -- the Krylov vector Q (= A p, dead between the CG update and the next A p)
		const int rv = (bv == MGB_VEC_R || xv == MGB_VEC_R || sv == MGB_VEC_R) ? MGB_VEC_Q : MGB_VEC_R;
		if (rv == MGB_VEC_Q && (l != 0 || bv == rv || xv == rv || sv == rv)) return fail(MGB_EINVAL, "ILU(0) smoothing: no free work vector on level %d", l);
		TRY(ilu_factor(e, l));
		if (guess_zero) { TRY(vec_zero(e, xv, l)); TRY(k_vecop<2>(e, l, rv, bv, 0.0)); }
		else TRY(k_residual(e, l, xv, bv, rv));
		for (int k = 0; k < its; ++k) {
			TRY(k_vecop<2>(e, l, sv, rv, 0.0));                                        // z <- r, solved in place
			TRY(wave(e, l, WV_ILU_FWD, false, S.v[sv], nullptr, nullptr, 1.0));
			TRY(wave(e, l, WV_ILU_BWD, true, S.v[sv], nullptr, nullptr, 1.0));
			TRY(k_vecop<0>(e, l, xv, sv, sm->scale));                                  // x <- x + scale z
			if (k + 1 < its) TRY(k_residual(e, l, xv, bv, rv));
		}
		return MGB_OK;
	}
	return fail(MGB_EINVAL, "unknown smoother type %d", sm->type);
}

static int check_smoother(mgb_engine *e, const mgb_smoother *s)
{
	if (!s) return fail(MGB_EINVAL, "null smoother");
	if (s->type == MGB_SMOOTH_RBSOR) {
		if (s->omega <= 0.0 || s->omega >= 2.0) return fail(MGB_EINVAL, "SOR omega must be in (0,2)");
		TRY(set_sor_omega(e, s->omega));
	} else if (s->type == MGB_SMOOTH_LEXSOR || s->type == MGB_SMOOTH_ILU0) {
		if (e->cfg.red_black_numbering) return fail(MGB_EINVAL, "lexicographic SOR / ILU(0) belong to the natural numbering (-map 0,1,2)");
		if (e->P > 1) return fail(MGB_EINVAL, "lexicographic SOR / ILU(0) are sequential across grid rows: they run on one GPU (PETSc's parallel "
		                          "versions are processor-local and give other numbers); use Jacobi or -map 3 red-black SOR on strips");
		if (s->type == MGB_SMOOTH_LEXSOR) {
			if (s->omega <= 0.0 || s->omega >= 2.0) return fail(MGB_EINVAL, "SOR omega must be in (0,2)");
			TRY(set_sor_omega(e, s->omega));
		} else {
			// factor every level now: allocations are not allowed later, inside a graph capture
			for (int l = 0; l < e->L; ++l) if (e->geo[l].coef_set) TRY(ilu_factor(e, l));
		}
	} else if (s->type != MGB_SMOOTH_JACOBI) return fail(MGB_EINVAL, "unknown smoother type %d", s->type);
	return MGB_OK;
}

// the coarse level as a strip sees it in a transfer from / to the distributed level above: its own strip of a
// distributed coarse level, or its rows [rows[r], rows[r+1]) of the whole first agglomerated level
static LevelDev coarse_view(const mgb_engine *e, const Strip &s, int lc, bool fine_dist, size_t *row_off)
{
	const LevelGeom &gc = e->geo[lc];
	*row_off = 0;
	if (gc.dist || e->P == 1 || !fine_dist) return ldev(e, s, lc);
	*row_off = (size_t)gc.rows[s.rank] * gc.pitch;
	return ldev_rows(e, s, lc, gc.rows[s.rank], gc.rows[s.rank + 1]);
}

// b[l+1] = res[l] * (fused ? bv[l] - A[l] xv[l] : rv[l])   (ref: src/solver.c:1534-1535).  Every strip of a
// distributed level produces its own coarse rows; when the coarse level is the first agglomerated one they are
// gathered on rank 0, otherwise the coarse ghost rows are exchanged.  Needs ghost depth 2 of xv and 1 of bv / rv.
static int restrict_streamed(mgb_engine *e, int l, int bv, int xv);
static int restrict_to_coarse(mgb_engine *e, int l, int bv, int xv, int rv, bool fused)
{
	TRY(flush_levels(e, l, l));
	const LevelGeom &gf = e->geo[l], &gc = e->geo[l + 1];
	if (fused && !e->cfg.red_black_numbering) {
		TRY(restrict_streamed(e, l, bv, xv));
		if (gc.dist) return halo(e, l + 1, MGB_VEC_B, 2);
		if (gf.dist) return gather_rows(e, l + 1, MGB_VEC_B);
		return MGB_OK;
	}
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &F = s.lev[l], &C = s.lev[l + 1];
		size_t coff; const LevelDev cd = coarse_view(e, s, l + 1, gf.dist, &coff);
		dim3 blk(32, 4), gr(cdiv(gc.pitch, 32), cdiv(cd.ni, 4));
		double *bc = C.v[MGB_VEC_B] + coff;
		if (fused) klaunch(k_restrict<1>, dim3(gr), dim3(blk), 0, s.stream, F.v[xv], F.v[bv], nullptr, bc, ldev(e, s, l), cd, e->R3);
		else       klaunch(k_restrict<0>, dim3(gr), dim3(blk), 0, s.stream, nullptr, nullptr, F.v[rv], bc, ldev(e, s, l), cd, e->R3);
		LAUNCHED(e); KCHECK();
	}
	if (gc.dist) return halo(e, l + 1, MGB_VEC_B, 2);
	if (gf.dist) return gather_rows(e, l + 1, MGB_VEC_B);
	return MGB_OK;
}
// x[l] += pro[l] * u[l+1]   (ref: src/solver.c:1540-1541 ; PCMG: MatInterpolateAdd); ghost rows of x valid on return
static int prolong_add(mgb_engine *e, int l, int xv, bool multadd)
{
	const LevelGeom &gf = e->geo[l], &gc = e->geo[l + 1];
	if (gf.dist && !gc.dist) TRY(bcast_rows(e, l + 1, MGB_VEC_U));
	TRY(flush_levels(e, l, l + 1));
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &F = s.lev[l], &C = s.lev[l + 1];
		size_t coff; const LevelDev cd = coarse_view(e, s, l + 1, gf.dist, &coff);
		dim3 blk(32, 4), gr(cdiv(gf.pitch, 64), cdiv(F.ni, 4));
		const double *uc = C.v[MGB_VEC_U] + coff;
		if (multadd) klaunch(k_prolong_add<1>, dim3(gr), dim3(blk), 0, s.stream, F.v[xv], uc, ldev(e, s, l), cd, e->P3);
		else         klaunch(k_prolong_add<0>, dim3(gr), dim3(blk), 0, s.stream, F.v[xv], uc, ldev(e, s, l), cd, e->P3);
		LAUNCHED(e); KCHECK();
	}
	return halo(e, l, xv, 2);
}


// ------------------------------------------------------------------------------------------------ fused legs (Jacobi)
#define HALO_DEPTH MGB_GHOST_ROWS

// the fused kernels' power-of-two path (OP = 2) also fuses the transfer weights into fma: only when those are exact too
static LevelDev fused_ldev(const mgb_engine *e, const Strip &s, int l)
{
	LevelDev d = ldev(e, s, l);
	if (d.uniform == 2 && e->L > 1 && !e->transfer_pow2) d.uniform = 1;
	return d;
}

template <int D, int PRE, int POST, int SMK = 0>
static void launch_jfused(const FusedArgs &a, dim3 grid, cudaStream_t st)
{
	// > 48 KB of dynamic shared memory needs the opt-in once per kernel AND per device (a function attribute lives in
	// the device's context): one flag per device, set under a lock
	static std::mutex mu;
	static bool optin[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	{
		std::lock_guard<std::mutex> lk(mu);
		if (dev >= 0 && dev < 64 && !optin[dev]) {
			if (cudaFuncSetAttribute(k_jfused<D, PRE, POST, SMK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jf_smem_bytes<D>()) == cudaSuccess)
				optin[dev] = true;                 // on failure the launch below reports the error through KCHECK
		}
	}
	klaunch(k_jfused<D, PRE, POST, SMK>, dim3(grid), dim3(FJ_THREADS), jf_smem_bytes<D>(), st, a);
}

template <int D, int SMK = 0>
static int dispatch_jfused(int pre, int post, const FusedArgs &a, dim3 grid, cudaStream_t st)
{
#define JF(PRE_, POST_) if (pre == PRE_ && post == POST_) { launch_jfused<D, PRE_, POST_, SMK>(a, grid, st); return MGB_OK; }
	JF(PRE_GIVEN, POST_NONE) JF(PRE_GIVEN, POST_RESTRICT) JF(PRE_GIVEN, POST_NORM)
	JF(PRE_ZERO, POST_NONE) JF(PRE_ZERO, POST_RESTRICT) JF(PRE_ZERO, POST_NORM)
	JF(PRE_PROLONG, POST_NONE) JF(PRE_PROLONG, POST_NORM)
	JF(PRE_PROLONG_MULTADD, POST_NONE)
	if (SMK == 0) { JF(PRE_PROLONG_MULTADD, POST_DOT) }
#undef JF
	return fail(MGB_EINVAL, "fused kernel: combination pre %d / post %d not built", pre, post);
}

// rows per block of the fused kernel.  The kernel runs 4 blocks per SM (registers and shared memory): 592 resident
// blocks on 148 SMs.  Large levels get exactly one wave (no partially filled second wave), i.e. as many row chunks as
// fit next to the column tiles; small levels get chunks of at least 4 rows (the row loop of a block is a serial chain: short chunks, many blocks).
static int pick_rows(const LevelGeom &g, int ni)
{
	const int tiles = cdiv(g.pitch, FJ_VALID);
	int chunks = (148 * (512 / FJ_THREADS)) / tiles;      // resident blocks: 512 threads per SM
	if (chunks < 1) chunks = 1;
	int r = cdiv(ni, chunks);
	if (r < 4) r = 4;
	return (r + 1) & ~1;
}

// b[l+1] = res[l] * (b - A x) as a streaming pass (the fused kernel with zero sweeps): natural numbering only
static int restrict_streamed(mgb_engine *e, int l, int bv, int xv)
{
	const LevelGeom &g = e->geo[l];
	TRY(flush_levels(e, l, l));
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		SLevel &S = s.lev[l];
		FusedArgs a; memset(&a, 0, sizeof a);
		a.u_in = S.v[xv]; a.b = S.v[bv]; a.u_out = nullptr;
		a.F = fused_ldev(e, s, l); a.scale = 1.0; a.gni = g.gni;
		a.rows = pick_rows(g, S.ni);
		a.R3 = e->R3; a.P3 = e->P3;
		size_t coff; a.C = coarse_view(e, s, l + 1, g.dist, &coff);
		a.bc = s.lev[l + 1].v[MGB_VEC_B] + coff;
		int tiles = cdiv(g.pitch, FJ_VALID);
		const int need = cdiv(2 * e->geo[l + 1].pitch, FJ_VALID);
		if (need > tiles) tiles = need;
		dim3 grid(tiles, cdiv(S.ni, a.rows));
		launch_jfused<0, PRE_GIVEN, POST_RESTRICT>(a, grid, s.stream);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}

// Jacobi on the natural numbering, or red-black SOR (-map 3) in the sweep orders that use k_rb_half's variant 0
static bool fusable(const mgb_engine *e, const mgb_smoother *sm)
{
	if (sm->type == MGB_SMOOTH_JACOBI) return !e->cfg.red_black_numbering;
	if (sm->type == MGB_SMOOTH_RBSOR)
		return e->cfg.red_black_numbering && sm->scale == 1.0 && (sm->sor_sweep == MGB_SOR_SYMMETRIC || sm->sor_sweep == MGB_SOR_FORWARD);
	return false;
}
// Does the fused leg pay on level l?  Jacobi: always (3 sweeps + residual + transfer in one pass, and the persistent bottom
// kernel below 64 rows).  Red-black SOR: a V(3,3) leg is 7 half sweeps = 3 passes of the fused kernel; that beats 7 + 1
// one-sweep launches only where the level is bandwidth-bound (at 1023 rows and below the data sits in L2 and a one-sweep
// launch costs ~3 us against 8-20 us of pipeline latency per fused pass).  On strips the distributed levels are fused
// (their ghost rows travel inside the kernels), the agglomerated ones follow the same size rule on rank 0.
static bool fuse_level(const mgb_engine *e, const mgb_smoother *sm, int l)
{
	if (sm->type != MGB_SMOOTH_RBSOR) return true;
	if (e->P > 1) return e->geo[l].dist;
	return e->geo[l].gni >= e->rb_fuse_min_rows;
}
// the stages of one smoothing call: `its` Jacobi sweeps (colour -1) or the half sweeps of MatSOR on the red-first numbering
// in the order smooth() launches them (0 red, 1 black)
static std::vector<int> smoother_stages(const mgb_smoother *sm, int its)
{
	std::vector<int> st;
	if (sm->type == MGB_SMOOTH_JACOBI) { st.assign(its, -1); return st; }
	const int total = its * (sm->sor_its > 0 ? sm->sor_its : 1);
	for (int k = 0; k < total; ++k) {
		if (sm->sor_sweep == MGB_SOR_SYMMETRIC) {
			if (!(k > 0 && sm->omega == 1.0)) st.push_back(0);     // a red half sweep right after a red one recomputes the same values
			st.push_back(1);
			if (sm->omega != 1.0) st.push_back(1);
			st.push_back(0);
		} else { st.push_back(0); st.push_back(1); }
	}
	return st;
}

// One leg on level l: [x += pro * u[l+1]] -> `its` Jacobi sweeps on (b = bv, x = xv) -> [b[l+1] = res * (b - A x)] or
// [scal[norm_slot] = ||b - A x||].  Sweeps beyond FJ_MAXD are chained as extra passes.  On return the iterate is in v[xv]
// with valid ghost rows; a restricted right-hand side has been exchanged / gathered.
// ---- strip-to-strip traffic of a fused leg, done inside k_jfused (FusedComm, mgb_fused.cuh)
static unsigned long long *flag_at(mgb_engine *e, int dst_rank, int chan, int src_rank) { return flags_of(e, dst_rank) + (size_t)chan * MGB_MAX_RANKS + src_rank; }
static void add_wait(FusedComm &X, int which, int &n, mgb_engine *e, const Strip &s, int chan, int src_rank)
{
	if (n >= FJ_MAXWAIT) return;
	X.w_flag[which][n] = flag_at(e, s.rank, chan, src_rank);
	X.w_ver[which][n] = ver_of(e, s) + chan;
	++n;
}
// Fill a.X for strip s.  out_phys: physical buffer of u_out; bcast_out: level l is the first agglomerated one and its
// result feeds the prolongation of every rank.  Returns through a.bc a pointer into rank 0's HBM when the restricted
// right-hand side is gathered there.
static void fused_comm(mgb_engine *e, Strip &s, int l, int D, int pre_k, int post_k, int bv, int xv, int out_phys, bool bcast_out,
                       dim3 grid, FusedArgs &a)
{
	FusedComm &X = a.X;
	const int r = s.rank, P = e->P;
	const LevelGeom &g = e->geo[l];
	X.status = status_of(e, s); X.status_host = s.status_host_dev; X.spin_limit = e->spin_limit;
	const bool prolong = pre_k == PRE_PROLONG || pre_k == PRE_PROLONG_MULTADD;
	int first_chan = -1;
	auto channel = [&](int chan) -> int { const int c = X.nch++; X.ver[c] = ver_of(e, s) + chan; X.nsig[c] = 0; if (first_chan < 0) first_chan = chan; return c; };
	if (g.dist) {
		const size_t pitch = g.pitch;
		const int ni = e->lay[r].ni[l];
		// ghost rows this leg reads
		if (pre_k != PRE_ZERO) {
			const int ch = CH_HALO(l, s.lev[l].phys[xv]);
			if (r > 0) add_wait(X, 0, X.nw_top, e, s, ch, r - 1);
			if (r < P - 1) add_wait(X, 1, X.nw_bot, e, s, ch, r + 1);
		}
		{
			const int ch = CH_HALO(l, s.lev[l].phys[bv]);
			if (r > 0) add_wait(X, 0, X.nw_top, e, s, ch, r - 1);
			if (r < P - 1) add_wait(X, 1, X.nw_bot, e, s, ch, r + 1);
		}
		if (prolong) {
			if (e->geo[l + 1].dist) {
				const int ch = CH_HALO(l + 1, s.lev[l + 1].phys[MGB_VEC_U]);
				if (r > 0) add_wait(X, 0, X.nw_top, e, s, ch, r - 1);
				if (r < P - 1) add_wait(X, 1, X.nw_bot, e, s, ch, r + 1);
			} else if (r != 0) add_wait(X, 2, X.nw_all, e, s, CH_BCAST(l + 1), 0);
		}
		// ghost rows of u_out pushed to the neighbours
		if (D > 0) {
			const int ch = CH_HALO(l, out_phys), c = channel(ch);
			X.pu[X.npu++] = {r > 0 ? peer_vec(e, r - 1, l, out_phys) + (size_t)e->lay[r - 1].ni[l] * pitch : nullptr, 0, HALO_DEPTH};
			X.pu[X.npu++] = {r < P - 1 ? peer_vec(e, r + 1, l, out_phys) - (size_t)ni * pitch : nullptr, ni - HALO_DEPTH, ni};
			if (r > 0) X.sig[c][X.nsig[c]++] = flag_at(e, r - 1, ch, r);
			if (r < P - 1) X.sig[c][X.nsig[c]++] = flag_at(e, r + 1, ch, r);
		}
		if (post_k == POST_RESTRICT) {
			const LevelGeom &gc = e->geo[l + 1];
			const int kB = s.lev[l + 1].phys[MGB_VEC_B];
			if (gc.dist) {
				const int ch = CH_HALO(l + 1, kB), c = channel(ch);
				const int nic = e->lay[r].ni[l + 1];
				X.pb[X.npb++] = {r > 0 ? peer_vec(e, r - 1, l + 1, kB) + (size_t)e->lay[r - 1].ni[l + 1] * gc.pitch : nullptr, 0, HALO_DEPTH};
				X.pb[X.npb++] = {r < P - 1 ? peer_vec(e, r + 1, l + 1, kB) - (size_t)nic * gc.pitch : nullptr, nic - HALO_DEPTH, nic};
				if (r > 0) X.sig[c][X.nsig[c]++] = flag_at(e, r - 1, ch, r);
				if (r < P - 1) X.sig[c][X.nsig[c]++] = flag_at(e, r + 1, ch, r);
			} else {
				// gather: this rank's coarse rows go straight into rank 0's copy of the first agglomerated level
				const int ch = CH_GATHER(l + 1), c = channel(ch);
				a.bc = peer_vec(e, 0, l + 1, kB) + (size_t)gc.rows[r] * gc.pitch;
				X.bc_remote = 1;
				if (r != 0) X.sig[c][X.nsig[c]++] = flag_at(e, 0, ch, r);
			}
		}
	} else if (P > 1 && l == e->La && r == 0) {
		// the first agglomerated level on rank 0: its right-hand side was gathered, its result may be broadcast
		if (pre_k == PRE_ZERO) for (int t = 1; t < P; ++t) add_wait(X, 2, X.nw_all, e, s, CH_GATHER(l), t);
		if (bcast_out && D > 0) {
			const int ch = CH_BCAST(l), c = channel(ch);
			for (int t = 1; t < P && X.npu < FJ_MAXPUSH; ++t) {
				int c0 = g.rows[t] - 3, c1 = g.rows[t + 1] + 3;               // the fused up leg reads 3 coarse ghost rows
				if (c0 < 0) c0 = 0;
				if (c1 > g.gni) c1 = g.gni;
				X.pu[X.npu++] = {peer_vec(e, t, l, out_phys), c0, c1};
				X.sig[c][X.nsig[c]++] = flag_at(e, t, ch, 0);
			}
		}
	}
	// blocks that take a ticket: the row chunks that intersect a push range (all of them when bc is remote), times the column tiles
	if (X.nch) {
		int chunks = 0;
		const int ni = s.lev[l].ni;
		for (int y0 = 0; y0 < ni; y0 += a.rows) {
			const int y1 = y0 + a.rows < ni ? y0 + a.rows : ni;
			bool takes = X.bc_remote && post_k == POST_RESTRICT;
			for (int k = 0; k < X.npu; ++k) takes |= (D > 0 && y0 < X.pu[k].hi && y1 > X.pu[k].lo);
			for (int k = 0; k < X.npb; ++k) takes |= (post_k == POST_RESTRICT && (y0 >> 1) < X.pb[k].hi && (y1 >> 1) > X.pb[k].lo);
			chunks += takes ? 1 : 0;
		}
		X.npushblocks = chunks * (int)grid.x;
		X.ticket = ticket_of(e, s) + first_chan;
		if (X.npushblocks == 0) X.nch = 0;
	}
}
// rank 0 consumes a gathered right-hand side in a kernel that cannot wait by itself (the persistent bottom kernel)
static int wait_gather(mgb_engine *e, int l)
{
	if (e->P == 1) return MGB_OK;
	for (auto &s : e->strips) {
		if (s.rank != 0) continue;
		XferArgs a; memset(&a, 0, sizeof a);
		const int chan = CH_GATHER(l);
		a.ver = ver_of(e, s) + chan; a.ticket = ticket_of(e, s) + chan; a.status = status_of(e, s); a.status_host = s.status_host_dev;
		a.spin_limit = e->spin_limit; a.do_push = 0; a.do_wait = 1;
		for (int t = 1; t < e->P; ++t) a.wait_flag[a.nwait++] = flag_at(e, 0, chan, t);
		k_xfer<<<1, 32, 0, s.stream>>>(a);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}

// the ghost rows of v[which] on level l were pushed by the neighbours from inside a fused leg: a consumer that cannot wait by
// itself gets a wait-only launch on the channel (passes at once when the rows came by an ordinary exchange)
static int wait_halo(mgb_engine *e, int l, int which)
{
	if (e->P == 1 || !e->geo[l].dist) return MGB_OK;
	TRY(flush_levels(e, l, l));
	for (auto &s : e->strips) {
		if (!computes(s, l)) continue;
		XferArgs a; memset(&a, 0, sizeof a);
		const int chan = CH_HALO(l, s.lev[l].phys[which]);
		a.ver = ver_of(e, s) + chan; a.ticket = ticket_of(e, s) + chan; a.status = status_of(e, s); a.status_host = s.status_host_dev;
		a.spin_limit = e->spin_limit; a.do_push = 0; a.do_wait = 1;
		if (s.rank > 0) a.wait_flag[a.nwait++] = flag_at(e, s.rank, chan, s.rank - 1);
		if (s.rank < e->P - 1) a.wait_flag[a.nwait++] = flag_at(e, s.rank, chan, s.rank + 1);
		k_xfer<<<1, 32, 0, s.stream>>>(a);
		LAUNCHED(e); KCHECK();
	}
	return MGB_OK;
}

// before a prolongation from the first agglomerated level lc onto distributed strips: its rows must reach every rank
static int need_bcast(mgb_engine *e, int lc)
{
	if (e->bcast_done == lc) { e->bcast_done = -1; return MGB_OK; }     // they travelled inside the fused leg that produced them
	return bcast_rows(e, lc, MGB_VEC_U);
}

static int fused_leg(mgb_engine *e, int l, const mgb_smoother *sm, int its, int pre, int post, int bv, int xv, int sv, int norm_slot,
                     bool bcast_out = false)
{
	const LevelGeom &g = e->geo[l];
	if (its < 1) return fail(MGB_EINVAL, "fused leg needs at least one sweep");
	const std::vector<int> stages = smoother_stages(sm, its);
	const bool rb = sm->type == MGB_SMOOTH_RBSOR;
	its = (int)stages.size();
	int done = 0;
	while (done < its) {
		const int D = (its - done > FJ_MAXD) ? FJ_MAXD : its - done;
		const bool firstc = done == 0, lastc = done + D == its;
		const int pre_k = firstc ? pre : PRE_GIVEN, post_k = lastc ? post : POST_NONE;
		std::vector<int> nb;
		TRY(flush_levels(e, l, (pre_k == PRE_PROLONG || pre_k == PRE_PROLONG_MULTADD) ? l + 1 : l));
		const bool inkernel = e->P > 1 && e->inkernel;     // ghost rows / gather / broadcast travel inside k_jfused
		const bool bcast_k = bcast_out && lastc && inkernel && !g.dist && l == e->La;
		if (bcast_k) e->bcast_done = l;
		for (auto &s : e->strips) {
			if (!computes(s, l)) {
				// the ranks that do not compute on a broadcast level keep their version counter of the channel in step
				if (bcast_k) { klaunch(k_bump, dim3(1), dim3(1), 0, s.stream, ver_of(e, s) + CH_BCAST(l)); LAUNCHED(e); KCHECK(); }
				continue;
			}
			SLevel &S = s.lev[l];
			FusedArgs a; memset(&a, 0, sizeof a);
			a.u_in = S.v[xv]; a.b = S.v[bv]; a.u_out = S.v[sv];
			a.F = fused_ldev(e, s, l); a.scale = rb ? sm->omega : sm->scale; a.gni = g.gni;
			for (int k = 0; k < D; ++k) if (stages[done + k] == 1) a.rbmask |= 1 << k;
			a.rows = pick_rows(g, S.ni);
			a.R3 = e->R3; a.P3 = e->P3;
			int tiles = cdiv(g.pitch, FJ_VALID);
			if (pre_k == PRE_PROLONG || pre_k == PRE_PROLONG_MULTADD) {
				size_t coff; a.C = coarse_view(e, s, l + 1, g.dist, &coff);
				a.uc = s.lev[l + 1].v[MGB_VEC_U] + coff;
			}
			if (post_k == POST_RESTRICT) {
				size_t coff; a.C = coarse_view(e, s, l + 1, g.dist, &coff);
				a.bc = s.lev[l + 1].v[MGB_VEC_B] + coff;
				const int need = cdiv(2 * e->geo[l + 1].pitch, FJ_VALID);
				if (need > tiles) tiles = need;
			}
			dim3 grid(tiles, cdiv(S.ni, a.rows));
			if (post_k == POST_NORM || post_k == POST_DOT) {
				if ((size_t)grid.x * grid.y > s.partial_cap) return fail(MGB_EINVAL, "grid too large for the partial-sum buffer");
				a.partial = s.partial;
				nb.push_back((int)(grid.x * grid.y));
			}
			if (inkernel) fused_comm(e, s, l, D, pre_k, post_k, bv, xv, S.phys[sv], bcast_k, grid, a);
			int rc;
			switch (D + (rb ? 10 : 0)) {
			case 1: rc = dispatch_jfused<1>(pre_k, post_k, a, grid, s.stream); break;
			case 2: rc = dispatch_jfused<2>(pre_k, post_k, a, grid, s.stream); break;
			case 3: rc = dispatch_jfused<3>(pre_k, post_k, a, grid, s.stream); break;
			case 11: rc = dispatch_jfused<1, 1>(pre_k, post_k, a, grid, s.stream); break;
			case 12: rc = dispatch_jfused<2, 1>(pre_k, post_k, a, grid, s.stream); break;
			default: rc = dispatch_jfused<3, 1>(pre_k, post_k, a, grid, s.stream); break;
			}
			TRY(rc);
			LAUNCHED(e); KCHECK();
		}
		swap_vec(e, l, xv, sv);
		if (!inkernel) {
			TRY(halo(e, l, xv, HALO_DEPTH));
			if (post_k == POST_RESTRICT) {
				if (e->geo[l + 1].dist) TRY(halo(e, l + 1, MGB_VEC_B, HALO_DEPTH));
				else if (g.dist) TRY(gather_rows(e, l + 1, MGB_VEC_B));
			}
		}
		if (post_k == POST_NORM) TRY(reduce_tail(e, l, nb, norm_slot, 1));
		if (post_k == POST_DOT) TRY(reduce_tail(e, l, nb, norm_slot, 0, TAIL_BETA));     // z'r of CG, with beta / ratio derived on the device
		done += D;
	}
	return MGB_OK;
}


// ------------------------------------------------------------------------------------------------ persistent bottom of the cycle
// first level handled by k_coarse_cycle: the finest level l >= 1 that is whole on rank 0 and has at most
// coarse_threshold rows (L if there is none or the bottom has too many levels for one launch)
// (Measured for red-black SOR at 1025^2 / 7 levels, MGB_RB_BOTTOM_ROWS: cluster kernel from 63 rows 4.0-4.1 k cycles/s, from 127
// rows 3.6 k, from 255 rows 2.9 k -- 16 cluster phases per level cost more than 17 one-sweep launches of ~3 us; 63 for both smoothers.)
static int bottom_start(const mgb_engine *e, const mgb_smoother *sm)
{
	if (e->coarse_threshold <= 0) return e->L;
	const int thr = (sm && sm->type == MGB_SMOOTH_RBSOR) ? e->rb_coarse_threshold : e->coarse_threshold;
	for (int l = 1; l < e->L; ++l)
		if (!e->geo[l].dist && e->geo[l].gni <= thr && e->geo[l].nj <= 2 * thr)
			return (e->L - l <= CC_MAXLEV) ? l : e->L;
	return e->L;
}
// levels lp .. L-1 in one launch on rank 0: zero-guess smoothing and restriction down, the coarsest smoothing,
// correction and smoothing up.  b[lp] must be complete on rank 0; on return u[lp] holds the correction.
static int bottom_cycle(mgb_engine *e, int lp, const mgb_smoother *sm_level, int its_level, const mgb_smoother *sm_coarse, int its_coarse, bool multadd)
{
	TRY(flush_levels(e, lp, MGB_MAXL));
	const bool rb = sm_level->type == MGB_SMOOTH_RBSOR;
	// red-black SOR: the half sweeps of one smoothing call (colours in launch order), at most 32 per call
	const std::vector<int> st_level = smoother_stages(sm_level, its_level), st_coarse = smoother_stages(sm_coarse, its_coarse);
	if (rb && (st_level.size() > 32 || st_coarse.size() > 32)) return fail(MGB_EINVAL, "too many half sweeps per smoothing call for the bottom kernel");
	auto mask_of = [](const std::vector<int> &st) { unsigned m = 0u; for (size_t k = 0; k < st.size(); ++k) if (st[k] == 1) m |= 1u << k; return m; };
	for (auto &s : e->strips) {
		if (s.rank != 0) continue;
		CoarseArgs a; memset(&a, 0, sizeof a);
		a.nlev = e->L - lp; a.multadd = multadd ? 1 : 0; a.R3 = e->R3; a.P3 = e->P3;
		a.rb = rb ? 1 : 0; a.omega = sm_level->omega;
		for (int l = lp; l < e->L; ++l) {
			SLevel &S = s.lev[l]; CLevel &c = a.lev[l - lp];
			const bool last = l == e->L - 1;
			c.x = S.v[MGB_VEC_U]; c.w = S.v[MGB_VEC_W]; c.b = S.v[MGB_VEC_B]; c.coef = S.coef;
			c.ni = S.ni; c.nj = e->geo[l].nj; c.pitch = e->geo[l].pitch; c.uniform = e->geo[l].uniform;
			if (rb) {
				c.its_down = (int)(last ? st_coarse.size() : st_level.size()); c.its_up = last ? 0 : (int)st_level.size();
				c.mask_down = last ? mask_of(st_coarse) : mask_of(st_level); c.mask_up = last ? 0u : mask_of(st_level);
				c.scale = 1.0;
			} else {
				c.its_down = last ? its_coarse : its_level; c.its_up = last ? 0 : its_level;
				c.scale = last ? sm_coarse->scale : sm_level->scale;
			}
		}
		{
			// 152 KB of dynamic shared memory: opt-in once per device (as in launch_jfused)
			static std::mutex mu; static bool optin[64] = {};
			int dev = 0; cudaGetDevice(&dev);
			std::lock_guard<std::mutex> lk(mu);
			if (dev >= 0 && dev < 64 && !optin[dev] &&
			    cudaFuncSetAttribute(k_coarse_cycle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CC_SMEM_BYTES) == cudaSuccess) optin[dev] = true;
		}
		klaunch(k_coarse_cycle, dim3(CC_CTAS), dim3(CC_THREADS), CC_SMEM_BYTES, s.stream, a);
		LAUNCHED(e); KCHECK();
	}
	if (!rb)                                                  // red-black half sweeps are in place: no ping-pong
		for (int l = lp; l < e->L; ++l) {
			const bool last = l == e->L - 1;
			const int swaps = last ? its_coarse - 1 : (its_level - 1) + its_level;
			if (swaps & 1) swap_vec(e, l, MGB_VEC_U, MGB_VEC_W);
		}
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ single ops (C-ABI)
// The single operations are collective in a multi-rank run.  They refresh the ghost rows of their inputs first
// (vectors may have been set from the host); the solvers keep ghosts valid incrementally.
#define NEED(e) do { if (!(e)) return fail(MGB_EINVAL, "null engine"); } while (0)
#define GHOSTS(e, l, which) TRY(halo(e, l, which, 2))

extern "C" int mgb_error_norms_separable(mgb_engine *e, const double *sx, const double *sy, double error[3])
{
	if (!e || !sx || !sy || !error) return fail(MGB_EINVAL, "null argument");
	if (e->P > 1 && e->strips.size() == 1)
		return fail(MGB_EINVAL, "error norms of a multi-process run are combined by the caller from the per-rank rows");
	const LevelGeom &g = e->geo[0];
	const bool multi = e->strips.size() > 1;
	double mx = 0.0, s1 = 0.0, s2 = 0.0;
	for (auto &s : e->strips) {
		SLevel &S = s.lev[0];
		CU(cudaMemcpyAsync(s.tab_x, sx, sizeof(double) * g.nj, cudaMemcpyHostToDevice, s.stream));
		CU(cudaMemcpyAsync(s.tab_y, sy, sizeof(double) * g.gni, cudaMemcpyHostToDevice, s.stream));
		const size_t total = (size_t)S.ni * g.pitch;
		const int blocks = (int)((total + MGB_RED_THREADS - 1) / MGB_RED_THREADS < 1184 ? (total + MGB_RED_THREADS - 1) / MGB_RED_THREADS : 1184);
		k_error<<<blocks, MGB_RED_THREADS, 0, s.stream>>>(S.v[MGB_VEC_U], s.tab_x, s.tab_y, ldev(e, s, 0), s.partial);
		LAUNCHED(e); KCHECK();
		k_error2<<<1, 32, 0, s.stream>>>(s.partial, blocks, s.scal + 8, multi ? 0 : 1);
		LAUNCHED(e); KCHECK();
		k_publish<<<1, 32, 0, s.stream>>>(s.scal_host_dev, s.scal, 8, 3);
		LAUNCHED(e); KCHECK();
		CU(cudaStreamSynchronize(s.stream));
		mx = fmax(mx, s.scal_host[8]); s1 += s.scal_host[9]; s2 += s.scal_host[10];
	}
	error[0] = mx; error[1] = s1; error[2] = multi ? sqrt(s2) : s2;
	return MGB_OK;
}

extern "C" int mgb_op_apply(mgb_engine *e, int level, int x_vec, int y_vec)
{
	NEED(e); TRY(require_ops(e, false)); TRY(check_vec(e, x_vec, level)); TRY(check_vec(e, y_vec, level));
	if (x_vec == y_vec) return fail(MGB_EINVAL, "x and y must differ");
	GHOSTS(e, level, x_vec);
	TRY(k_apply(e, level, x_vec, y_vec));
	return sync(e);
}
extern "C" int mgb_op_residual(mgb_engine *e, int level)
{
	NEED(e); TRY(require_ops(e, false)); TRY(check_vec(e, MGB_VEC_U, level));
	GHOSTS(e, level, MGB_VEC_U);
	TRY(k_residual(e, level, MGB_VEC_U, MGB_VEC_B, MGB_VEC_R));
	return sync(e);
}
extern "C" int mgb_op_residual_norm(mgb_engine *e, int level, double *norm)
{
	NEED(e); TRY(require_ops(e, false)); TRY(check_vec(e, MGB_VEC_U, level));
	GHOSTS(e, level, MGB_VEC_U);
	TRY(k_resnorm(e, level, MGB_VEC_U, MGB_VEC_B, 0));
	TRY(read_scalars(e, 0, 1)); TRY(sync(e));
	*norm = host_scal(e)[0];
	return MGB_OK;
}
extern "C" int mgb_op_smooth(mgb_engine *e, int level, const mgb_smoother *s, int its, int guess_zero)
{
	NEED(e); TRY(require_ops(e, false)); TRY(check_smoother(e, s)); TRY(check_vec(e, MGB_VEC_U, level));
	GHOSTS(e, level, MGB_VEC_U);
	TRY(smooth(e, level, s, its, guess_zero != 0, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W));
	return sync(e);
}
extern "C" int mgb_op_restrict(mgb_engine *e, int level, int fused)
{
	NEED(e); TRY(require_ops(e, true));
	if (level < 0 || level + 1 >= e->L) return fail(MGB_EINVAL, "level %d has no coarser level", level);
	GHOSTS(e, level, MGB_VEC_U); GHOSTS(e, level, MGB_VEC_B); GHOSTS(e, level, MGB_VEC_R);
	TRY(restrict_to_coarse(e, level, MGB_VEC_B, MGB_VEC_U, MGB_VEC_R, fused != 0));
	return sync(e);
}
extern "C" int mgb_op_prolong(mgb_engine *e, int level, int multadd)
{
	NEED(e); TRY(require_ops(e, true));
	if (level < 0 || level + 1 >= e->L) return fail(MGB_EINVAL, "level %d has no coarser level", level);
	GHOSTS(e, level + 1, MGB_VEC_U);
	TRY(prolong_add(e, level, MGB_VEC_U, multadd != 0));
	return sync(e);
}
extern "C" int mgb_op_norm2(mgb_engine *e, int which, int level, double *out)
{
	NEED(e); TRY(need_peers(e)); TRY(check_vec(e, which, level));
	TRY(k_reduce(e, level, which, -1, 0, 1)); TRY(read_scalars(e, 0, 1)); TRY(sync(e));
	*out = host_scal(e)[0]; return MGB_OK;
}
extern "C" int mgb_op_dot(mgb_engine *e, int xw, int yw, int level, double *out)
{
	NEED(e); TRY(need_peers(e)); TRY(check_vec(e, xw, level)); TRY(check_vec(e, yw, level));
	TRY(k_reduce(e, level, xw, yw, 0, 0)); TRY(read_scalars(e, 0, 1)); TRY(sync(e));
	*out = host_scal(e)[0]; return MGB_OK;
}
extern "C" int mgb_op_axpy(mgb_engine *e, int yw, double alpha, int xw, int level)
{
	NEED(e); TRY(check_vec(e, xw, level)); TRY(check_vec(e, yw, level));
	TRY(k_vecop<0>(e, level, yw, xw, alpha)); return sync(e);
}
extern "C" int mgb_op_aypx(mgb_engine *e, int yw, double beta, int xw, int level)
{
	NEED(e); TRY(check_vec(e, xw, level)); TRY(check_vec(e, yw, level));
	TRY(k_vecop<1>(e, level, yw, xw, beta)); return sync(e);
}

static int csr_spmv_dev(mgb_engine *e, const Csr *c, const double *x, int xn, int xp, double *y, int yn, int yp)
{
	k_csr_spmv<<<cdiv(c->m, 256), 256, 0, e->strips[0].stream>>>(c->rowptr, c->col, c->val, c->m, x, xn, xp, y, yn, yp);
	LAUNCHED(e); KCHECK(); return MGB_OK;
}
extern "C" int mgb_csr_spmv_vec(mgb_engine *e, int which, int level, int x_vec, int y_vec)
{
	const Csr *c = nullptr; TRY(pick_csr(e, -1, which, level, &c, nullptr));
	if (e->P > 1) return fail(MGB_EINVAL, "CSR SpMV runs on a single rank (strips apply the operator matrix-free)");
	const int lx = (which == MGB_MAT_PRO) ? level + 1 : level;
	const int ly = (which == MGB_MAT_RES) ? level + 1 : level;
	if (lx == ly && x_vec == y_vec) return fail(MGB_EINVAL, "x and y must differ");
	TRY(check_vec(e, x_vec, lx)); TRY(check_vec(e, y_vec, ly));
	Strip &s = e->strips[0];
	TRY(csr_spmv_dev(e, c, s.lev[lx].v[x_vec], e->geo[lx].nj, e->geo[lx].pitch, s.lev[ly].v[y_vec], e->geo[ly].nj, e->geo[ly].pitch));
	return sync(e);
}
extern "C" int mgb_csr_spmv(mgb_engine *e, int which, int level, const double *x, double *y)
{
	const Csr *c = nullptr; TRY(pick_csr(e, -1, which, level, &c, nullptr));
	if (e->P > 1) return fail(MGB_EINVAL, "CSR SpMV runs on a single rank (strips apply the operator matrix-free)");
	if (!x || !y) return fail(MGB_EINVAL, "null argument");
	cudaStream_t st = e->strips[0].stream;
	double *dx, *dy;
	CU(cudaMalloc(&dx, sizeof(double) * (size_t)c->n)); CU(cudaMalloc(&dy, sizeof(double) * (size_t)c->m));
	CU(cudaMemcpyAsync(dx, x, sizeof(double) * (size_t)c->n, cudaMemcpyHostToDevice, st));
	int r = csr_spmv_dev(e, c, dx, c->n, 0, dy, c->m, 0);   // dense vectors: one "row" of length n
	if (r == MGB_OK) {
		cudaError_t ce = cudaMemcpyAsync(y, dy, sizeof(double) * (size_t)c->m, cudaMemcpyDeviceToHost, st);
		if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
		if (ce != cudaSuccess) r = fail(MGB_ECUDA, "csr spmv copy back: %s", cudaGetErrorString(ce));
	}
	cudaFree(dx); cudaFree(dy);
	return r;
}

// ------------------------------------------------------------------------------------------------ cycle 0
// One V-cycle exactly as the body of the reference's while loop (ref: src/solver.c:1531-1546), followed by
// the fine residual norm into scal[0].
static int vcycle_body(mgb_engine *e, const mgb_vcycle_params *p, bool first)
{
	const int Lc = e->L;
	const mgb_smoother *s = &p->smoother;
	if (!p->no_fuse && fusable(e, s) && p->v0 >= 1 && (Lc == 1 || p->v1 >= 1)) {
		// the same cycle with each leg of a level done in one pass (mgb_fused.cuh): identical arithmetic per value.  Levels
		// on which the fused leg does not pay (fuse_level) take the one-kernel-per-operation sequence instead.
		const int B = MGB_VEC_B, U = MGB_VEC_U, W = MGB_VEC_W, R = MGB_VEC_R;
		const bool ik = e->P > 1 && e->inkernel;
		auto FL = [&](int l) { return fuse_level(e, s, l); };
		if (Lc == 1) {
			TRY(fused_leg(e, 0, s, p->v0, first ? PRE_ZERO : PRE_GIVEN, POST_NORM, B, U, W, 0));
		} else {
			// levels lp .. Lc-1: one persistent launch (Jacobi, or red-black SOR in the fusable sweep orders)
			const int lp = p->no_bottom ? Lc : bottom_start(e, s);
			for (int l = 0; l < Lc - 1 && l < lp; ++l) {                                                    // :1531-1536
				const bool zero = l > 0 || first;
				if (FL(l)) TRY(fused_leg(e, l, s, p->v0, zero ? PRE_ZERO : PRE_GIVEN, POST_RESTRICT, B, U, W, 0));
				else {
					if (ik && l == e->La) TRY(wait_gather(e, l));
					TRY(smooth(e, l, s, p->v0, zero, B, U, W));
					TRY(restrict_to_coarse(e, l, B, U, R, true));
				}
			}
			if (lp < Lc) {
				if (lp == e->La && ik) TRY(wait_gather(e, lp));        // the bottom kernel consumes the gathered right-hand side
				TRY(bottom_cycle(e, lp, s, p->v0, s, p->v1, false));
			} else if (FL(Lc - 1)) TRY(fused_leg(e, Lc - 1, s, p->v1, PRE_ZERO, POST_NONE, B, U, W, 0, true));    // :1536 coarsest
			else {
				if (ik && Lc - 1 == e->La) TRY(wait_gather(e, Lc - 1));
				TRY(smooth(e, Lc - 1, s, p->v1, true, B, U, W));
			}
			for (int l = (lp < Lc ? lp - 1 : Lc - 2); l >= 0; --l) {                                           // :1540-1546
				if (FL(l)) {
					if (e->geo[l].dist && !e->geo[l + 1].dist) TRY(need_bcast(e, l + 1));
					TRY(fused_leg(e, l, s, p->v0, PRE_PROLONG, l == 0 ? POST_NORM : POST_NONE, B, U, W, 0, true));
				} else {
					TRY(prolong_add(e, l, U, false));
					TRY(smooth(e, l, s, p->v0, false, B, U, W));
					if (l == 0) TRY(k_resnorm(e, 0, U, B, 0));
				}
			}
		}
		TRY(flush_all(e));                                     // scal[0] is already in the mapped host mirror (k_reduce_tail)
		return MGB_OK;
	}
	TRY(smooth(e, 0, s, p->v0, first, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W));                          // :1531-1532
	for (int l = 1; l < Lc; ++l) {
		TRY(restrict_to_coarse(e, l - 1, MGB_VEC_B, MGB_VEC_U, MGB_VEC_R, true));                 // :1534-1535
		TRY(smooth(e, l, s, (l == Lc - 1) ? p->v1 : p->v0, true, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W)); // :1536
	}
	for (int l = Lc - 2; l >= 0; --l) {
		TRY(prolong_add(e, l, MGB_VEC_U, false));                                                 // :1540-1541
		TRY(smooth(e, l, s, p->v0, false, MGB_VEC_B, MGB_VEC_U, MGB_VEC_W));                       // :1542
	}
	TRY(k_resnorm(e, 0, MGB_VEC_U, MGB_VEC_B, 0));                                                // :1545-1546
	TRY(flush_all(e));
	return MGB_OK;
}

static void drop_graphs(mgb_engine *e)
{
	for (auto &g : e->gcache) cudaGraphExecDestroy(g.exec);
	e->gcache.clear();
}

static void save_state(mgb_engine *e, std::vector<PtrState> &st)
{
	st.resize(e->strips.size());
	for (size_t i = 0; i < e->strips.size(); ++i)
		for (int l = 0; l < e->L; ++l)
			for (int k = 0; k < MGB_NVEC; ++k) { st[i].v[l][k] = e->strips[i].lev[l].v[k]; st[i].phys[l][k] = e->strips[i].lev[l].phys[k]; }
}
static void load_state(mgb_engine *e, const std::vector<PtrState> &st)
{
	for (size_t i = 0; i < e->strips.size(); ++i)
		for (int l = 0; l < e->L; ++l)
			for (int k = 0; k < MGB_NVEC; ++k) { e->strips[i].lev[l].v[k] = st[i].v[l][k]; e->strips[i].lev[l].phys[k] = st[i].phys[l][k]; }
}
// Replay `body` (kernel launches on the strips' stream, no synchronisation) as a CUDA graph.  A captured body is valid for one
// set of parameters (`key`) and one assignment of the ping-pong pointers, which is appended to the key; the pointer state
// after the body is a pure function of the state before it and is restored on replay.
template <class F>
static int run_graphed(mgb_engine *e, std::vector<char> key, F body)
{
	Strip &s0 = e->strips[0];
	for (auto &s : e->strips)
		for (int l = 0; l < e->L; ++l) {
			const char *b = (const char *)s.lev[l].v;
			key.insert(key.end(), b, b + sizeof s.lev[l].v);
		}
	GraphEntry *g = nullptr;
	for (auto &c : e->gcache) if (c.key == key) { g = &c; break; }
	if (!g) {
		cudaGraph_t graph = nullptr;
		const long long l0 = e->launches;
		std::vector<PtrState> before; save_state(e, before);
		CU(cudaStreamBeginCapture(s0.stream, cudaStreamCaptureModeThreadLocal));
		int r = body();
		cudaError_t ce = cudaStreamEndCapture(s0.stream, &graph);     // always ends the capture, also after a failed body
		if (r != MGB_OK || ce != cudaSuccess) {
			if (graph) cudaGraphDestroy(graph);
			load_state(e, before); e->launches = l0; e->pending.clear();
			if (r != MGB_OK) return r;
			return fail(MGB_ECUDA, "graph capture failed: %s", cudaGetErrorString(ce));
		}
		GraphEntry ne; ne.key = key; ne.exec = nullptr;
		ce = cudaGraphInstantiate(&ne.exec, graph, 0);
		cudaGraphDestroy(graph);
		if (ce != cudaSuccess) { load_state(e, before); e->launches = l0; return fail(MGB_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); }
		ne.launches = e->launches - l0;
		e->launches = l0;
		save_state(e, ne.after);
		e->gcache.push_back(ne);
		g = &e->gcache.back();
	} else {
		load_state(e, g->after);
	}
	CU(cudaGraphLaunch(g->exec, s0.stream));
	e->launches += g->launches;
	return MGB_OK;
}

// the cycle loop of MultigridVcycle on the right-hand side in B[0] (ref: src/solver.c:1512-1558); rnorm holds
// max_iter+1 entries and is returned normalised by rnorm[0]
static int vcycle_solve(mgb_engine *e, const mgb_vcycle_params *p, double *rnorm, int *num_iter, bool mark_loop_start)
{
	Strip &s0 = e->strips[0];
	// bnorm = ||b0|| ; u0 = 0 ; rnorm[0] = ||A0 u0 - b0||                                      (:1512-1520)
	TRY(halo(e, 0, MGB_VEC_B, HALO_DEPTH));
	TRY(k_reduce(e, 0, MGB_VEC_B, -1, 1, 1));
	TRY(vec_zero(e, MGB_VEC_U, 0));
	TRY(k_resnorm(e, 0, MGB_VEC_U, MGB_VEC_B, 0));
	TRY(read_scalars(e, 0, 2));
	const double bnorm = host_scal(e)[1];
	double rn = host_scal(e)[0];
	rnorm[0] = rn;
	int iter = 0;
	if (e->gcache.size() > 16) drop_graphs(e);
	// the timed region of the reference starts here, after rnorm[0] (t0 = MPI_Wtime() at src/solver.c:1526)
	if (mark_loop_start) CU(cudaEventRecord(s0.ev0, s0.stream));
	while (iter < p->max_iter && 100000000.0 * bnorm > rn && rn > p->rtol * bnorm) {              // :1530
		if (!p->use_graph) {
			TRY(vcycle_body(e, p, iter == 0));
		} else {
			const bool first = iter == 0;                            // the first cycle starts from u = 0 (PRE_ZERO leg): its own graph
			std::vector<char> key(sizeof *p + 1, first ? 'F' : 'V');
			memcpy(key.data() + 1, p, sizeof *p);
			TRY(run_graphed(e, key, [&]() { return vcycle_body(e, p, first); }));
		}
		CU(cudaStreamSynchronize(s0.stream));
		TRY(quick_status(e));
		rn = s0.scal_host[0];
		iter = iter + 1;
		rnorm[iter] = rn;
	}
	const double r0 = rnorm[0];
	for (int i = 0; i <= iter; ++i) rnorm[i] = rnorm[i] / r0;                                     // :1554-1557
	*num_iter = iter;
	return MGB_OK;
}

static int vcycle_check(mgb_engine *e, const mgb_vcycle_params *p)
{
	TRY(require_ops(e, true));
	if (e->dead) return fail(MGB_ESTATE, "the engine is unusable after a timed-out strip exchange: destroy it on every rank"); TRY(check_smoother(e, &p->smoother));
	if (p->v0 < 0 || p->v1 < 0 || p->max_iter < 0) return fail(MGB_EINVAL, "negative sweep or iteration count");
	return MGB_OK;
}

extern "C" int mgb_solve_vcycle(mgb_engine *e, const mgb_vcycle_params *p, double *rnorm, int *num_iter, double *seconds)
{
	NEED(e);
	if (!p || !rnorm || !num_iter) return fail(MGB_EINVAL, "null argument");
	TRY(vcycle_check(e, p));
	Strip &s0 = e->strips[0];
	TRY(sync(e));
	const auto t0 = std::chrono::steady_clock::now();
	TRY(vcycle_solve(e, p, rnorm, num_iter, true));
	CU(cudaEventRecord(s0.ev1, s0.stream));
	CU(cudaStreamSynchronize(s0.stream));
	const auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	{ float ms = 0.f; CU(cudaEventElapsedTime(&ms, s0.ev0, s0.ev1)); e->last_solve_ms = ms; }
	return check_status(e);
}

// A stream of independent right-hand sides (host arrays, whole-grid, natural order; pinned memory for real overlap):
// while solve k runs, the right-hand side of solve k+1 is uploaded into a spare level-0 vector and the solution of
// solve k-1 is downloaded from another one, on two copy streams (both PCIe directions busy, the SMs never wait).
// The reference's driver solves one right-hand side per process run (ref: src/poisson.c:118-125: Assemble, Solve,
// GetSol); this is the same Solve() applied to many, with the operator assembled once.
extern "C" int mgb_solve_vcycle_many(mgb_engine *e, const mgb_vcycle_params *p, int nrhs, const double *const *b_hosts,
                                     double *const *u_hosts, int *num_iter, double *final_rnorm, double *seconds)
{
	NEED(e);
	if (!p || nrhs < 1 || !b_hosts || !u_hosts) return fail(MGB_EINVAL, "null argument");
	TRY(vcycle_check(e, p));
	const LevelGeom &g = e->geo[0];
	const int SPARE_B = MGB_VEC_Q;                           // a Krylov work vector, idle in cycle 0
	std::vector<double> rn((size_t)p->max_iter + 2);
	auto upload = [&](const double *host, int which, bool async) -> int {
		for (auto &s : e->strips) {
			SLevel &S = s.lev[0];
			TRY(put_rows(e, s, 0, which, host + (size_t)S.r0 * g.nj, async ? s.copy_in : s.stream));
			if (async) CU(cudaEventRecord(s.ev_in, s.copy_in));
		}
		return MGB_OK;
	};
	TRY(sync(e));
	const auto t0 = std::chrono::steady_clock::now();
	const bool trace = getenv("MGB_TRACE") != nullptr;
	auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
	TRY(upload(b_hosts[0], MGB_VEC_B, false));
	for (int k = 0; k < nrhs; ++k) {
		if (trace) fprintf(stderr, "[mgb] rhs %d: t=%.2f ms upload of next issued\n", k, now_ms());
		if (k + 1 < nrhs) TRY(upload(b_hosts[k + 1], SPARE_B, true));
		int it = 0;
		if (trace) fprintf(stderr, "[mgb] rhs %d: t=%.2f ms solve starts\n", k, now_ms());
		TRY(vcycle_solve(e, p, rn.data(), &it, false));
		if (trace) fprintf(stderr, "[mgb] rhs %d: t=%.2f ms solve done (%d cycles)\n", k, now_ms(), it);
		if (num_iter) num_iter[k] = it;
		if (final_rnorm) final_rnorm[k] = rn[it];
		// solution k: packed into the dense staging buffer on the compute stream (after the previous download has left
		// it), then device-to-host on the second copy stream
		for (auto &s : e->strips) {
			SLevel &S = s.lev[0];
			CU(cudaStreamWaitEvent(s.stream, s.ev_out, 0));
			TRY(get_rows(e, s, 0, MGB_VEC_U, u_hosts[k] + (size_t)S.r0 * g.nj, s.stream, s.copy_out, s.ev_sol));
			CU(cudaEventRecord(s.ev_out, s.copy_out));
			// the next right-hand side must have arrived before it becomes B
			if (k + 1 < nrhs) CU(cudaStreamWaitEvent(s.stream, s.ev_in, 0));
		}
		if (k + 1 < nrhs) swap_vec(e, 0, MGB_VEC_B, SPARE_B);
	}
	for (auto &s : e->strips) { CU(cudaStreamSynchronize(s.copy_out)); CU(cudaStreamSynchronize(s.stream)); }
	const auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return check_status(e);
}

// ------------------------------------------------------------------------------------------------ cycle 8
// PCApply_MG (multiplicative V, one cycle, x = 0 on entry) on level l with right-hand side bv and iterate xv.
// dot_slot >= 0 (level 0 only): also leave scal[dot_slot] = xv . bv (the z'r of CG, with TAIL_BETA) when the last leg can
// produce it in passing (*dot_done = true), else the caller computes it
static int pcmg_cycle(mgb_engine *e, const mgb_pcmg_params *p, int l, int bv, int xv, int dot_slot = -1, bool *dot_done = nullptr)
{
	const int Lc = e->L;
	// the gathered right-hand side of the first agglomerated level: the fused legs wait for it themselves, every other consumer
	// (bottom kernel, LU, one-sweep kernels) gets a wait launch on rank 0 (harmless when a fused leg follows)
	if (e->P > 1 && e->inkernel && l == e->La && l >= 1) TRY(wait_gather(e, l));
	if (l >= 1 && !p->no_fuse && !p->no_bottom && p->coarse == MGB_COARSE_RICHARDSON && fusable(e, &p->level_smoother) &&
	    fusable(e, &p->coarse_smoother) && p->level_smoother.type == p->coarse_smoother.type &&
	    p->level_its >= 1 && p->coarse_its >= 1 && l == bottom_start(e, &p->level_smoother) &&
	    bv == MGB_VEC_B && xv == MGB_VEC_U)
		return bottom_cycle(e, l, &p->level_smoother, p->level_its, &p->coarse_smoother, p->coarse_its, true);
	if (l == Lc - 1) {
		if (p->coarse == MGB_COARSE_RICHARDSON) {
			if (!p->no_fuse && fusable(e, &p->coarse_smoother) && fuse_level(e, &p->coarse_smoother, l) && p->coarse_its >= 1)
				return fused_leg(e, l, &p->coarse_smoother, p->coarse_its, PRE_ZERO, POST_NONE, bv, xv, MGB_VEC_W, 0, true);
			return smooth(e, l, &p->coarse_smoother, p->coarse_its, true, bv, xv, MGB_VEC_W);
		}
		TRY(flush_levels(e, l, l));
		for (auto &s : e->strips) {
			if (!computes(s, l)) continue;
			SLevel &S = s.lev[l];
			if (bandlu_solve(S.lu, S.v[bv], S.v[xv], S.ni, e->geo[l].nj, e->geo[l].pitch, s.stream)) return fail(MGB_ECUDA, "coarse LU solve launch failed");
			LAUNCHED(e);
		}
		return MGB_OK;
	}
	if (!p->no_fuse && fusable(e, &p->level_smoother) && fuse_level(e, &p->level_smoother, l) && p->level_its >= 1) {
		TRY(fused_leg(e, l, &p->level_smoother, p->level_its, PRE_ZERO, POST_RESTRICT, bv, xv, MGB_VEC_W, 0));
		TRY(pcmg_cycle(e, p, l + 1, MGB_VEC_B, MGB_VEC_U));
		if (e->geo[l].dist && !e->geo[l + 1].dist) TRY(need_bcast(e, l + 1));
		const bool dot = l == 0 && dot_slot >= 0 && dot_done && p->level_smoother.type == MGB_SMOOTH_JACOBI && p->level_its <= FJ_MAXD;
		TRY(fused_leg(e, l, &p->level_smoother, p->level_its, PRE_PROLONG_MULTADD, dot ? POST_DOT : POST_NONE, bv, xv, MGB_VEC_W, dot ? dot_slot : 0, true));
		if (dot) *dot_done = true;
		return MGB_OK;
	}
	TRY(smooth(e, l, &p->level_smoother, p->level_its, true, bv, xv, MGB_VEC_W));     // pre-smooth from x = 0
	TRY(restrict_to_coarse(e, l, bv, xv, MGB_VEC_R, true));                           // b_c = R (b - A x)
	TRY(pcmg_cycle(e, p, l + 1, MGB_VEC_B, MGB_VEC_U));                               // x_c = 0 ; recurse
	TRY(prolong_add(e, l, xv, true));                                                 // x = x + P x_c (MatMultAdd)
	TRY(smooth(e, l, &p->level_smoother, p->level_its, false, bv, xv, MGB_VEC_W));    // post-smooth
	return MGB_OK;
}

// KSPConvergedDefault
static int ksp_converged(const mgb_pcmg_params *p, int it, double rn, double *rnorm0, double *ttol)
{
	if (it == 0) { *rnorm0 = rn; *ttol = fmax(p->rtol * rn, p->abstol); }
	if (rn != rn || rn - rn != 0.0) return -9;     // KSP_DIVERGED_NANORINF
	if (rn <= *ttol) return (rn < p->abstol) ? 3 : 2;
	if (rn >= p->dtol * (*rnorm0)) return -4;
	return 0;
}

extern "C" int mgb_solve_pcmg(mgb_engine *e, const mgb_pcmg_params *p, double *rnorm, int *num_iter, int *reason_out, double *seconds)
{
	NEED(e);
	if (!p || !rnorm || !num_iter) return fail(MGB_EINVAL, "null argument");
	TRY(require_ops(e, true));
	if (e->dead) return fail(MGB_ESTATE, "the engine is unusable after a timed-out strip exchange: destroy it on every rank");
	const int Lc = e->L;
	if (Lc < 2) return fail(MGB_EINVAL, "cycle 8 needs at least two levels");
	TRY(check_smoother(e, &p->level_smoother));
	if (p->coarse == MGB_COARSE_RICHARDSON) {
		TRY(check_smoother(e, &p->coarse_smoother));
		if (p->coarse_smoother.type == MGB_SMOOTH_RBSOR && p->level_smoother.type == MGB_SMOOTH_RBSOR &&
		    p->coarse_smoother.omega != p->level_smoother.omega)
			return fail(MGB_EINVAL, "different SOR omegas on levels and coarse grid are not supported");
	} else if (p->coarse == MGB_COARSE_LU) {
		if (e->geo[Lc - 1].dist) return fail(MGB_EINVAL, "coarse LU needs the coarsest level agglomerated on rank 0");
		for (auto &s : e->strips) {
			if (!computes(s, Lc - 1)) continue;
			SLevel &C = s.lev[Lc - 1];
			if (bandlu_factor(C.lu, e->geo[Lc - 1].coef_host.data(), C.ni, e->geo[Lc - 1].nj, g_err, sizeof g_err)) return MGB_EINVAL;
		}
	} else return fail(MGB_EINVAL, "unknown coarse solver %d", p->coarse);
	if (p->outer != MGB_KSP_CG && p->outer != MGB_KSP_RICHARDSON) return fail(MGB_EINVAL, "unknown outer KSP %d", p->outer);

	const int X = MGB_VEC_U, B = MGB_VEC_B, R = MGB_VEC_R, Z = MGB_VEC_Z, Pv = MGB_VEC_P, Q = MGB_VEC_Q;
	Strip &s0 = e->strips[0];
	TRY(sync_all(e));
	const auto t0 = std::chrono::steady_clock::now();
	CU(cudaEventRecord(s0.ev0, s0.stream));
	int reason = 0, its = 0, nlog = 0;
	double rnorm0 = 0.0, ttol = 0.0, dp = 0.0;
	auto logr = [&](double v) { if (nlog < p->max_iter) rnorm[nlog++] = v; };   // KSPSetResidualHistory(na = numIter)
	for (int i = 0; i <= p->max_iter; ++i) rnorm[i] = NAN;
	double *hs = host_scal(e);
	// z = B r : one multigrid cycle on (r, z); the ghost rows of r are refreshed first (the restriction reads them)
	auto precond = [&]() -> int { TRY(halo(e, 0, R, HALO_DEPTH)); return pcmg_cycle(e, p, 0, R, Z); };

	TRY(vec_zero(e, X, 0));                                  // KSPSolve: zero initial guess
	TRY(k_vecop<2>(e, 0, R, B, 0.0));                        // r = b
	if (p->outer == MGB_KSP_CG) {
		// KSPSolve_CG, KSP_NORM_UNPRECONDITIONED (ref: src/solver.c:1922).  One iteration is ONE graph launch and ONE host
		// read-back (||r||): beta = z'r, b = beta / betaold, dpi = p'w and a = beta / dpi are derived on the device where the
		// dot products finish (k_reduce_tail, TAIL_*), the vector kernels read them from HBM.  The host sees all of them in
		// the mapped mirror after the synchronisation and applies PETSc's tests in PETSc's order.  On a beta == 0 /
		// indefinite-preconditioner / indefinite-operator exit PETSc leaves before x is updated: with the deferred x update
		// (k_cg_pstep) x is exactly PETSc's; only the internal work vector r has taken the update of that iteration (with
		// MGB_CG_FUSE=0 x has taken it too).
		double beta = 0.0, betaold = 1.0, dpi = 0.0, dpiold;
		TRY(k_reduce(e, 0, R, -1, 0, 1)); TRY(read_scalars(e, 0, 1)); dp = hs[0];
		logr(dp);
		reason = ksp_converged(p, 0, dp, &rnorm0, &ttol);
		if (!reason) {
			TRY(vec_zero(e, Pv, 0));                                                               // p = 0: the first p = z + 0 p
			for (auto &s : e->strips) { klaunch(k_set_scalar, dim3(1), dim3(1), 0, s.stream, s.scal + SC_BETA, INFINITY); LAUNCHED(e); KCHECK(); }
			// the direction step fused with the operator apply and the deferred x update (k_cg_pstep); MGB_CG_FUSE=0: separate passes
			const bool fuse_dir = e->cg_fuse;
			if (fuse_dir) for (auto &s : e->strips) { klaunch(k_set_scalar, dim3(1), dim3(1), 0, s.stream, s.scal + SC_ALPHA, 0.0); LAUNCHED(e); KCHECK(); }
			bool ran = false;
			auto iteration = [&]() -> int {
				bool dot_done = false;
				TRY(halo(e, 0, R, HALO_DEPTH));
				TRY(pcmg_cycle(e, p, 0, R, Z, 2, &dot_done));                                      // z = B r  [; beta = z'r]
				if (!dot_done) TRY(k_reduce(e, 0, Z, R, 2, 0, TAIL_BETA));                         // beta = z'r ; b = beta / betaold
				if (fuse_dir) {
					// x += a_prev p ; p = z + b p ; w = A p ; dpi = p'w ; a = beta / dpi -- one pass, p's ghost rows derived locally
					TRY(wait_halo(e, 0, Z));
					TRY(k_cg_dir(e, 0, Z, Pv, MGB_VEC_W, Q, X, 3, TAIL_DPI));
					TRY(k_cg_step(e, 0, -1, Pv, R, Q, 0.0, 0, SC_ALPHA));                              // r -= a w ; ||r||
				} else {
					TRY(k_vecop<1>(e, 0, Pv, Z, 0.0, SC_RATIO));                                       // p = z + b p
					TRY(halo(e, 0, Pv, 2));
					TRY(k_apply_dot(e, 0, Pv, Q, 3, TAIL_DPI));                                        // w = A p ; dpi = p'w ; a = beta / dpi
					TRY(k_cg_step(e, 0, X, Pv, R, Q, 0.0, 0, SC_ALPHA));                               // x += a p ; r -= a w ; ||r||
				}
				return flush_all(e);
			};
			if (e->gcache.size() > 16) drop_graphs(e);
			int i = 0;
			do {
				its = i + 1;
				if (p->no_graph) TRY(iteration());
				else {
					std::vector<char> key(sizeof *p + 1, 'C');
					memcpy(key.data() + 1, p, sizeof *p);
					TRY(run_graphed(e, key, iteration));
				}
				CU(cudaStreamSynchronize(s0.stream));
				TRY(quick_status(e));
				ran = true;
				betaold = beta; dpiold = dpi;
				beta = hs[SC_BETA]; dpi = hs[SC_DPI]; dp = hs[0];
				// PETSc leaves these three exits BEFORE x takes the update of the iteration: with the deferred x update that is
				// exactly what has happened here (the update of this iteration is still pending and is dropped)
				if (beta == 0.0) { reason = 3; ran = false; break; }
				else if (i > 0 && beta * betaold < 0.0) { reason = -8; ran = false; break; }       // KSP_DIVERGED_INDEFINITE_PC
				if (dpi == 0.0 || (i > 0 && dpi * dpiold <= 0.0)) { reason = -10; ran = false; break; }   // KSP_DIVERGED_INDEFINITE_MAT
				logr(dp);
				reason = ksp_converged(p, i + 1, dp, &rnorm0, &ttol);
				if (reason) break;
				i++;
			} while (i < p->max_iter);
			if (i >= p->max_iter && !reason) reason = -3;                                          // KSP_DIVERGED_ITS
			if (fuse_dir && ran) TRY(k_vecop<0>(e, 0, X, Pv, 0.0, SC_ALPHA));                      // the pending x += a p of the last iteration
		}
	} else {
		// KSPSolve_Richardson, general path (residual norm logged every iteration), scale 1
		for (int i = 0; i < p->max_iter; ++i) {
			TRY(k_reduce(e, 0, R, -1, 0, 1)); TRY(read_scalars(e, 0, 1)); dp = hs[0];
			logr(dp);
			reason = ksp_converged(p, i, dp, &rnorm0, &ttol);
			if (reason) break;
			TRY(precond());                                                                        // z = B r
			TRY(k_vecop<0>(e, 0, X, Z, 1.0));                                                      // x = x + scale z
			its++;
			TRY(halo(e, 0, X, 2));
			TRY(k_residual(e, 0, X, B, R));                                                        // r = b - A x
		}
		if (!reason) {
			TRY(k_reduce(e, 0, R, -1, 0, 1)); TRY(read_scalars(e, 0, 1)); dp = hs[0];
			logr(dp);
			if (its >= p->max_iter) { reason = ksp_converged(p, its, dp, &rnorm0, &ttol); if (!reason) reason = -3; }
		}
	}
	TRY(flush_all(e));
	CU(cudaEventRecord(s0.ev1, s0.stream));
	TRY(sync_all(e));
	const auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	{ float ms = 0.f; CU(cudaEventElapsedTime(&ms, s0.ev0, s0.ev1)); e->last_solve_ms = ms; }
	TRY(check_status(e));
	// ref: src/solver.c:1971-1976 -- numIter = KSPGetIterationNumber ; rnorm[i] /= rnorm[0]
	const double r0 = rnorm[0];
	for (int i = 0; i < its + 1 && i <= p->max_iter; ++i) rnorm[i] = rnorm[i] / r0;
	*num_iter = its;
	if (reason_out) *reason_out = reason;
	return MGB_OK;
}

// ------------------------------------------------------------------------------------------------ measurement
extern "C" int mgb_time_op(mgb_engine *e, int op, int level, int reps, double *ms_per_launch)
{
	NEED(e); TRY(require_ops(e, true));
	if (level < 0 || level >= e->L) return fail(MGB_EINVAL, "level %d out of range", level);
	if (reps < 1 || !ms_per_launch) return fail(MGB_EINVAL, "bad reps / null output");
	if (e->strips.size() != 1) return fail(MGB_EINVAL, "mgb_time_op times one strip per process");
	mgb_smoother jac = {MGB_SMOOTH_JACOBI, 0.8, 1.0, 0, 1};
	const bool has_coarse = level + 1 < e->L;
	if ((op == 4 || op == 5 || op == 11 || op == 12 || op == 15) && !has_coarse) return fail(MGB_EINVAL, "level %d has no coarser level", level);
	if (op == 7 && !e->csr_built) return fail(MGB_ESTATE, "mgb_assemble_csr was not called");
	TRY(set_sor_omega(e, 1.0));
	Strip &s = e->strips[0];
	const int U = MGB_VEC_U, B = MGB_VEC_B, R = MGB_VEC_R;
	for (int pass = 0; pass < 2; ++pass) {
		const int n = pass == 0 ? 2 : reps;                   // pass 0: warm-up
		if (pass == 1) CU(cudaEventRecord(s.ev0, s.stream));
		for (int k = 0; k < n; ++k) {
			switch (op) {
			case 0: TRY(k_apply(e, level, U, R)); break;
			case 1: TRY(k_residual(e, level, U, B, R)); break;
			case 2: TRY(smooth(e, level, &jac, 1, false, B, U, MGB_VEC_W)); break;
			case 3: TRY(k_rb(e, level, U, B, 0, 1.0, 0)); TRY(k_rb(e, level, U, B, 1, 1.0, 0)); break;
			case 4: TRY(restrict_to_coarse(e, level, B, U, R, true)); break;
			case 5: TRY(prolong_add(e, level, U, false)); break;
			case 6: TRY(k_resnorm(e, level, U, B, 0)); break;
			case 7: { SLevel &S = s.lev[level]; TRY(csr_spmv_dev(e, &S.A, S.v[U], e->geo[level].nj, e->geo[level].pitch, S.v[R], e->geo[level].nj, e->geo[level].pitch)); } break;
			case 8: TRY(k_reduce(e, level, U, -1, 0, 1)); break;
			case 9: TRY(k_reduce(e, level, U, B, 0, 0)); break;
			case 10: TRY(k_vecop<0>(e, level, R, U, 0.5)); break;
			case 11: TRY(fused_leg(e, level, &jac, 3, PRE_GIVEN, POST_RESTRICT, B, U, MGB_VEC_W, 0)); break;
			case 12: TRY(fused_leg(e, level, &jac, 3, PRE_PROLONG, POST_NORM, B, U, MGB_VEC_W, 0)); break;
			case 13: TRY(fused_leg(e, level, &jac, 3, PRE_GIVEN, POST_NONE, B, U, MGB_VEC_W, 0)); break;
			case 14: TRY(fused_leg(e, level, &jac, 1, PRE_GIVEN, POST_NONE, B, U, MGB_VEC_W, 0)); break;
			case 15: TRY(fused_leg(e, level, &jac, 3, PRE_ZERO, POST_RESTRICT, B, U, MGB_VEC_W, 0)); break;
			case 17: TRY(halo(e, level, U, HALO_DEPTH)); TRY(flush_all(e)); break;
			case 18: TRY(k_reduce(e, level, U, -1, 0, 1)); break;
			case 16: if (level < 1) return fail(MGB_EINVAL, "the bottom kernel starts at level >= 1");
			         TRY(bottom_cycle(e, level, &jac, 3, &jac, 3, false)); break;
			default: return fail(MGB_EINVAL, "unknown op %d", op);
			}
			TRY(flush_all(e));               // transfers requested by the operation belong to its cost
		}
		if (pass == 1) CU(cudaEventRecord(s.ev1, s.stream));
		CU(cudaStreamSynchronize(s.stream));
	}
	float ms = 0.f;
	CU(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
	*ms_per_launch = (double)ms / reps;
	return check_status(e);
}
