/*
 * solver_b200.c -- replacement of the reference's src/solver.c for the accelerated path.
 *
 * Same entry points, same argument meaning, same call order as include/solver.h:81-98; PETSc's CPU linear
 * algebra is replaced by the sm_100a engine behind include/mgb200.h (libmgb200.so).  There is no CPU path
 * here: when the engine cannot be created (no GPU, library missing) the run stops with an error.
 *
 *   SetUpSolver / DestroySolver         ref: src/solver.c:107-149 (+ SetUpAssembly/DestroyAssembly :33-105)
 *   Assemble                            ref: src/solver.c:1156-1209 -> levelMatrixA/fillJacobians :185-253,489-510,
 *                                            levelvecb :558-620, Res :1035-1094, Pro :1096-1154
 *   Solve                               ref: src/solver.c:2617-2630 -> MultigridVcycle :1414-1575 (cycle 0),
 *                                            MultigridPetscPCMG :1884-1989 (cycle 8)
 *   SetUpPostProcess / Postprocessing / DestroyPostProcess   ref: src/solver.c:151-183, 1211-1380
 *
 * Solver options are read from the options database exactly where the reference's KSPSetFromOptions calls
 * would read them (un-prefixed keys for cycle 0, mg_levels_ / mg_coarse_ prefixes for cycle 8).
 * Supported: -pc_type jacobi (any -ksp_richardson_scale) ; -pc_type sor (-pc_sor_omega, -pc_sor_its, -pc_sor_lits,
 * -pc_sor_forward/backward/symmetric): PETSc's lexicographic MatSOR on the natural numbering (wavefront kernels, one GPU)
 * or the same arithmetic on the red-black numbering with -map 3 ; -pc_type ilu or NO -pc_type: PETSc's default ILU(0)
 * (wavefront kernels, one GPU) -- the shipped poisson.in runs unmodified ; -ksp_max_it ;
 * cycle 8: -ksp_type richardson|cg, -ksp_rtol, -ksp_atol, -ksp_divtol, -ksp_max_it, -mg_levels_*, -mg_coarse_*.
 * Extensions: -rtol (cycle-0 tolerance, 1e-7 in the reference), -mgb_graph 0|1, -mgb_csr 0|1.
 * Everything else on this path (Chebyshev, the research cycles 1-7, 9, 10, several grids per level) is refused with a
 * message, never approximated.
 */
#include "pb_api.h"
#include "mgb200.h"
#include <string.h>
#include <time.h>

/* the engine handle lives in the Assembly slot the accelerated path never uses (A2: second-operator array of
 * the delayed cycles), so the struct layout of include/solver.h:39-52 is untouched */
static mgb_engine *ENGINE(Assembly *assem) { return (mgb_engine *)(void *)assem->A2; }
static void set_engine(Assembly *assem, mgb_engine *e) { assem->A2 = (Mat *)(void *)e; }

static __thread char g_msg[600];      /* per thread: two sessions on two threads keep their own message and trap */
const char *pb200_last_error(void) { return g_msg; }

/* Error convention of the reference: its API returns void and prints (ERROR_MSG, src/solver.c:3-6).  Here an
 * error on the accelerated path is fatal: print and exit(1) -- or, when the pipeline runs as a library call
 * (pb200_run), unwind to it through pb200_trap so that the caller gets a status and the message. */
#include <setjmp.h>
__thread jmp_buf *pb200_trap = NULL;

static void bail(void)
{
	fprintf(stderr, "poisson_b200 ERROR: %s\n", g_msg);
	if (pb200_trap) longjmp(*pb200_trap, 1);
	exit(1);
}
static void die(const char *what)
{
	char tmp[600];
	snprintf(tmp, sizeof tmp, "%s: %s", what, mgb_last_error());
	memcpy(g_msg, tmp, sizeof g_msg);
	bail();
}
static void refuse(const char *why)
{
	if (why != g_msg) snprintf(g_msg, sizeof g_msg, "%s", why);
	bail();
}

struct mgb_engine *pb200_engine(Solver *solver) { return solver && solver->assem ? ENGINE(solver->assem) : NULL; }

/* ---------------------------------------------------------------- setup / teardown */
void SetUpSolver(Indices *indices, Solver *solver, Cycle cyc)
{
	solver->cycle = cyc;
	solver->assem = calloc(1, sizeof(Assembly));
	solver->assem->levels = indices->levels;
	solver->assem->moreInfo = solver->moreInfo;
	solver->rnorm = malloc(((size_t)solver->numIter + 1) * sizeof(double));
	solver->grids = 0; solver->rNormGrid = NULL; solver->rNormGlobal = NULL;
	if (cyc != VCYCLE && cyc != PetscPCMG)
		refuse("only -cycle 0 (V-cycle) and -cycle 8 (PCMG) run on the B200 engine; the research cycles stay with the PETSc build");
	if (solver->moreInfo != 0)
		refuse("-moreNorm 1 only concerns the delayed cycles (D1/D2/D1PS), which are not on the accelerated path");
}

void DestroySolver(Solver *solver)
{
	if (!solver->assem) return;
	mgb_destroy(ENGINE(solver->assem));
	free(solver->assem); solver->assem = NULL;
	free(solver->rnorm); solver->rnorm = NULL;
}

/* ---------------------------------------------------------------- assembly */
static int int_pow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

void Assemble(Problem *prob, Mesh *mesh, Indices *indices, Operator *op, Solver *solver)
{
	Assembly *assem = solver->assem;
	const int L = assem->levels;
	int map_style = 2, want_csr = 1;
	pbopt_get_int("-map", &map_style);
	pbopt_get_int("-mgb_csr", &want_csr);
	for (int l = 0; l < L; l++)
		if (indices->level[l].grids != 1)
			refuse("the B200 engine handles one grid per level (-grids must equal -levels)");

	mgb_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.levels = L;
	cfg.ni = indices->level[0].grid[0].ni;
	cfg.nj = indices->level[0].grid[0].nj;
	cfg.device = -1;
	cfg.red_black_numbering = (map_style == 3);
	/* row strips (B200 extension; replaces mpiexec -n P): -mgb_ranks P [-mgb_rank r] [-mgb_agglomerate n]
	 * [-mgb_emulate 1: all strips in this process on one GPU, for tests] */
	cfg.rank = 0; cfg.nranks = 1;
	pbopt_get_int("-mgb_ranks", &cfg.nranks);
	pbopt_get_int("-mgb_rank", &cfg.rank);
	pbopt_get_int("-mgb_agglomerate", &cfg.agglomerate_below);
	pbopt_get_int("-mgb_emulate", &cfg.emulate);
	pbopt_get_int("-mgb_device", &cfg.device);
	mgb_engine *e = NULL;
	if (mgb_create(&cfg, &e) != MGB_OK) die("mgb_create");
	set_engine(assem, e);

	/* operator: one OpA evaluation per grid row (the reference does it per matrix row, src/solver.c:231-236;
	 * its three metric functions depend on y only, which is checked here rather than assumed) */
	for (int l = 0; l < L; l++) {
		Level *lv = &indices->level[l];
		const int g = lv->gridId[0], ni = lv->grid[0].ni, nj = lv->grid[0].nj;
		int eni, enj;
		mgb_level_dims(e, l, &eni, &enj);
		if (eni != ni || enj != nj) refuse("level sizes of Indices and engine disagree");
		const int f = int_pow(indices->coarseningFactor, g);
		double *rows = malloc((size_t)ni * 5 * sizeof(double));
		for (int i = 0; i < ni; i++) {
			const int ifine = f * (i + 1) - 1;
			const int probe[3] = {0, nj / 2, nj - 1};
			for (int k = 0; k < 3; k++) {
				const int jfine = f * (probe[k] + 1) - 1;
				double metrics[5], As[5];
				mesh->MetricCoefficients(mesh, mesh->coord[0][jfine + 1], mesh->coord[1][ifine + 1], metrics);
				prob->OpA(As, metrics, lv->h[0]);
				if (k == 0) memcpy(rows + (size_t)i * 5, As, sizeof As);
				else if (memcmp(rows + (size_t)i * 5, As, sizeof As) != 0)
					refuse("stencil coefficients vary along x: only y-dependent metrics are supported by the engine");
			}
		}
		if (mgb_set_level_operator(e, l, rows) != MGB_OK) die("mgb_set_level_operator");
		free(rows);
	}
	if (L > 1) {
		if (op->res[0].ni != 3 || op->res[0].nj != 3 || op->pro[0].ni != 3 || op->pro[0].nj != 3)
			refuse("transfer stencils must be 3x3");
		if (mgb_set_transfer(e, op->res[0].data, op->pro[0].data) != MGB_OK) die("mgb_set_transfer");
	}
	if (want_csr && map_style != 3)
		if (mgb_assemble_csr(e) != MGB_OK) die("mgb_assemble_csr");

	/* right-hand side on the finest level: b[i][j] = F(x_{j+1}, y_{i+1})  (ref: src/solver.c:594-597) */
	{
		const int ni = cfg.ni, nj = cfg.nj;
		double *b = malloc((size_t)ni * nj * sizeof(double));
		if (!b) refuse("out of host memory for the right-hand side");
#pragma omp parallel for schedule(static)
		for (int i = 0; i < ni; i++)
			for (int j = 0; j < nj; j++)
				b[(size_t)i * nj + j] = prob->Ffunc(mesh->coord[0][j + 1], mesh->coord[1][i + 1]);
		if (mgb_set_rhs(e, b) != MGB_OK) die("mgb_set_rhs");
		free(b);
	}
}

/* ---------------------------------------------------------------- options -> smoother */
/* what KSPSetFromOptions + PCSetFromOptions would configure for a KSP with the given options prefix */
static void read_smoother(const char *prefix, int rb_numbering, mgb_smoother *s, int *max_it, char *ksp_type, size_t klen)
{
	char key[128], pc[64] = "";
	s->type = -1; s->scale = 1.0; s->omega = 1.0; s->sor_sweep = MGB_SOR_SYMMETRIC; s->sor_its = 1;
	snprintf(key, sizeof key, "-%sksp_type", prefix);
	if (ksp_type) pbopt_get_string(key, ksp_type, klen);
	snprintf(key, sizeof key, "-%sksp_max_it", prefix);
	pbopt_get_int(key, max_it);
	snprintf(key, sizeof key, "-%sksp_richardson_scale", prefix);
	pbopt_get_real(key, &s->scale);
	snprintf(key, sizeof key, "-%spc_type", prefix);
	pbopt_get_string(key, pc, sizeof pc);
	if (!strcmp(pc, "jacobi")) s->type = MGB_SMOOTH_JACOBI;
	else if (!strcmp(pc, "ilu")) {
		if (rb_numbering) refuse("-pc_type ilu is offered on the natural numbering (-map 0,1,2) only");
		s->type = MGB_SMOOTH_ILU0;                     /* ILU(0), natural ordering: PETSc's defaults */
	}
	else if (!strcmp(pc, "sor")) {
		/* natural numbering: PETSc's lexicographic MatSOR as anti-diagonal wavefronts (same operations, same order);
		 * -map 3: the same MatSOR arithmetic on the red-black numbering (fully parallel) */
		s->type = rb_numbering ? MGB_SMOOTH_RBSOR : MGB_SMOOTH_LEXSOR;
		int its = 1, lits = 1;
		snprintf(key, sizeof key, "-%spc_sor_omega", prefix); pbopt_get_real(key, &s->omega);
		snprintf(key, sizeof key, "-%spc_sor_its", prefix); pbopt_get_int(key, &its);
		snprintf(key, sizeof key, "-%spc_sor_lits", prefix); pbopt_get_int(key, &lits);
		s->sor_its = its * lits;
		/* later flags win, in PCSetFromOptions_SOR's order */
		snprintf(key, sizeof key, "-%spc_sor_symmetric", prefix); if (pbopt_get_bool(key)) s->sor_sweep = MGB_SOR_SYMMETRIC;
		snprintf(key, sizeof key, "-%spc_sor_backward", prefix); if (pbopt_get_bool(key)) s->sor_sweep = MGB_SOR_BACKWARD;
		snprintf(key, sizeof key, "-%spc_sor_forward", prefix); if (pbopt_get_bool(key)) s->sor_sweep = MGB_SOR_FORWARD;
		snprintf(key, sizeof key, "-%spc_sor_local_symmetric", prefix); if (pbopt_get_bool(key)) s->sor_sweep = MGB_SOR_SYMMETRIC;
		snprintf(key, sizeof key, "-%spc_sor_local_backward", prefix); if (pbopt_get_bool(key)) s->sor_sweep = MGB_SOR_BACKWARD;
		snprintf(key, sizeof key, "-%spc_sor_local_forward", prefix); if (pbopt_get_bool(key)) s->sor_sweep = MGB_SOR_FORWARD;
	} else if (pc[0] == '\0') s->type = -1;
	else {
		snprintf(g_msg, sizeof g_msg, "-%spc_type %s is not available on the B200 engine (have: jacobi, sor, ilu)", prefix, pc);
		refuse(g_msg);
	}
}

/* ---------------------------------------------------------------- solve */
/* what the level KSPs of cycle 0 are configured with (ref: src/solver.c:1463-1510 + KSPSetFromOptions) */
static void vcycle_params(Solver *solver, mgb_vcycle_params *pp)
{
	int map_style = 2; pbopt_get_int("-map", &map_style);
	mgb_vcycle_params p;
	memset(&p, 0, sizeof p);
	char ksp_type[64] = "richardson";
	int max_it = -1;
	read_smoother("", map_style == 3, &p.smoother, &max_it, ksp_type, sizeof ksp_type);
	if (strcmp(ksp_type, "richardson"))
		refuse("cycle 0 smooths with KSPRICHARDSON (src/solver.c:1464); other -ksp_type values are not offered");
	if (p.smoother.type < 0) {
		/* no -pc_type: PETSc's default for a sequential AIJ matrix, ILU(0) -- what the shipped poisson.in runs with */
		if (map_style == 3) refuse("no -pc_type given: PETSc's default ILU(0) is offered on the natural numbering (-map 0,1,2) only");
		p.smoother.type = MGB_SMOOTH_ILU0;
	}
	/* -ksp_max_it, if present, overrides the sweep counts of every level KSP (KSPSetFromOptions comes last) */
	p.v0 = (max_it >= 0) ? max_it : solver->v[0];
	p.v1 = (max_it >= 0) ? max_it : solver->v[1];
	p.max_iter = solver->numIter;
	p.rtol = 1.e-7;
	pbopt_get_real("-rtol", &p.rtol);
	p.use_graph = 1;
	pbopt_get_int("-mgb_graph", &p.use_graph);
	{ int fuse = 1, bottom = 1; pbopt_get_int("-mgb_fuse", &fuse); pbopt_get_int("-mgb_bottom", &bottom); p.no_fuse = !fuse; p.no_bottom = !bottom; }
	*pp = p;
}

static void solve_vcycle(Solver *solver)
{
	mgb_engine *e = ENGINE(solver->assem);
	mgb_vcycle_params p;
	vcycle_params(solver, &p);
	int iters = 0; double seconds = 0.0;
	const clock_t c0 = clock();
	if (mgb_solve_vcycle(e, &p, solver->rnorm, &iters, &seconds) != MGB_OK) die("mgb_solve_vcycle");
	const clock_t c1 = clock();
	solver->numIter = iters;
	printf("rank = [%d]; Solver cputime:                %lf\n", 0, (double)(c1 - c0) / CLOCKS_PER_SEC);
	printf("rank = [%d]; Solver walltime:               %lf\n", 0, seconds);
}

static void solve_pcmg(Solver *solver)
{
	mgb_engine *e = ENGINE(solver->assem);
	int map_style = 2; pbopt_get_int("-map", &map_style);
	mgb_pcmg_params p;
	memset(&p, 0, sizeof p);
	char ktype[64] = "richardson", lk[64] = "chebyshev", ck[64] = "preonly", cpc[64] = "lu";
	pbopt_get_string("-ksp_type", ktype, sizeof ktype);
	if (!strcmp(ktype, "cg")) p.outer = MGB_KSP_CG;
	else if (!strcmp(ktype, "richardson")) p.outer = MGB_KSP_RICHARDSON;
	else refuse("cycle 8: -ksp_type must be richardson or cg on the B200 engine");
	p.rtol = 1.e-7; p.abstol = 1.e-50; p.dtol = 1.e4; p.max_iter = solver->numIter;       /* src/solver.c:1924 */
	pbopt_get_real("-ksp_rtol", &p.rtol);
	pbopt_get_real("-ksp_atol", &p.abstol);
	pbopt_get_real("-ksp_divtol", &p.dtol);
	pbopt_get_int("-ksp_max_it", &p.max_iter);
	if (p.max_iter > solver->numIter) p.max_iter = solver->numIter;   /* rnorm holds numIter+1 entries */
	double oscale = 1.0;
	if (pbopt_get_real("-ksp_richardson_scale", &oscale) && oscale != 1.0)
		refuse("cycle 8: outer -ksp_richardson_scale other than 1 is not offered");

	p.level_its = 2;                                                 /* PCMG default: 2 smoothing steps */
	read_smoother("mg_levels_", map_style == 3, &p.level_smoother, &p.level_its, lk, sizeof lk);
	if (strcmp(lk, "richardson") || p.level_smoother.type < 0)
		refuse("PCMG's default level smoother (Chebyshev with GMRES eigenvalue estimates + SOR) is not reproducible and not "
		       "offered: pass -mg_levels_ksp_type richardson -mg_levels_pc_type jacobi|sor [-mg_levels_ksp_richardson_scale w] "
		       "-mg_levels_ksp_max_it nu");
	pbopt_get_string("-mg_coarse_ksp_type", ck, sizeof ck);
	pbopt_get_string("-mg_coarse_pc_type", cpc, sizeof cpc);
	if (!strcmp(ck, "preonly") && !strcmp(cpc, "lu")) p.coarse = MGB_COARSE_LU;
	else if (!strcmp(ck, "richardson")) {
		p.coarse = MGB_COARSE_RICHARDSON;
		p.coarse_its = 1;                                            /* coarse KSP max_it defaults to 1 */
		read_smoother("mg_coarse_", map_style == 3, &p.coarse_smoother, &p.coarse_its, NULL, 0);
		if (p.coarse_smoother.type < 0) refuse("-mg_coarse_ksp_type richardson needs -mg_coarse_pc_type jacobi|sor");
	} else refuse("cycle 8: coarse solver must be preonly+lu (default) or richardson+jacobi|sor");

	{ int fuse = 1, bottom = 1, graph = 1; pbopt_get_int("-mgb_fuse", &fuse); pbopt_get_int("-mgb_bottom", &bottom); pbopt_get_int("-mgb_graph", &graph);
	  p.no_fuse = !fuse; p.no_bottom = !bottom; p.no_graph = !graph; }
	int iters = 0, reason = 0; double seconds = 0.0;
	if (mgb_solve_pcmg(e, &p, solver->rnorm, &iters, &reason, &seconds) != MGB_OK) die("mgb_solve_pcmg");
	solver->numIter = iters;
	printf("rank = [%d]; Solver walltime:               %lf\n", 0, seconds);
	printf("KSP converged reason: %d\n", reason);
}

/* B200 extension: Solve() for a stream of right-hand sides (cycle 0), host<->device copies overlapped with the solves */
int pb200_solve_many_impl(Solver *solver, int nrhs, const double *const *b, double *const *u, int *iters, double *finals, double *seconds)
{
	if (solver->cycle != VCYCLE) refuse("the pipelined multi-right-hand-side solve is offered for -cycle 0");
	mgb_vcycle_params p;
	vcycle_params(solver, &p);
	if (mgb_solve_vcycle_many(ENGINE(solver->assem), &p, nrhs, b, u, iters, finals, seconds) != MGB_OK) die("mgb_solve_vcycle_many");
	return 0;
}

void Solve(Solver *solver)
{
	if (solver->cycle == VCYCLE) solve_vcycle(solver);
	else if (solver->cycle == PetscPCMG) solve_pcmg(solver);
	else refuse("cycle not available on the B200 engine");
}

/* ---------------------------------------------------------------- post-processing */
void SetUpPostProcess(PostProcess *pp)
{
	pp->solData = fopen("uData.dat", "w");
	pp->resData = fopen("rData.dat", "w");
	pp->errData = fopen("eData.dat", "w");
	pp->XgridData = fopen("XgridData.dat", "w");
	pp->YgridData = fopen("YgridData.dat", "w");
}

void DestroyPostProcess(PostProcess *pp)
{
	FILE **f[5] = {&pp->solData, &pp->resData, &pp->errData, &pp->XgridData, &pp->YgridData};
	for (int k = 0; k < 5; k++) if (*f[k]) { fclose(*f[k]); *f[k] = NULL; }
}

/* error triple against the analytic solution, summed in row-major order like GetError (src/solver.c:1227-1236) */
static void error_norms(Problem *prob, Mesh *mesh, const double *u, int ni, int nj, double *error)
{
	error[0] = 0.0; error[1] = 0.0; error[2] = 0.0;
	for (int i = 0; i < ni; i++)
		for (int j = 0; j < nj; j++) {
			const double sol = prob->SOLfunc(mesh->coord[0][j + 1], mesh->coord[1][i + 1]);
			const double diff = fabs(u[(size_t)i * nj + j] - sol);
			error[0] = fmax(diff, error[0]);
			error[1] = error[1] + diff;
			error[2] = error[2] + diff * diff;
		}
	error[2] = sqrt(error[2]);
}

void Postprocessing(Problem *prob, Mesh *mesh, Indices *indices, Solver *solver, PostProcess *pp)
{
	const int ni = indices->level[0].grid[0].ni, nj = indices->level[0].grid[0].nj;
	double *u = malloc((size_t)ni * nj * sizeof(double));
	if (!u) refuse("out of host memory for the solution");
	if (mgb_get_solution(ENGINE(solver->assem), u) != MGB_OK) die("mgb_get_solution");       /* GetSol */
	error_norms(prob, mesh, u, ni, nj, pp->error);
	for (int k = 0; k < 3; k++) {
		printf("\nerror[%d] = %.16e\n", k, pp->error[k]);
		if (pp->errData) fprintf(pp->errData, "%.16e\n", pp->error[k]);
	}
	/* file formats of src/solver.c:1337-1353 (note: X/YgridData print coord[.][j], not [j+1], as the reference does) */
	if (pp->solData && pp->XgridData && pp->YgridData) {
		for (int i = 0; i < ni; i++) {
			for (int j = 0; j < nj; j++) {
				fprintf(pp->XgridData, "%lf    ", mesh->coord[0][j]);
				fprintf(pp->YgridData, "%lf    ", mesh->coord[1][i]);
				fprintf(pp->solData, "%.16e    ", u[(size_t)i * nj + j]);
			}
			fprintf(pp->XgridData, "\n"); fprintf(pp->YgridData, "\n"); fprintf(pp->solData, "\n");
		}
	}
	if (pp->resData) {
		for (int k = 0; k < solver->numIter + 1; k++) fprintf(pp->resData, "%.16e ", solver->rnorm[k]);
		fprintf(pp->resData, "\n");
	}
	printf("Relative residual = %.16e ", solver->rnorm[solver->numIter]);
	free(u);
}
