/*
 * poisson_b200.c -- the driver: the reference's main() (ref: src/poisson.c:27-138) with the same option keys,
 * the same guards and the same call order, on top of the B200 solver layer.  Also exported as pb200_run() so
 * that tests and benchmarks can run the whole pipeline in-process from an option string.
 *
 * Unlike the reference (which reads its keys into uninitialised variables, src/poisson.c:51-59), missing
 * required keys (-npts, -iter, -levels) are an error here; the optional ones default to the shipped
 * poisson.in values (-mesh 0 -cycle 0 -map 2 -v 3,3 -moreNorm 0, -grids = -levels).
 */
#include "pb_api.h"
#include "mgb200.h"
#include <setjmp.h>
#include <string.h>
#include <unistd.h>

extern __thread jmp_buf *pb200_trap;
long long mgb_launch_count(const struct mgb_engine *e);
double mgb_last_solve_ms(const struct mgb_engine *e);

/* ref: src/poisson.c:165-214 */
static void print_info(Mesh *mesh, Indices *indices, Operator *op, Solver *solver, int cyc, int meshflag, int mapflag)
{
	printf("=============================================================\n");
	printf("Size:				%d x %d\n", mesh->n[0], mesh->n[1]);
	if (meshflag == 0) printf("Mesh Type:			Uniform\n");
	if (meshflag == 1 || meshflag == 2) printf("Mesh Type:			Non Uniform\n");
	printf("Number of grids:		%d\n", op->totalGrids);
	printf("Number of levels:		%d\n", solver->assem->levels);
	printf("Number of grids per level:	");
	for (int l = 0; l < indices->levels; l++) printf("%d	", indices->level[l].grids);
	printf("\n");
	printf("Number of unknowns per level:	");
	for (int l = 0; l < indices->levels; l++) printf("%d	", indices->level[l].global.ni);
	printf("\n");
	if (mapflag == 0) printf("Mapping style :			Grid after grid\n");
	if (mapflag == 1) printf("Mapping style :			Through the grids\n");
	if (mapflag == 2) printf("Mapping style :			Local grid after grid\n");
	if (mapflag == 3) printf("Mapping style :			Red-black (B200 extension)\n");
	if (cyc == 8) printf("Cycle :				Petsc-V-Cycle\n");
	if (cyc == 0) printf("Cycle :				V-Cycle\n");
	printf("Number of smoothing steps :	%d(fine) %d(coarsest)\n", solver->v[0], solver->v[1]);
	printf("Number of processes:		%d\n", 1);
	printf("Number of iterations:		%d\n", solver->numIter);
	printf("=============================================================\n");
}

/* One driver run split in the reference's phases, so that callers can stop between Assemble and Solve. */
typedef struct pb200_session {
	Problem prob; Mesh mesh; Indices indices; Operator op; Solver solver; PostProcess pp;
	int cyc, meshflag, mapflag;
	int assembled, solved;
	int max_iter;                       /* -iter as given (Solve overwrites solver.numIter with the count done) */
} pb200_session;

/* SetUpProblem .. Assemble (ref: src/poisson.c:45-118) from the options database */
static int session_setup(pb200_session *s)
{
	int vmax = 2;
	memset(s, 0, sizeof *s);
	s->mapflag = 2;
	SetUpProblem(&s->prob);
	s->solver.v[0] = 3; s->solver.v[1] = 3; s->solver.moreInfo = 0;
	const int has_npts = pbopt_get_int("-npts", s->mesh.n);
	pbopt_get_int("-mesh", &s->meshflag);
	const int has_iter = pbopt_get_int("-iter", &s->solver.numIter);
	const int has_levels = pbopt_get_int("-levels", &s->indices.levels);
	if (!pbopt_get_int("-grids", &s->indices.totalGrids)) s->indices.totalGrids = s->indices.levels;
	pbopt_get_int("-cycle", &s->cyc);
	pbopt_get_int("-map", &s->mapflag);
	pbopt_get_int_array("-v", s->solver.v, &vmax);
	pbopt_get_int("-moreNorm", &s->solver.moreInfo);
	if (!has_npts || !has_iter || !has_levels) {
		fprintf(stderr, "poisson_b200 ERROR: options -npts, -iter and -levels are required\n");
		return 2;
	}
	if (s->mesh.n[0] < 3 || s->indices.levels < 1 || s->solver.numIter < 0 || s->meshflag < 0 || s->meshflag > 2 ||
	    s->mapflag < 0 || s->mapflag > 3) {
		fprintf(stderr, "poisson_b200 ERROR: invalid option value\n");
		return 2;
	}
	if (s->cyc != 0 && s->cyc != 8) {
		fprintf(stderr, "poisson_b200 ERROR: only -cycle 0 (V-cycle) and -cycle 8 (PCMG) run on the B200 engine\n");
		return 2;
	}
	if (s->indices.totalGrids != s->indices.levels) {
		fprintf(stderr, "poisson_b200 ERROR: one grid per level only (-grids must equal -levels)\n");
		return 2;
	}
	if (((s->mesh.n[0] - 1) >> (s->indices.levels - 1)) < 2) {
		fprintf(stderr, "poisson_b200 ERROR: too many levels for -npts %d\n", s->mesh.n[0]);
		return 2;
	}
	/* square grid on the unit square (ref: src/poisson.c:73-82) */
	for (int i = 1; i < DIMENSION; i++) s->mesh.n[i] = s->mesh.n[0];
	for (int i = 0; i < DIMENSION; i++) { s->mesh.bounds[i * 2] = 0.0; s->mesh.bounds[i * 2 + 1] = 1.0; }

	SetUpMesh(&s->mesh, s->meshflag == 0 ? UNIFORM : (s->meshflag == 1 ? NONUNIFORM1 : NONUNIFORM2));
	s->indices.coarseningFactor = 2;
	SetUpIndices(&s->mesh, &s->indices);
	mapping(&s->indices, s->mapflag);
	SetUpOperator(&s->indices, &s->op);
	GridTransferOperators(s->op, s->indices);
	SetUpSolver(&s->indices, &s->solver, s->cyc == 0 ? VCYCLE : PetscPCMG);
	Assemble(&s->prob, &s->mesh, &s->indices, &s->op, &s->solver);
	s->assembled = 1;
	s->max_iter = s->solver.numIter;
	return 0;
}

/* Solve .. PrintInfo (ref: src/poisson.c:123-128) */
static int session_solve(pb200_session *s, const char *dir, int verbose)
{
	Solve(&s->solver);
	s->solved = 1;
	char cwd[1024] = "";
	if (dir) {
		if (!getcwd(cwd, sizeof cwd) || chdir(dir) != 0) { fprintf(stderr, "poisson_b200 ERROR: cannot enter %s\n", dir); return 2; }
		SetUpPostProcess(&s->pp);
	}
	Postprocessing(&s->prob, &s->mesh, &s->indices, &s->solver, &s->pp);
	printf("\n");
	if (verbose) print_info(&s->mesh, &s->indices, &s->op, &s->solver, s->cyc, s->meshflag, s->mapflag);
	if (dir) { DestroyPostProcess(&s->pp); if (chdir(cwd) != 0) return 2; }
	return 0;
}

/* Destroy* (ref: src/poisson.c:130-134) */
static void session_teardown(pb200_session *s)
{
	if (s->solver.assem) DestroySolver(&s->solver);
	if (s->op.res) DestroyOperator(&s->op);
	if (s->indices.level) DestroyIndices(&s->indices);
	DestroyMesh(&s->mesh);
}

static void session_results(pb200_session *s, pb200_result *res, double *u_out, double *rnorm_out, int rnorm_cap)
{
	if (res) {
		res->num_iter = s->solver.numIter;
		res->ni = s->indices.level[0].grid[0].ni; res->nj = s->indices.level[0].grid[0].nj;
		memcpy(res->error, s->pp.error, sizeof s->pp.error);
		res->levels = s->indices.levels;
		res->gpu_launches = mgb_launch_count(pb200_engine(&s->solver));
		res->solve_seconds = 1e-3 * mgb_last_solve_ms(pb200_engine(&s->solver));   /* cycle loop only, CUDA events */
	}
	if (rnorm_out) for (int k = 0; k <= s->solver.numIter && k < rnorm_cap; k++) rnorm_out[k] = s->solver.rnorm[k];
	if (u_out) mgb_get_solution(pb200_engine(&s->solver), u_out);
}

/* stdout of the library entry points goes to /dev/null unless -pb_verbose is given */
static int quiet_begin(void)
{
	fflush(stdout);
	int saved = dup(1);
	if (!pbopt_get_bool("-pb_verbose")) {
		FILE *dn = fopen("/dev/null", "w");
		if (dn) { dup2(fileno(dn), 1); fclose(dn); }
	}
	return saved;
}
static void quiet_end(int saved)
{
	fflush(stdout);
	if (saved >= 0) { dup2(saved, 1); close(saved); }
}

int pb200_open(const char *options, pb200_session **out)
{
	jmp_buf trap;
	pbopt_clear();
	pbopt_insert_string(options);
	pb200_session *s = malloc(sizeof *s);
	const int saved = quiet_begin();
	int rc;
	pb200_trap = &trap;
	if (setjmp(trap) == 0) rc = session_setup(s);
	else rc = 1;
	pb200_trap = NULL;
	quiet_end(saved);
	if (rc != 0) { session_teardown(s); free(s); s = NULL; }
	*out = s;
	return rc;
}

struct mgb_engine *pb200_session_engine(pb200_session *s) { return s ? pb200_engine(&s->solver) : NULL; }

int pb200_solve(pb200_session *s, const char *dir, pb200_result *res, double *u, double *rnorm, int rnorm_cap)
{
	jmp_buf trap;
	const int saved = quiet_begin();
	int rc;
	pb200_trap = &trap;
	if (setjmp(trap) == 0) rc = session_solve(s, dir, 1);
	else rc = 1;
	pb200_trap = NULL;
	quiet_end(saved);
	if (rc == 0) session_results(s, res, u, rnorm, rnorm_cap);
	return rc;
}

/* One more solve on an assembled session with a caller-supplied right-hand side (host, ni*nj doubles, natural
 * order -- what levelvecb would have put into b[0], ref: src/solver.c:594-597) and the solution copied back to a
 * host buffer (what GetSol does, ref: src/solver.c:1255-1313).  No files, no error norms. */
int pb200_solve_rhs(pb200_session *s, const double *b, double *u, pb200_result *res, double *rnorm, int rnorm_cap)
{
	jmp_buf trap;
	if (!s || !b) return 2;
	const int saved = quiet_begin();
	int rc = 0;
	pb200_trap = &trap;
	if (setjmp(trap) == 0) {
		s->solver.numIter = s->max_iter;
		if (mgb_set_rhs(pb200_engine(&s->solver), b) != MGB_OK) rc = 1;
		else Solve(&s->solver);
	} else rc = 1;
	pb200_trap = NULL;
	quiet_end(saved);
	if (rc == 0) {
		memset(s->pp.error, 0, sizeof s->pp.error);
		session_results(s, res, u, rnorm, rnorm_cap);
	}
	return rc;
}

/* the same for a stream of right-hand sides: uploads, solves and downloads are pipelined (mgb_solve_vcycle_many) */
int pb200_solve_many_impl(Solver *solver, int nrhs, const double *const *b, double *const *u, int *iters, double *finals, double *seconds);
int pb200_solve_rhs_many(pb200_session *s, int nrhs, const double *const *b, double *const *u, int *iters, double *finals, double *seconds)
{
	jmp_buf trap;
	if (!s || !b || !u || nrhs < 1) return 2;
	const int saved = quiet_begin();
	int rc = 0;
	pb200_trap = &trap;
	if (setjmp(trap) == 0) {
		s->solver.numIter = s->max_iter;
		rc = pb200_solve_many_impl(&s->solver, nrhs, b, u, iters, finals, seconds);
	} else rc = 1;
	pb200_trap = NULL;
	quiet_end(saved);
	return rc;
}

void pb200_close(pb200_session *s)
{
	if (!s) return;
	session_teardown(s);
	free(s);
}

int pb200_run(const char *options, const char *dir, pb200_result *res, double *u, double *rnorm, int rnorm_cap)
{
	pb200_session *s = NULL;
	int rc = pb200_open(options, &s);
	if (rc == 0) rc = pb200_solve(s, dir, res, u, rnorm, rnorm_cap);
	pb200_close(s);
	return rc;
}

#ifndef PB200_NO_MAIN
int main(int argc, char *argv[])
{
	pbopt_clear();
	pbopt_insert_file("poisson.in");          /* PetscInitialize(&argc, &argv, "poisson.in", 0) */
	pbopt_insert_args(argc, argv);
	pb200_session s;
	int rc = session_setup(&s);
	if (rc == 0) rc = session_solve(&s, ".", 1);
	session_teardown(&s);
	return rc;
}
#endif
