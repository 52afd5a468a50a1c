"""Host logic of the PETSc surface (multigrid-petsc_b200/host/petsc_b200/petsc_b200.c) that needs no GPU: the options
database (file first, argv wins, '#' comments, prefixes, negative numbers are values not keys, integer arrays, flags), the
1-rank MPI stubs, and the loud failure of the first device object without a CUDA device.  A small C program is compiled
against the product header exactly as the reference's sources are (-Ipetsc_b200, #include <petscksp.h>)."""
import importlib
import os
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PB = os.path.join(ROOT, "multigrid-petsc_b200", "host", "petsc_b200")
LIB = os.path.join(ROOT, "multigrid-petsc_b200", "lib")

PROG = r"""
#include <petscksp.h>
#include <assert.h>
int main(int argc, char **argv)
{
	PetscInt n = -1, v[4] = {0, 0, 0, 0}, nv = 4, m = 7; PetscReal w = 0.0, neg = 0.0; char s[32] = ""; PetscBool set, has;
	int rank = 5, size = 5;
	PetscInitialize(&argc, &argv, "poisson.in", 0);
	MPI_Comm_rank(PETSC_COMM_WORLD, &rank); MPI_Comm_size(PETSC_COMM_WORLD, &size);
	assert(rank == 0 && size == 1 && MPI_Wtime() > 0.0);
	PetscOptionsGetInt(NULL, NULL, "-npts", &n, &set);            assert(set && n == 65);        /* argv overrides the file's 17 */
	PetscOptionsGetIntArray(NULL, NULL, "-v", v, &nv, &set);      assert(set && nv == 2 && v[0] == 3 && v[1] == 2);
	PetscOptionsGetInt(NULL, NULL, "-absent", &m, &set);          assert(!set && m == 7);       /* untouched */
	PetscOptionsGetReal(NULL, "mg_levels_", "-ksp_richardson_scale", &w, &set);  assert(set && w == 0.8);
	PetscOptionsGetReal(NULL, NULL, "-shift", &neg, &set);        assert(set && neg == -1.5);   /* "-1.5" is a value, not a key */
	PetscOptionsGetString(NULL, NULL, "-pc_type", s, sizeof s, &set);            assert(set && !strcmp(s, "jacobi"));
	PetscOptionsHasName(NULL, NULL, "-pc_sor_forward", &has);     assert(has);                  /* a flag without a value */
	PetscOptionsHasName(NULL, NULL, "-commented_out", &has);      assert(!has);
	PetscOptionsInsertString(NULL, "-npts 9 -late 4");
	PetscOptionsGetInt(NULL, NULL, "-npts", &n, &set);            assert(n == 9);               /* later insertions win */
	PetscPrintf(PETSC_COMM_WORLD, "options ok %d\n", n);
	if (argc > 3) {                                                /* the first device object: loud without a GPU */
		Vec x; VecCreateSeq(PETSC_COMM_SELF, 10, &x);
		PetscPrintf(PETSC_COMM_WORLD, "vector created\n");
	}
	PetscFinalize();
	return 0;
}
"""


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_options_database_mpi_stubs_and_loud_failure(tmp_path):
    importlib.import_module("multigrid-petsc_b200").build()
    (tmp_path / "t.c").write_text(PROG)
    (tmp_path / "poisson.in").write_text(textwrap.dedent("""\
        # a comment line
        -npts 17
        -v 3,2
        -pc_type jacobi          # trailing comment
        -mg_levels_ksp_richardson_scale 0.8
        -shift -1.5
        -pc_sor_forward
        # -commented_out 1
        """))
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=gnu99", "-O1", "-I" + PB, "-I" + os.path.join(ROOT, "include"), "-o", str(exe), str(tmp_path / "t.c"),
                    os.path.join(PB, "petsc_b200.c"), "-L" + LIB, "-lmgb200", "-Wl,-rpath," + LIB, "-lm"], check=True)
    out = subprocess.run([str(exe), "-npts", "65"], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "options ok 9" in out.stdout
    out = subprocess.run([str(exe), "-npts", "65", "-device"], cwd=tmp_path, capture_output=True, text=True)
    if _has_gpu():
        assert out.returncode == 0 and "vector created" in out.stdout
    else:
        assert out.returncode != 0 and "vector created" not in out.stdout
        assert "petsc_b200 error" in out.stderr and "no CUDA device" in out.stderr and "no CPU fallback" in out.stderr
