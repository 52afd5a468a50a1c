"""CPU tests of bench.py's host-side pieces: the poisson.in writer and the reference arm (the reference's own sources
over mini-PETSc, or the oracle port when oracle/_ref is absent) on a small grid."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_options_file_round_trip():
    txt = bench.options_file(bench.options(129, 4, 7, "-mgb_csr 0"))
    lines = txt.strip().splitlines()
    assert "-npts 129" in lines and "-iter 7" in lines and "-v 3,3" in lines
    assert "-ksp_richardson_scale 0.8" in lines and "-pc_type jacobi" in lines and "-mgb_csr 0" in lines
    assert all(l.startswith("-") for l in lines)


def test_reference_arm_small_grid():
    v, cores, kind, wall, done = bench.run_reference(129, 4, 3, threads=2)
    assert done == 3 and v > 0 and wall > 0 and cores == 2
    assert kind in ("reference", "port")
