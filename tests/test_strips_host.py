"""CPU tests of the host side of the row-strip path (no GPU, no compute calls): the partition arithmetic exported by
the C-ABI library (mgb_strip_rows), that the library loads and exports every symbol include/mgb200.h declares, and the
world_size-2 handle exchange over gloo."""
import importlib
import os
import re
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mgb = importlib.import_module("multigrid-petsc_b200")


@pytest.fixture(scope="module", autouse=True)
def _built():
    mgb.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    names = sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 30
    lib = mgb.engine_lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    host = mgb.host_lib()
    for n in ("SetUpSolver", "Assemble", "Solve", "DestroySolver", "SetUpPostProcess", "Postprocessing", "DestroyPostProcess",
              "SetUpIndices", "DestroyIndices", "mapping", "SetUpOperator", "DestroyOperator", "GridTransferOperators",
              "SetUpMesh", "DestroyMesh", "SetUpProblem", "pb200_run", "pb200_open", "pb200_solve", "pb200_solve_rhs", "pb200_close"):
        assert hasattr(host, n), n


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mgb.MgbError, match="no CPU fallback|no CUDA device"):
        mgb.Engine(2, 15)
    with pytest.raises(mgb.MgbError):
        mgb.run_poisson("-npts 17 -iter 5 -grids 2 -levels 2 -pc_type jacobi")


@pytest.mark.parametrize("n,levels,P,aggl", [(8191, 13, 2, 0), (8191, 13, 4, 0), (8191, 13, 8, 0), (4095, 12, 8, 0),
                                             (1023, 10, 3, 0), (127, 4, 2, 31), (127, 4, 4, 31), (99, 3, 2, 24)])
def test_strip_partition_properties(n, levels, P, aggl):
    first_whole = None
    for l in range(levels):
        nl = (n + 1) // (1 << l) - 1
        rows = [mgb.strip_rows(levels, n, n, P, l, r, aggl) for r in range(P)]
        dist = rows[0][2]
        assert all(d == dist for _, _, d in rows)
        if dist:
            assert first_whole is None                      # distributed levels are the finest ones
            assert rows[0][0] == 0 and rows[-1][1] == nl    # cover the grid ...
            for r in range(P - 1):
                assert rows[r][1] == rows[r + 1][0]         # ... contiguously, whole rows
            assert min(b - a for a, b, _ in rows) >= 12
            if l > 0:
                # the fine strip starts on the even fine row 2*c, right above the first coarse row of the same rank
                assert [a for a, _, _ in prev] == [2 * a for a, _, _ in rows]
        else:
            if first_whole is None:
                first_whole = l
                assert rows[0][0] == 0 and rows[-1][1] == nl and all(rows[r][1] == rows[r + 1][0] for r in range(P - 1))
                if l > 0:
                    assert [a for a, _, _ in prev] == [2 * a for a, _, _ in rows]
            else:
                assert all((a, b) == (0, nl) for a, b, _ in rows)
        prev = rows
    assert first_whole is not None and first_whole > 0


def test_strip_partition_errors():
    with pytest.raises(mgb.MgbError):
        mgb.strip_rows(4, 127, 127, 8, 0, 0, 31)       # strips thinner than twice the ghost depth
    with pytest.raises(mgb.MgbError):
        mgb.strip_rows(4, 127, 127, 2, 0, 0, 127)      # nothing to distribute
    with pytest.raises(mgb.MgbError):
        mgb.strip_rows(4, 127, 127, 2, 0, 5, 31)       # rank out of range


def test_handle_exchange_world_size_2_gloo(tmp_path):
    """Two CPU processes over gloo: every rank ends up with all handles concatenated in rank order."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import importlib, os, sys
        sys.path.insert(0, {ROOT!r})
        strips = importlib.import_module("multigrid-petsc_b200.strips")
        rank, world = strips.init_distributed("gloo")
        mine = bytes([rank + 1]) * 64
        allh = strips.exchange_handles(mine)
        assert len(allh) == 64 * world and allh[:64] == bytes([1]) * 64 and allh[64:128] == bytes([2]) * 64
        opts = strips.strip_options("-npts 8193 -levels 13", rank, world)
        assert f"-mgb_ranks {{world}} -mgb_rank {{rank}}" in opts and "-mgb_csr 0" in opts
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()
        print("ok" + str(rank), flush=True)
    """))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under multigrid-petsc_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "multigrid-petsc_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        if os.sep + "lib" in d or "__pycache__" in d:
            continue
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"\boracle\b|mgoracle|minipetsc|mg_oracle", txt):
                    # mentions in comments that point at the oracle as the checker are fine; imports / includes / links are not
                    for line in txt.splitlines():
                        if re.search(r"^\s*(from|import)\s+oracle|#include\s*[\"<].*(mg_oracle|minipetsc|petscksp)|-lmgoracle|-[IL]\S*oracle|import_module\(.oracle", line):
                            offenders.append((f, line.strip()))
    assert not offenders, offenders
