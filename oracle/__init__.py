"""TEST INFRASTRUCTURE ONLY: the CPU oracle (see oracle/mg_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (multigrid-petsc_b200/) never does.
"""
from .binding import Oracle, build_oracle, oracle_lib_path, ref_binary_path, ref_l2_path  # noqa: F401
