#!/usr/bin/env python
"""Multi-process strip check (run under torchrun on >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/strips_check.py
Solves the 1025^2 / 10-level golden case on WORLD_SIZE row strips (one process per GPU, P2P ghost rows) and compares
the gathered solution with the committed single-strip SHA-256 and iteration count."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
strips = importlib.import_module("multigrid-petsc_b200.strips")

if __name__ == "__main__":
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["n1025_l10_jacobi"]
    sys.exit(strips.selfcheck(1025, 10, g["u_sha256"], g["num_iter"]))
