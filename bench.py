#!/usr/bin/env python
"""bench.py -- V-cycles/s and smoother HBM GB/s of the B200 multigrid engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[3] -- 2-D Poisson, 8193 x 8193 points (8191^2 unknowns), fp64,
13-level V(3,3) cycle of the reference's cycle 0 (src/solver.c:1530-1550) with the PETSc-comparable smoother
Richardson + PCJACOBI, scale 0.8.  One "step" = one V-cycle including the fine-level residual norm and its host
read-back (the reference's loop body).  At N > 1 the same grid is split into N row strips (strong scaling).

  value      V-cycles/s, data resident in HBM, K cycles timed with CUDA events on the engine's stream
  e2e        V-cycles/s through the host C layer (pb200_solve_rhs: the reference-facing Solve() with HOST buffers):
             every solve uploads the right-hand side from pinned host memory, iterates to the reference's
             tolerance (1e-7) and copies the solution back; cycles done / wall time of the calls
  roofline   the dominant kernel = the fused fine-level down leg (3 Jacobi sweeps + residual + restriction in one pass):
             26 B per unknown x 8191^2 / CUDA-event time per launch; fine_level_ops lists the one-sweep kernels too
  cpu_baseline  the reference's own sources (oracle/_ref/poisson_ref: src/*.c over the in-repo mini-PETSc, real
             PETSc is not in the image) on the host cores, on a bounded sample
  --impl reference   that CPU program on the full workload for a few cycles (rank 0 only)
"""
import argparse
import importlib
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NPTS = 8193
LEVELS = 13
SMOOTHER = "-pc_type jacobi -ksp_richardson_scale 0.8"
JACOBI_BYTES_PER_UNKNOWN = 24          # SURVEY.md 8d / DESIGN.md: read x, read b, write x
VCYCLE_BYTES_PER_FINE_UNKNOWN = 264    # unfused per-sweep byte count of a V(3,3) cycle, all levels
# the dominant kernel of the product path is the fused down leg k_jfused<3,PRE_GIVEN,POST_RESTRICT> on level 0
# (3 Jacobi sweeps + residual + restriction in one pass): it reads u and b once, writes u once and 1/4 coarse value
FUSED_DOWN_BYTES_PER_UNKNOWN = 26      # 8 + 8 + 8 + 2
FUSED_DOWN_UNFUSED_BYTES = 90          # the SURVEY.md 8d count of what it replaces: 3 x 24 (sweeps) + 18 (residual+restrict)
NCU_TRAFFIC_FUSED_DOWN = 1.780e9       # dram read + write bytes per launch, ncu --set full (profiles/r1_ncu_fused_down.txt)


def options(npts, levels, iters, extra=""):
    return (f"-npts {npts} -mesh 0 -iter {iters} -grids {levels} -levels {levels} -cycle 0 -map 2 -v 3,3 "
            f"-moreNorm 0 {SMOOTHER} {extra}").strip()


def options_file(opts):
    """option string -> poisson.in text (one '-key value' per line)"""
    lines, toks = [], opts.split()
    for t in toks:
        if t.startswith("-") and not re.fullmatch(r"-[0-9.].*", t):
            lines.append(t)
        else:
            lines[-1] += " " + t
    return "\n".join(lines) + "\n"


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU reference
def run_reference(npts, levels, cycles, threads=None):
    """The reference's own driver on the host cores: returns (V-cycles/s, cores, kind, walltime, cycles)."""
    from oracle import ref_binary_path
    ncores = threads or os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(ncores))
    exe = ref_binary_path()
    if os.path.exists(exe):
        with tempfile.TemporaryDirectory() as d:
            with open(os.path.join(d, "poisson.in"), "w") as f:
                f.write(options_file(options(npts, levels, cycles)))
            out = subprocess.run([exe], cwd=d, env=env, capture_output=True, text=True, timeout=3000)
        m = re.search(r"Solver walltime:\s+([0-9.eE+-]+)", out.stdout)
        it = re.search(r"Number of iterations:\s+(\d+)", out.stdout)
        if out.returncode != 0 or not m or not it:
            raise RuntimeError("reference run failed: " + out.stderr[-400:] + out.stdout[-400:])
        wall, done = float(m.group(1)), int(it.group(1))
        return done / wall, ncores, "reference", wall, done
    # no oracle/_ref on this machine: the oracle's restatement of the same loop (kind "port")
    from oracle import Oracle
    o = Oracle(options(npts, levels, cycles))
    t0 = time.perf_counter()
    done, _ = o.solve()
    wall = time.perf_counter() - t0
    o.close()
    return done / wall, ncores, "port", wall, done


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cycles = max(1, min(a.steps, 3))
    n = NPTS - 2
    v, cores, kind, wall, done = run_reference(NPTS, LEVELS, cycles)
    sample = (f"{done} V-cycles of the full {NPTS}^2 / {LEVELS}-level workload, cycle loop only as the reference times it "
              f"(src/solver.c:1526-1553); {kind}: reference src/*.c over the in-repo mini-PETSc (real PETSc absent), "
              f"OpenMP over {cores} host threads")
    line = {"impl": "reference", "metric": "V-cycles/sec (fp64, 8193^2 grid)", "value": v, "unit": "V-cycles/s",
            "n_gpus": a.gpus, "steps": done, "warmup": 0, "ms_per_step": 1e3 * wall / done, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"2D Poisson {NPTS}^2 fp64, {LEVELS}-level V(3,3), Richardson+Jacobi 0.8", "unknowns": n * n},
            "cpu_baseline": {"value": v, "unit": "V-cycles/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------- B200 arm
def b200_arm(a):
    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 engine has no CPU path (use --impl reference for the CPU arm)")
    if a.gpus != world:
        raise SystemExit(f"bench.py: --gpus {a.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run --nproc-per-node {a.gpus}")
    torch.cuda.set_device(local)
    mgb = importlib.import_module("multigrid-petsc_b200")
    if a.workload == "weak":
        strips = importlib.import_module("multigrid-petsc_b200.strips")
        return strips.bench_weak(a, ClockSampler, hbm_peak)
    if world > 1:
        from importlib import import_module
        strips = import_module("multigrid-petsc_b200.strips")
        return strips.bench_strips(a, NPTS, LEVELS, ClockSampler, hbm_peak, options)

    n = NPTS - 2
    steps, warm = a.steps, max(a.warmup, 3)
    s = mgb.Session(options(NPTS, LEVELS, 1000, "-mgb_csr 0"))
    e = s.engine
    sm = mgb.jacobi(0.8)
    # ---- value: K cycles, data resident
    e.solve_vcycle(sm, 3, 3, max_iter=warm, rtol=0.0)
    l0 = e.launch_count()
    clk = ClockSampler(local)
    clk.start()
    torch.cuda.synchronize()
    it, rn, _ = e.solve_vcycle(sm, 3, 3, max_iter=steps, rtol=0.0)
    torch.cuda.synchronize()
    ms = e.last_solve_ms()
    launches = e.launch_count() - l0
    # keep the GPU busy a little longer so that the clock sampler sees the loaded state
    t_end = time.time() + (0.0 if a.profile else 1.0)
    while time.time() < t_end:
        e.solve_vcycle(sm, 3, 3, max_iter=steps, rtol=0.0)
    clocks = clk.stop()
    assert it == steps, (it, steps)
    value = steps / (ms * 1e-3)
    peak, peak_kind = hbm_peak()
    # ---- roofline of the dominant kernel (fused down leg on the fine level), CUDA events on the engine's stream
    t_f = e.time_op("fused_down", 0, 20)
    ach = FUSED_DOWN_BYTES_PER_UNKNOWN * n * n / (t_f * 1e-3) / 1e9
    per_op = {}
    for op in ("fused_down", "fused_up", "jacobi", "residual", "residual_norm", "residual_restrict", "prolong_correct"):
        t = t_f if op == "fused_down" else e.time_op(op, 0, 20)
        per_op[op] = {"us": t * 1e3, "gbs_unfused_count": mgb.OPS[op][1] * n * n / (t * 1e-3) / 1e9}
        if op in mgb.FUSED_OWN_BYTES:
            per_op[op]["gbs_own_traffic"] = mgb.FUSED_OWN_BYTES[op] * n * n / (t * 1e-3) / 1e9
    # ---- e2e: reference-facing Solve() with host buffers
    b_host = torch.empty(n * n, dtype=torch.float64, pin_memory=True)
    u_host = torch.empty(n * n, dtype=torch.float64, pin_memory=True)
    x = np.linspace(0.0, 1.0, NPTS)[1:-1]
    rng = np.random.default_rng(0)
    b_host.numpy().reshape(n, n)[:] = np.outer(np.sin(np.pi * x), -2 * np.pi ** 2 * np.sin(np.pi * x))
    b_host.numpy()[:] += 1e-3 * rng.standard_normal(n * n)        # synthetic right-hand side, seed 0
    b2_host = torch.empty(n * n, dtype=torch.float64, pin_memory=True)
    u2_host = torch.empty(n * n, dtype=torch.float64, pin_memory=True)
    b2_host.copy_(b_host)
    b2_host.numpy()[:] += 1e-3 * rng.standard_normal(n * n)       # a second, different right-hand side
    r = s.solve_rhs(b_host.data_ptr(), u_host.data_ptr())         # warm-up solve
    # (a) one Solve() at a time: upload, solve, download in sequence
    nsolve, cycles1 = (1 if a.profile else 3), 0
    t0 = time.perf_counter()
    for _ in range(nsolve):
        r = s.solve_rhs(b_host.data_ptr(), u_host.data_ptr())
        cycles1 += r["num_iter"]
    t_single = time.perf_counter() - t0
    # (b) a stream of right-hand sides through the same Solve(): the copies of neighbouring solves overlap the running one
    nstream = 2 if a.profile else 8
    bp = [b_host.data_ptr(), b2_host.data_ptr()] * (nstream // 2)
    up = [u_host.data_ptr(), u2_host.data_ptr()] * (nstream // 2)
    s.solve_rhs_many(bp[:2], up[:2])                              # warm-up (graphs for both pointer states)
    t0 = time.perf_counter()
    its, fin, _ = s.solve_rhs_many(bp, up)
    t_e2e = time.perf_counter() - t0
    cycles = sum(its)
    e2e = cycles / t_e2e
    bytes_per_solve = 8 * n * n
    s.close()
    # ---- CPU baseline on a bounded sample
    cpu = None
    if not a.no_cpu_baseline and not a.profile:
        sn, sl, sc = 4097, 12, 3
        v, cores, kind, wall, done = run_reference(sn, sl, sc)
        scale = ((sn - 2) ** 2) / float(n * n)
        cpu = {"value": v * scale, "unit": "V-cycles/s", "cores": cores, "kind": kind,
               "sample": (f"{done} V-cycles at {sn}^2 / {sl} levels ({wall:.2f} s of cycle loop; {v:.3f} cycles/s there), scaled by the "
                          f"unknown ratio {scale:.4f} to {NPTS}^2; {kind} = reference src/*.c over the in-repo mini-PETSc, "
                          f"OpenMP on {cores} host threads; `--impl reference` runs the full size")}
    line = {"metric": "V-cycles/sec (fp64, 8193^2 grid)", "value": value, "unit": "V-cycles/s", "n_gpus": 1, "steps": steps,
            "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"2D Poisson {NPTS}^2 fp64, {LEVELS}-level V(3,3), Richardson+Jacobi 0.8 (BASELINE configs[3] at N=1)",
                       "unknowns": n * n, "l2": "inputs larger than L2 (537 MB per fine vector)", "parallelism": "1 strip"},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_jfused<3,PRE_GIVEN,POST_RESTRICT> level 0 (3 Jacobi sweeps + residual + restriction, one pass)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_kind": peak_kind,
                         "traffic": NCU_TRAFFIC_FUSED_DOWN, "algorithmic_bytes_per_launch": FUSED_DOWN_BYTES_PER_UNKNOWN * n * n,
                         "note": "temporal blocking: the kernel's own compulsory traffic is 26 B/unknown; by the unfused SURVEY 8d count "
                                 "(90 B/unknown for the 3 sweeps + residual + restriction it replaces) it delivers "
                                 f"{FUSED_DOWN_UNFUSED_BYTES * n * n / (t_f * 1e-3) / 1e9:.0f} GB/s-equivalent; it is fp64-issue / shared-memory bound, "
                                 "not HBM bound (DESIGN.md section 4). The one-sweep kernels it replaces are listed in fine_level_ops.",
                         "vcycle_gbs_unfused_count": VCYCLE_BYTES_PER_FINE_UNKNOWN * n * n * value / 1e9,
                         "fine_level_ops": per_op},
            "e2e": {"value": e2e, "unit": "V-cycles/s", "h2d_bytes_per_step": bytes_per_solve * nstream / cycles,
                    "d2h_bytes_per_step": bytes_per_solve * nstream / cycles, "solves": nstream, "cycles_per_solve": cycles / nstream,
                    "single_solve_value": cycles1 / t_single,
                    "note": "a stream of right-hand sides through the host C layer (pb200_solve_rhs_many): per solve H2D rhs from "
                            "pinned memory + V-cycles to 1e-7 + D2H solution, the copies of neighbouring solves overlapping the running "
                            "one; wall clock over the whole batch incl. the un-overlapped first upload and last download; bytes are per "
                            "V-cycle. single_solve_value = the same without overlap (one pb200_solve_rhs at a time)"},
            "gpu_launches": launches,
            "final_relative_residual": float(rn[-1])}
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="strong", choices=["strong", "weak"],
                    help="strong (default): 8193^2 split over the GPUs (BASELINE configs[3]); weak: 4096 x 4097 points per GPU (configs[4])")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: no busy loop, one e2e solve, no CPU baseline")
    a = ap.parse_args()
    if a.impl == "reference":
        return reference_arm(a)
    return b200_arm(a)


if __name__ == "__main__":
    sys.exit(main())
