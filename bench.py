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
NCU_FUSED_DOWN_SUMMARY = os.path.join(ROOT, "profiles", "r2_ncu_fused_down.txt")   # ncu --set full summary of that kernel
WORKLOAD = f"2D Poisson {NPTS}^2 fp64, {LEVELS}-level V(3,3), Richardson+Jacobi 0.8 (BASELINE configs[3])"
METRIC = "V-cycles/sec (fp64, 8193^2 grid)"
PARITY_TOL = 1e-10                     # north_star: per-cycle residual norms within 1e-10 relative
MGJ = ("-mg_levels_ksp_type richardson -mg_levels_pc_type jacobi -mg_levels_ksp_richardson_scale 0.8 "
       "-mg_levels_ksp_max_it 3")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu summary"""
    try:
        for line in open(NCU_FUSED_DOWN_SUMMARY):
            m = re.match(r"\s*traffic = dram read \+ write\s+([0-9.eE+]+)\s+byte", line)
            if m:
                return float(m.group(1)), os.path.relpath(NCU_FUSED_DOWN_SUMMARY, ROOT)
    except OSError:
        pass
    return None, None


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
def _record_paths(tag):
    return [os.path.join(ROOT, "gpurun_out", f"ref_record_{tag}.json"), os.path.join(tempfile.gettempdir(), f"mgb200_ref_record_{tag}.json")]


def save_record(tag, rec):
    for path in _record_paths(tag):
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            json.dump(rec, open(path, "w"))
        except OSError:
            pass


def load_record(tag, max_age_s=6 * 3600):
    """what a reference run on THIS box left behind (bench.py --impl reference runs right before the B200 arm)"""
    for path in _record_paths(tag):
        try:
            if time.time() - os.path.getmtime(path) <= max_age_s:
                rec = json.load(open(path))
                if rec.get("host") == os.uname().nodename:
                    return rec
        except (OSError, ValueError):
            pass
    return None


def run_reference_opts(opts, threads=None):
    """The reference's own driver (oracle/_ref/poisson_ref: unmodified src/*.c over the in-repo mini-PETSc) with the
    given poisson.in options on the host cores.  Returns dict(iters, wall, rnorm[], error[3], cores, kind)."""
    from oracle import ref_binary_path
    ncores = threads or os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(ncores))
    exe = ref_binary_path()
    if os.path.exists(exe):
        with tempfile.TemporaryDirectory() as d:
            with open(os.path.join(d, "poisson.in"), "w") as f:
                f.write(options_file(opts))
            # the solution / grid dumps (~27 bytes per unknown each, src/solver.c:1337-1346) are not needed here: they go to /dev/null
            for name in ("uData.dat", "XgridData.dat", "YgridData.dat"):
                os.symlink(os.devnull, os.path.join(d, name))
            out = subprocess.run([exe], cwd=d, env=env, capture_output=True, text=True, timeout=3000)
            m = re.search(r"Solver walltime:\s+([0-9.eE+-]+)", out.stdout)
            it = re.search(r"Number of iterations:\s+(\d+)", out.stdout)
            if out.returncode != 0 or not m or not it:
                raise RuntimeError("reference run failed: " + out.stderr[-400:] + out.stdout[-400:])
            rn = [float(t) for t in open(os.path.join(d, "rData.dat")).read().split()]
            er = [float(t) for t in open(os.path.join(d, "eData.dat")).read().split()]
        return {"iters": int(it.group(1)), "wall": float(m.group(1)), "rnorm": rn, "error": er, "cores": ncores, "kind": "reference"}
    # no oracle/_ref on this machine: the oracle's restatement of the same loop (kind "port")
    from oracle import Oracle
    o = Oracle(opts)
    t0 = time.perf_counter()
    done, rn = o.solve()
    wall = time.perf_counter() - t0
    _, err = o.postprocess()
    o.close()
    return {"iters": int(done), "wall": wall, "rnorm": [float(x) for x in rn], "error": [float(x) for x in err], "cores": ncores, "kind": "port"}


def run_reference(npts, levels, cycles, threads=None):
    """cycle-0 workload of this bench on the host cores: returns (V-cycles/s, cores, kind, walltime, cycles)."""
    r = run_reference_opts(options(npts, levels, cycles), threads)
    run_reference.last = r
    return r["iters"] / r["wall"], r["cores"], r["kind"], r["wall"], r["iters"]


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if a.workload == "cg":
        return reference_arm_cg(a)
    cycles = max(1, min(a.steps, 3))
    n = NPTS - 2
    v, cores, kind, wall, done = run_reference(NPTS, LEVELS, cycles)
    r = run_reference.last
    sample = (f"{done} V-cycles of the full {NPTS}^2 / {LEVELS}-level workload, cycle loop only as the reference times it "
              f"(src/solver.c:1526-1553); {kind}: reference src/*.c over the in-repo mini-PETSc (real PETSc absent), "
              f"OpenMP over {cores} host threads")
    save_record(f"vcycle_{NPTS}", {"host": os.uname().nodename, "npts": NPTS, "levels": LEVELS, "cycles": done, "rnorm": r["rnorm"],
                                   "value": v, "cores": cores, "kind": kind, "wall": wall, "sample": sample})
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "V-cycles/s",
            "n_gpus": a.gpus, "steps": done, "warmup": 0, "ms_per_step": 1e3 * wall / done, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "unknowns": n * n},
            "cpu_baseline": {"value": v, "unit": "V-cycles/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rnorm": r["rnorm"][: done + 1], "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def parity_block(rn_gpu, rec):
    """per-cycle relative residual norms of the B200 arm against the reference run's rData.dat on the same input"""
    k = min(len(rn_gpu), len(rec["rnorm"]), rec["cycles"] + 1)
    diffs = [abs(float(rn_gpu[i]) - rec["rnorm"][i]) / abs(rec["rnorm"][i]) for i in range(k)]
    return {"against": f"oracle/_ref/poisson_ref ({rec['kind']}: reference src/*.c over mini-PETSc) on this box, same {rec['npts']}^2 input",
            "cycles": k - 1, "max_rel_diff": max(diffs), "tol": PARITY_TOL, "ok": bool(max(diffs) <= PARITY_TOL),
            "rnorm_b200": [float(x) for x in rn_gpu[:k]], "rnorm_reference": rec["rnorm"][:k]}


def cpu_block(a):
    """cpu_baseline + the record the parity check runs against: what `--impl reference` left on this box, else run it now"""
    rec = load_record(f"vcycle_{NPTS}")
    how = "measured by the preceding `bench.py --impl reference` run on this box"
    if rec is None:
        if a.no_cpu_baseline or a.profile:
            return None, None
        v, cores, kind, wall, done = run_reference(NPTS, LEVELS, 3)
        r = run_reference.last
        rec = {"host": os.uname().nodename, "npts": NPTS, "levels": LEVELS, "cycles": done, "rnorm": r["rnorm"], "value": v,
               "cores": cores, "kind": kind, "wall": wall,
               "sample": (f"{done} V-cycles of the full {NPTS}^2 / {LEVELS}-level workload, cycle loop only (src/solver.c:1526-1553); "
                          f"{kind}: reference src/*.c over the in-repo mini-PETSc (real PETSc absent), OpenMP over {cores} host threads")}
        save_record(f"vcycle_{NPTS}", rec)
        how = "measured inside this run"
    cpu = {"value": rec["value"], "unit": "V-cycles/s", "cores": rec["cores"], "kind": rec["kind"],
           "sample": rec["sample"] + "; " + how + " (mini-PETSc, not real PETSc: a stated baseline)"}
    return cpu, rec


def bits_fingerprint(u):
    """partition-independent, bit-sensitive checksum of an fp64 array: (sum, xor) of the 64-bit patterns"""
    import numpy as np
    w = np.ascontiguousarray(u, dtype="<f8").view(np.uint64).reshape(-1)
    return int(np.add.reduce(w, dtype=np.uint64)), int(np.bitwise_xor.reduce(w))



# ---------------------------------------------------------------------------------------------- CG workload
MGC1 = "-mg_coarse_ksp_type richardson -mg_coarse_pc_type jacobi -mg_coarse_ksp_max_it 1"   # exact on the 1x1 coarsest grid


def cg_levels(npts):
    return (npts - 1).bit_length() - 1          # down to the 1 x 1 grid: 12 levels at 4097, 13 at 8193


def cg_options(npts, iters=100, extra=""):
    L = cg_levels(npts)
    return (f"-npts {npts} -mesh 0 -iter {iters} -grids {L} -levels {L} -cycle 8 -map 2 -v 3,3 -moreNorm 0 "
            f"-ksp_type cg -ksp_rtol 1e-10 {MGJ} {MGC1} {extra}").strip()


def cg_workload(npts):
    return (f"2D Poisson {npts}^2 fp64, CG preconditioned by one {cg_levels(npts)}-level V(3,3) cycle (Richardson+Jacobi 0.8), "
            f"to 1e-10 relative residual (BASELINE configs[2] / north-star time-to-solution)")


CG_METRIC = "time to 1e-10 relative residual, MG-preconditioned CG (fp64)"


def reference_arm_cg(a):
    r = run_reference_opts(cg_options(a.npts))
    n = a.npts - 2
    ms = 1e3 * r["wall"]
    sample = (f"the full solve: {r['iters']} CG iterations to 1e-10 at {a.npts}^2, KSPSolve bracket as the reference times it "
              f"(src/solver.c:1958-1968); {r['kind']}: reference src/*.c over the in-repo mini-PETSc (real PETSc absent), OpenMP over "
              f"{r['cores']} host threads")
    save_record(f"cg_{a.npts}", {"host": os.uname().nodename, "npts": a.npts, "iters": r["iters"], "rnorm": r["rnorm"], "wall": r["wall"],
                                 "cores": r["cores"], "kind": r["kind"], "sample": sample})
    line = {"impl": "reference", "metric": CG_METRIC, "value": ms, "unit": "ms", "n_gpus": a.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cg_workload(a.npts), "unknowns": n * n}, "iterations": r["iters"],
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rnorm": r["rnorm"][: r["iters"] + 1], "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def b200_arm_cg(a, mgb):
    """Solve time to 1e-10 with the MG-preconditioned CG (cycle 8) at 1..8 GPUs, iteration count and residual history beside
    the same-box CPU solve of the reference (src/solver.c:1884-1989)."""
    import numpy as np
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = 0
    n = a.npts - 2
    opts = cg_options(a.npts, extra="-mgb_csr 0")
    if world > 1:
        strips = importlib.import_module("multigrid-petsc_b200.strips")
        s = strips.StripSession(opts)
        rank = s.rank
        import torch.distributed as dist
    else:
        s = mgb.Session(opts)
    e = s.engine
    jac = mgb.jacobi(0.8)
    steps, warm = max(a.steps, 1), max(a.warmup, 3)

    def solve():
        return e.solve_pcmg(mgb.KSP_CG, jac, 3, coarse=mgb.COARSE_RICHARDSON, coarse_smoother=mgb.jacobi(1.0), coarse_its=1,
                            rtol=1e-10, max_iter=100)
    for _ in range(warm):
        it, rn, reason, _ = solve()
    l0 = e.launch_count()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    times = []
    for _ in range(steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        it, rn, reason, _ = solve()
        torch.cuda.synchronize()
        ms = e.last_solve_ms()
        if world > 1:
            ms = strips._max_over_ranks(ms)
        times.append(ms)
    launches = e.launch_count() - l0
    t_end = time.time() + (0.0 if a.profile else 1.0)
    k_busy = 0
    while True:
        go = 1.0 if time.time() < t_end else 0.0
        if world > 1:
            go = strips._max_over_ranks(go)
        if go == 0.0 or k_busy > 200:
            break
        solve()
        k_busy += 1
    clocks = clk.stop() if rank == 0 else None
    # fingerprint of the solution (partition independent); at N > 1 the per-rank checksums are combined on rank 0
    r0, r1 = e.local_rows(0)
    fsum, fxor = bits_fingerprint(e.get_vec(mgb.VEC_U, 0)[r0:r1])
    if world > 1:
        launches = strips._sum_over_ranks(launches)
        parts = [None] * world
        dist.all_gather_object(parts, (fsum, fxor))
        fsum = sum(p[0] for p in parts) & 0xFFFFFFFFFFFFFFFF
        fxor = 0
        for p_ in parts:
            fxor ^= p_[1]
    s.close()
    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return 0
    ms = float(np.median(times))
    # the CPU solve of the reference on this box: iterations and time (from the preceding --impl reference run, else run now)
    rec = load_record(f"cg_{a.npts}")
    if rec is None and not a.no_cpu_baseline and not a.profile:
        r = run_reference_opts(cg_options(a.npts))
        rec = {"host": os.uname().nodename, "npts": a.npts, "iters": r["iters"], "rnorm": r["rnorm"], "wall": r["wall"], "cores": r["cores"],
               "kind": r["kind"], "sample": f"the full solve at {a.npts}^2 on {r['cores']} host threads, measured inside this run"}
        save_record(f"cg_{a.npts}", rec)
    line = {"metric": CG_METRIC, "value": ms, "unit": "ms", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cg_workload(a.npts), "unknowns": n * n, "parallelism": f"{world} row strip(s)",
                       "l2": "inputs larger than L2" if 8.0 * n * n / world > 126e6 else "per-strip fine vectors fit L2; solves run back to back"},
            "clocks": clocks, "iterations": it, "converged_reason": reason, "final_relative_residual": float(rn[-1]),
            "ms_per_iteration": ms / max(it, 1), "ms_min": float(min(times)), "gpu_launches": int(launches),
            "fingerprint": {"u_sum64": f"{fsum:#018x}", "u_xor64": f"{fxor:#018x}"}}
    if rec:
        k = min(len(rn), len(rec["rnorm"]))
        # rnorm[] is normalised by rnorm[0]: an entry cannot be resolved below one fp64 epsilon of the initial residual (the
        # recursively updated CG residual near 1e-11 differs by ~1e-20 absolute between summation orders): same rule as tests/
        atol = 2.0 ** -52
        diffs = [abs(float(rn[i]) - rec["rnorm"][i]) / abs(rec["rnorm"][i]) for i in range(k)]
        hist_ok = all(abs(float(rn[i]) - rec["rnorm"][i]) <= PARITY_TOL * abs(rec["rnorm"][i]) + atol for i in range(k))
        line["cpu_baseline"] = {"value": 1e3 * rec["wall"], "unit": "ms", "cores": rec["cores"], "kind": rec["kind"], "sample": rec["sample"],
                                "iterations": rec["iters"]}
        line["parity"] = {"against": "oracle/_ref/poisson_ref on this box, same input", "iterations_b200": it, "iterations_reference": rec["iters"],
                          "iterations_equal": bool(it == rec["iters"]), "history_entries": k, "max_rel_diff": max(diffs), "tol": PARITY_TOL, "atol": atol,
                          "ok": bool(it == rec["iters"] and hist_ok)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------- per-level sweep
def level_sweep(a, mgb):
    """BASELINE configs[4], second half: smoother and residual bandwidth on EVERY level of the hierarchy, B200 kernels
    (CUDA events, mgb_time_op) beside the reference's data path on the host cores (cpu_baseline leg: the CPU checker's
    PETSc restatement -- assembled CSR MatMult + separate vector passes, OpenMP over all host threads)."""
    import numpy as np
    npts = a.npts
    levels = cg_levels(npts)
    n = npts - 2
    e = mgb.Engine(levels, n)
    e.set_poisson_uniform()
    x = np.linspace(0, 1, npts)[1:-1]
    e.set_rhs_separable(-2 * np.pi ** 2 * np.sin(np.pi * x), np.sin(np.pi * x))
    e.solve_vcycle(mgb.jacobi(0.8), 3, 3, max_iter=2, rtol=0.0)          # non-trivial data on every level
    l0 = e.launch_count()
    peak, peak_kind = hbm_peak()
    rows = []
    for l in range(levels):
        ni, nj = e.dims(l)
        row = {"level": l, "n": ni}
        for op in ("jacobi", "residual", "fused_down", "fused_up"):
            if op.startswith("fused") and l == levels - 1:
                continue
            ms = e.time_op(op, l, a.steps)
            bpu = mgb.FUSED_OWN_BYTES.get(op, mgb.OPS[op][1])
            row[op] = {"us": ms * 1e3, "gbs": bpu * ni * nj / (ms * 1e-3) / 1e9, "frac": bpu * ni * nj / (ms * 1e-3) / 1e9 / peak}
        rows.append(row)
    launches = e.launch_count() - l0
    e.close()
    cpu = None
    if not a.no_cpu_baseline:
        from oracle import Oracle                                            # cpu_baseline leg (the checker as the CPU data path)
        o = Oracle(options(npts, levels, 1))
        cores = os.cpu_count() or 1
        for l, row in enumerate(rows):
            ni, nj = o.dims(l)
            rng = np.random.default_rng(l)
            b = rng.standard_normal(ni * nj); xx = rng.standard_normal(ni * nj)
            reps = 2 if ni > 4000 else (5 if ni > 1000 else 20)
            o.smooth(l, b, xx, 1, False); o.residual(l, b, xx)
            t0 = time.perf_counter()
            for _ in range(reps):
                o.smooth(l, b, xx, 1, False)
            t1 = time.perf_counter()
            for _ in range(reps):
                o.residual(l, b, xx)
            t2 = time.perf_counter()
            # bytes of the CSR data path per unknown: MatMult 80 + the vector passes of one Richardson+Jacobi
            # iteration (r = b - t: 24, z = r * dinv: 24, x += s z: 24) resp. of KSPBuildResidual (24)
            row["cpu_jacobi"] = {"us": (t1 - t0) / reps * 1e6, "gbs": 152.0 * ni * nj / ((t1 - t0) / reps) / 1e9}
            row["cpu_residual"] = {"us": (t2 - t1) / reps * 1e6, "gbs": 104.0 * ni * nj / ((t2 - t1) / reps) / 1e9}
            row["speedup_jacobi"] = row["cpu_jacobi"]["us"] / row["jacobi"]["us"]
            row["speedup_residual"] = row["cpu_residual"]["us"] / row["residual"]["us"]
        o.close()
        cpu = {"value": rows[0]["cpu_jacobi"]["gbs"], "unit": "GB/s (fine-level Jacobi sweep, CSR data path: 152 B/unknown)", "cores": cores, "kind": "port",
               "sample": f"one smoother sweep and one residual on every level of the {npts}^2 hierarchy, 2-20 repetitions each, the CPU "
                         "checker's PETSc restatement (mini-PETSc, not real PETSc) incl. one vector copy per call"}
    line = {"metric": "smoother / residual HBM GB/s per level (fp64)", "value": rows[0]["jacobi"]["gbs"], "unit": "GB/s", "n_gpus": 1,
            "steps": a.steps, "warmup": 2, "ms_per_step": rows[0]["jacobi"]["us"] / 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"per-level smoother / residual bandwidth sweep, {npts}^2, {levels} levels (BASELINE configs[4])",
                       "unknowns": n * n, "l2": "levels above 2047^2 exceed L2; smaller levels are L2-resident (stated per row by n)"},
            "roofline": {"bound": "hbm", "kernel": "k_stream5<ST_JACOBI> level 0", "achieved": rows[0]["jacobi"]["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": rows[0]["jacobi"]["frac"], "peak_kind": peak_kind, "traffic": None},
            "levels": rows, "cpu_baseline": cpu, "gpu_launches": int(launches)}
    print(json.dumps(line), flush=True)
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
    if a.workload == "cg":
        return b200_arm_cg(a, mgb)
    if a.workload == "sweep":
        return level_sweep(a, mgb)
    if a.workload == "weak":
        strips = importlib.import_module("multigrid-petsc_b200.strips")
        return strips.bench_weak(a, ClockSampler, hbm_peak)
    if world > 1:
        from importlib import import_module
        strips = import_module("multigrid-petsc_b200.strips")
        return strips.bench_strips(a, NPTS, LEVELS, ClockSampler, hbm_peak, options, sys.modules[__name__])

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
    fp_sum, fp_xor = bits_fingerprint(e.get_vec(mgb.VEC_U, 0))       # the iterate after exactly `steps` cycles from u = 0
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
    # ---- CPU baseline + parity at the full size: the reference binary's first cycles on the same 8193^2 input
    cpu, rec = cpu_block(a)
    traffic, traffic_src = ncu_traffic()
    line = {"metric": METRIC, "value": value, "unit": "V-cycles/s", "n_gpus": 1, "steps": steps,
            "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "unknowns": n * n, "l2": "inputs larger than L2 (537 MB per fine vector)", "parallelism": "1 strip"},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_jfused<3,PRE_GIVEN,POST_RESTRICT> level 0 (3 Jacobi sweeps + residual + restriction, one pass)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_kind": peak_kind,
                         "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": FUSED_DOWN_BYTES_PER_UNKNOWN * n * n,
                         "limiter": "HBM is the roofline this fraction is quoted against, but not the limiter: ncu shows DRAM ~63 % busy, the "
                                    "L1/shared-memory data pipe ~71 %, fp64 pipe ~39 %, issue slots ~50 % at 4 warps per scheduler",
                         "note": "temporal blocking: the kernel's own compulsory traffic is 26 B/unknown; by the unfused SURVEY 8d count "
                                 "(90 B/unknown for the 3 sweeps + residual + restriction it replaces) it delivers "
                                 f"{FUSED_DOWN_UNFUSED_BYTES * n * n / (t_f * 1e-3) / 1e9:.0f} GB/s-equivalent "
                                 "(DESIGN.md section 4). The one-sweep kernels it replaces are listed in fine_level_ops.",
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
            "final_relative_residual": float(rn[-1]),
            "fingerprint": {"cycles": steps, "u_sum64": f"{fp_sum:#018x}", "u_xor64": f"{fp_xor:#018x}",
                            "note": "sum and xor of the 64-bit patterns of the fine-level iterate after `steps` cycles from u = 0: "
                                    "independent of the strip partition, identical at every N when the strips are bit-identical"}}
    if cpu:
        line["cpu_baseline"] = cpu
    if rec:
        line["parity"] = parity_block(rn, rec)
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="strong", choices=["strong", "weak", "cg", "sweep"],
                    help="strong (default): 8193^2 split over the GPUs (BASELINE configs[3]); weak: 4096 x 4097 points per GPU (configs[4]); "
                         "cg: MG-preconditioned CG to 1e-10 (configs[2] and the north-star time-to-solution), --npts 4097|8193")
    ap.add_argument("--npts", type=int, default=8193, help="grid points per side for --workload cg")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: no busy loop, one e2e solve, no CPU baseline")
    a = ap.parse_args()
    if a.impl == "reference":
        return reference_arm(a)
    return b200_arm(a)


if __name__ == "__main__":
    sys.exit(main())
