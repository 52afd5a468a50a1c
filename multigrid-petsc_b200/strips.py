"""Row strips over several GPUs: one process per GPU (torchrun), torch.distributed only for the plumbing.

The data path has no NCCL call: ghost rows, the gather of the first agglomerated level onto rank 0, the broadcast
of its correction and the all-reduce of the norm partials are P2P stores into the peers' HBM followed by a flag
(csrc/mgb_halo.cuh).  What this module does with torch.distributed is what mpiexec + PETSc's communicator setup do
for the reference: rendezvous, the exchange of one 64-byte CUDA IPC handle per rank, barriers around timed regions
and max-over-ranks of the measured times.
"""
import importlib
import json
import os
import time

import numpy as np

_pkg = importlib.import_module("multigrid-petsc_b200")


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores next to GPU `index` (NVML's ideal CPU affinity), so that the pinned host buffers
    it allocates afterwards live on that socket and its PCIe copies do not cross the inter-socket link -- what
    `mpiexec --bind-to` / `numactl` do for the reference's ranks.  Best effort: returns the CPU list or None."""
    if os.environ.get("MGB_NO_AFFINITY"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except Exception:
        return None


def init_distributed(backend=None):
    """Join the torchrun rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            try:
                phys = int(vis.split(",")[local]) if vis else local
            except ValueError:
                phys = local
            bind_to_gpu_numa_node(phys)
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def exchange_handles(mine):
    """all-gather one bytes object per rank; returns their concatenation in rank order"""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(mine))
    if any(len(h) != len(mine) for h in out):
        raise _pkg.MgbError("ranks exported IPC handles of different sizes")
    return b"".join(out)


def connect(engine):
    """Make the peers' arenas addressable: export, all-gather, import (collective)."""
    engine.ipc_connect(exchange_handles(engine.ipc_export()))


def strip_options(options, rank, world, agglomerate=0):
    extra = f" -mgb_ranks {world} -mgb_rank {rank} -mgb_csr 0 -mgb_device {int(os.environ.get('LOCAL_RANK', rank))}"
    if agglomerate:
        extra += f" -mgb_agglomerate {agglomerate}"
    return options + extra


class StripSession(_pkg.Session):
    """Session (host C layer: SetUpProblem .. Assemble) of this rank's strip, connected to its peers."""

    def __init__(self, options, agglomerate=0):
        self.rank, self.world = init_distributed()
        super().__init__(strip_options(options, self.rank, self.world, agglomerate))
        connect(self.engine)


def gather_solution(engine, n, level=0, which=_pkg.VEC_U):
    """Whole-grid solution on every rank, assembled from the per-rank rows (test / post-processing helper)."""
    import torch
    import torch.distributed as dist
    u = engine.get_vec(which, level)                  # only the local rows are filled
    r0, r1 = engine.local_rows(level)
    rows = [None] * dist.get_world_size()
    dist.all_gather_object(rows, (r0, r1, u[r0:r1].copy()))
    full = np.zeros_like(u)
    for a, b, part in rows:
        full[a:b] = part
    return full


def _max_over_ranks(x):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(x):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def bench_strips(a, npts, levels, ClockSampler, hbm_peak, options_fn, bench):
    """bench.py at N > 1: the same 8193^2 V(3,3) workload split into N row strips (strong scaling)."""
    import torch
    import torch.distributed as dist
    rank, world = init_distributed("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = npts - 2
    steps, warm = a.steps, max(a.warmup, 3)
    aggl = int(os.environ.get("MGB_AGGLOMERATE", "0"))       # experiment knob: 0 = the engine's default (511 rows)
    s = StripSession(options_fn(npts, levels, 1000), agglomerate=aggl)
    e = s.engine
    sm = _pkg.jacobi(0.8)
    e.solve_vcycle(sm, 3, 3, max_iter=warm, rtol=0.0)
    l0 = e.launch_count()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    dist.barrier()
    torch.cuda.synchronize()
    it, rn, _ = e.solve_vcycle(sm, 3, 3, max_iter=steps, rtol=0.0)
    torch.cuda.synchronize()
    dist.barrier()
    ms = _max_over_ranks(e.last_solve_ms())
    launches = _sum_over_ranks(e.launch_count() - l0)
    t_end = time.time() + (0.0 if a.profile else 1.0)
    k_busy = 0
    while True:                                       # keep all ranks busy a little longer for the clock sampler
        go = _max_over_ranks(1.0 if time.time() < t_end else 0.0)
        if go == 0.0 or k_busy > 50:
            break
        e.solve_vcycle(sm, 3, 3, max_iter=steps, rtol=0.0)
        k_busy += 1
    clocks = clk.stop() if rank == 0 else None
    value = steps / (ms * 1e-3)
    # bit pattern checksum of the iterate after `steps` cycles from u = 0: must equal the 1-GPU one (bench.py prints it at N = 1)
    r0, r1 = e.local_rows(0)
    parts = [None] * world
    dist.all_gather_object(parts, bench.bits_fingerprint(e.get_vec(_pkg.VEC_U, 0)[r0:r1]))
    fp_sum = sum(p[0] for p in parts) & 0xFFFFFFFFFFFFFFFF
    fp_xor = 0
    for p_ in parts:
        fp_xor ^= p_[1]
    # dominant kernel on this rank's strip: the fused down leg incl. its ghost-row / coarse-rhs exchange (collective)
    dist.barrier()
    torch.cuda.synchronize()
    t_j = _max_over_ranks(e.time_op("fused_down", 0, 20))
    ach = 26.0 * n * n / (t_j * 1e-3) / 1e9          # aggregate over the N strips
    peak, peak_kind = hbm_peak()
    # e2e: a stream of right-hand sides; every rank uploads its rows from pinned memory and reads its rows of u back,
    # the copies of neighbouring solves overlapping the running one (pb200_solve_rhs_many)
    x = np.linspace(0.0, 1.0, npts)[1:-1]
    hosts = []
    for k in range(2):
        bh_t = torch.empty(n * n, dtype=torch.float64, pin_memory=True)
        uh_t = torch.empty(n * n, dtype=torch.float64, pin_memory=True)
        bh = bh_t.numpy().reshape(n, n)
        bh[r0:r1] = np.outer(np.sin(np.pi * x[r0:r1]), -2 * np.pi ** 2 * np.sin(np.pi * x))
        bh[r0:r1] += 1e-3 * np.random.default_rng(1000 * k + r0).standard_normal((r1 - r0, n))
        hosts.append((bh_t, uh_t))
    s.solve_rhs(hosts[0][0].data_ptr(), hosts[0][1].data_ptr())
    nsolve = 2 if a.profile else 8
    bp = [hosts[k % 2][0].data_ptr() for k in range(nsolve)]
    up = [hosts[k % 2][1].data_ptr() for k in range(nsolve)]
    s.solve_rhs_many(bp[:2], up[:2])
    # (a) one Solve() at a time
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cycles1 = 0
    for k in range(1 if a.profile else 3):
        cycles1 += s.solve_rhs(hosts[k % 2][0].data_ptr(), hosts[k % 2][1].data_ptr())["num_iter"]
    torch.cuda.synchronize()
    t_single = _max_over_ranks(time.perf_counter() - t0)
    # where a single Solve() spends its time: upload of this rank's rows, the cycles, download of its rows (max over ranks)
    def timed(fn):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        return 1e3 * _max_over_ranks(time.perf_counter() - t0)
    bview = hosts[0][0].numpy().reshape(n, n)
    uview = hosts[0][1].numpy().reshape(n, n)
    t_up = timed(lambda: e._ck(e.L.mgb_vec_set(e.h, _pkg.VEC_B, 0, _pkg._pd(bview))))
    t_sv = timed(lambda: e.solve_vcycle(sm, 3, 3, max_iter=1000, rtol=1e-7))
    t_dn = timed(lambda: e._ck(e.L.mgb_vec_get(e.h, _pkg.VEC_U, 0, _pkg._pd(uview))))
    # (b) the pipelined stream
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    its, fin, _ = s.solve_rhs_many(bp, up)
    torch.cuda.synchronize()
    t_e2e = _max_over_ranks(time.perf_counter() - t0)
    cycles = sum(its)
    bytes_per_solve = 8.0 * n * n
    s.close()
    # the same fingerprint on ONE GPU (rank 0, same process, after the strips are gone): proves strips == one strip bit for bit
    fp1 = None
    if rank == 0 and not a.profile:
        s1 = _pkg.Session(options_fn(npts, levels, 1000, "-mgb_csr 0"))
        s1.engine.solve_vcycle(sm, 3, 3, max_iter=steps, rtol=0.0)
        fp1 = bench.bits_fingerprint(s1.engine.get_vec(_pkg.VEC_U, 0))
        s1.close()
    if rank == 0:
        rec = bench.load_record(f"vcycle_{npts}")
        line = {"metric": bench.METRIC, "value": value, "unit": "V-cycles/s", "n_gpus": world,
                "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": bench.WORKLOAD,
                           "unknowns": n * n, "l2": "inputs larger than L2 on every strip at N<=4; per-strip fine vector "
                           f"{8.0 * n * n / world / 1e6:.0f} MB", "parallelism": f"{world} row strips, P2P ghost rows over NVLink, "
                           f"levels with <= {aggl or 511} rows agglomerated on rank 0"},
                "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "k_jfused<3,PRE_GIVEN,POST_RESTRICT> level 0 on every strip (+ its ghost-row exchange)", "achieved": ach,
                             "peak": peak * world, "unit": "GB/s", "frac": ach / (peak * world), "peak_kind": peak_kind + f" x {world} GPUs",
                             "traffic": None, "vcycle_gbs_unfused_count": 264.0 * n * n * value / 1e9},
                "e2e": {"value": cycles / t_e2e, "unit": "V-cycles/s", "h2d_bytes_per_step": bytes_per_solve * nsolve / cycles,
                        "d2h_bytes_per_step": bytes_per_solve * nsolve / cycles, "solves": nsolve, "cycles_per_solve": cycles / nsolve,
                        "single_solve_value": cycles1 / t_single,
                        "single_solve_breakdown_ms": {"upload_rhs_rows": t_up, "cycles_to_1e-7": t_sv, "download_solution_rows": t_dn,
                                                      "per_rank_bytes_each_way": bytes_per_solve / world},
                        "note": "a stream of right-hand sides (pb200_solve_rhs_many): per solve every rank uploads its rows of the rhs "
                                "(pinned) + V-cycles to 1e-7 + reads its rows of u, copies overlapping the neighbouring solves; bytes are "
                                "totals over the ranks, per V-cycle"},
                "gpu_launches": int(launches), "final_relative_residual": float(rn[-1]),
                "fingerprint": {"cycles": steps, "u_sum64": f"{fp_sum:#018x}", "u_xor64": f"{fp_xor:#018x}",
                                "one_gpu_u_sum64": None if fp1 is None else f"{fp1[0]:#018x}",
                                "one_gpu_u_xor64": None if fp1 is None else f"{fp1[1]:#018x}",
                                "equals_one_gpu": None if fp1 is None else bool(fp1 == (fp_sum, fp_xor)),
                                "note": "sum and xor of the 64-bit patterns of the fine-level iterate after `steps` cycles from u = 0, "
                                        "combined over the strips; one_gpu_* = the same solve on rank 0's GPU alone in this run"}}
        if rec:
            line["parity"] = bench.parity_block(rn, rec)
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


def selfcheck(npts=1025, levels=10, expect_sha=None, expect_iters=None):
    """Run under torchrun: solve the npts^2 case on WORLD_SIZE strips and compare with the 1-strip golden hash."""
    import hashlib
    import torch.distributed as dist
    rank, world = init_distributed("nccl")
    opts = (f"-npts {npts} -mesh 0 -iter 100000 -grids {levels} -levels {levels} -cycle 0 -map 2 -v 3,3 -moreNorm 0 "
            "-pc_type jacobi -ksp_richardson_scale 0.8")
    s = StripSession(opts)
    it, rn, _ = s.engine.solve_vcycle(_pkg.jacobi(0.8), 3, 3, max_iter=100000, rtol=1e-7)
    u = gather_solution(s.engine, npts - 2)
    sha = hashlib.sha256(np.ascontiguousarray(u, dtype="<f8").tobytes()).hexdigest()
    ok = (expect_sha is None or sha == expect_sha) and (expect_iters is None or it == expect_iters)
    s.close()
    if rank == 0:
        print(json.dumps({"strips": world, "iters": it, "sha256": sha, "final": float(rn[-1]), "ok": bool(ok)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


def bench_weak(a, ClockSampler, hbm_peak, rows_per_gpu=4096, ncols=4097):
    """BASELINE configs[4]: weak scaling, 4096 x 4097 grid points per GPU.  The reference forces square grids
    (ref: src/poisson.c:73-75), so P = 2, 8 are not expressible through the host C layer; this goes through the
    C-ABI engine directly with a rectangular (4096 P - 1) x 4095 grid of unknowns, operator 1/h^2 [1 1 -4 1 1] per
    level as the reference's OpA gives on a uniform mesh, separable synthetic right-hand side."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = 0, int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        rank, world = init_distributed("nccl")
    torch.cuda.set_device(local)
    ni, nj = rows_per_gpu * world - 1, ncols - 2
    levels = 1
    while ((nj + 1) >> levels) - 1 >= 1 and ((ni + 1) >> levels) - 1 >= 1:
        levels += 1
    e = _pkg.Engine(levels, ni, nj, device=local, rank=rank, nranks=world)
    e.set_poisson_uniform()
    if world > 1:
        connect(e)
    y = np.linspace(0.0, 1.0, ni + 2)[1:-1]
    x = np.linspace(0.0, 1.0, nj + 2)[1:-1]
    e.set_rhs_separable(-2 * np.pi ** 2 * np.sin(np.pi * x), np.sin(np.pi * y))
    sm = _pkg.jacobi(0.8)
    steps, warm = a.steps, max(a.warmup, 3)
    e.solve_vcycle(sm, 3, 3, max_iter=warm, rtol=0.0)
    l0 = e.launch_count()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    it, rn, _ = e.solve_vcycle(sm, 3, 3, max_iter=steps, rtol=0.0)
    torch.cuda.synchronize()
    ms = e.last_solve_ms()
    launches = e.launch_count() - l0
    if world > 1:
        dist.barrier()
        ms = _max_over_ranks(ms)
        launches = _sum_over_ranks(launches)
    for _ in range(0 if a.profile else 6):
        e.solve_vcycle(sm, 3, 3, max_iter=steps, rtol=0.0)
    clocks = clk.stop() if rank == 0 else None
    unknowns = float(ni) * nj
    value = steps / (ms * 1e-3)
    peak, peak_kind = hbm_peak()
    e.close()
    if rank == 0:
        line = {"metric": "V-cycles/sec (fp64, weak scaling 4096 x 4097 points per GPU)", "value": value, "unit": "V-cycles/s",
                "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"2D Poisson {ni + 2} x {nj + 2} fp64 (4096 rows per GPU), {levels}-level V(3,3), Richardson+Jacobi 0.8 "
                           "(BASELINE configs[4])", "unknowns": unknowns, "l2": "inputs larger than L2 (134 MB per strip vector)",
                           "parallelism": f"{world} row strips"},
                "clocks": clocks, "unknown_updates_per_s": unknowns * value,
                "roofline": {"bound": "hbm", "kernel": "whole V-cycle, unfused SURVEY 8d count (264 B per fine unknown)",
                             "achieved": 264.0 * unknowns * value / 1e9, "peak": peak * world, "unit": "GB/s",
                             "frac": 264.0 * unknowns * value / 1e9 / (peak * world), "peak_kind": peak_kind + f" x {world} GPUs", "traffic": None},
                "gpu_launches": int(launches), "final_relative_residual": float(rn[-1])}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0
