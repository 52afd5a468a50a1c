"""The ctypes mirrors in multigrid-petsc_b200/__init__.py must have exactly the layout of the C structs of
include/mgb200.h and host/pb_api.h (a silent mismatch would corrupt parameters at the C-ABI boundary).  A tiny C
program prints sizeof / offsetof; no GPU involved."""
import ctypes as C
import importlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mgb = importlib.import_module("multigrid-petsc_b200")

PROG = r'''
#include <stdio.h>
#include <stddef.h>
#include "mgb200.h"
#include "pb_api.h"
#define S(T) printf(#T " size %zu\n", sizeof(T))
#define O(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
	S(mgb_smoother); O(mgb_smoother, scale); O(mgb_smoother, omega); O(mgb_smoother, sor_sweep); O(mgb_smoother, sor_its);
	S(mgb_config); O(mgb_config, device); O(mgb_config, rank); O(mgb_config, nranks); O(mgb_config, agglomerate_below); O(mgb_config, emulate);
	S(mgb_vcycle_params); O(mgb_vcycle_params, v0); O(mgb_vcycle_params, max_iter); O(mgb_vcycle_params, rtol);
	O(mgb_vcycle_params, use_graph); O(mgb_vcycle_params, no_fuse); O(mgb_vcycle_params, no_bottom);
	S(mgb_pcmg_params); O(mgb_pcmg_params, rtol); O(mgb_pcmg_params, max_iter); O(mgb_pcmg_params, level_smoother);
	O(mgb_pcmg_params, level_its); O(mgb_pcmg_params, coarse); O(mgb_pcmg_params, coarse_smoother); O(mgb_pcmg_params, coarse_its);
	O(mgb_pcmg_params, no_fuse); O(mgb_pcmg_params, no_bottom); O(mgb_pcmg_params, no_graph);
	S(pb200_result); O(pb200_result, error); O(pb200_result, solve_seconds); O(pb200_result, levels); O(pb200_result, gpu_launches);
	printf("MGB_IPC_HANDLE_BYTES %d\nMGB_NVEC %d\n", MGB_IPC_HANDLE_BYTES, MGB_NVEC);
	return 0;
}
'''

PY = {"mgb_smoother": mgb.Smoother, "mgb_config": mgb.Config, "mgb_vcycle_params": mgb.VcycleParams,
      "mgb_pcmg_params": mgb.PcmgParams, "pb200_result": mgb.RunResult}


def test_ctypes_structs_match_the_c_headers(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text(PROG)
    exe = tmp_path / "abi"
    subprocess.run(["/usr/bin/gcc", "-std=gnu99", "-I", os.path.join(ROOT, "include"), "-I",
                    os.path.join(ROOT, "multigrid-petsc_b200", "host"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        name, *rest = line.split()
        if name == "MGB_IPC_HANDLE_BYTES":
            assert int(rest[0]) == mgb.IPC_HANDLE_BYTES
        elif name == "MGB_NVEC":
            assert int(rest[0]) == 7 and mgb.VEC_Q == 6
        elif rest[0] == "size":
            assert C.sizeof(PY[name]) == int(rest[1]), line
            seen += 1
        else:
            t, f = name.split(".")
            assert getattr(PY[t], f).offset == int(rest[0]), line
            seen += 1
    assert seen > 30
