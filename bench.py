#!/usr/bin/env python3
"""bench.py -- IPCS time steps per second, 3D Taylor-Green P2-P1 on an N^3 box (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mesh 96] [--workload cavity]

One "step" = one ``FractionalStep_AB_CN.solve(dt, nu, max_iter=1)`` (fracstep.py:660-696) on the
z-extruded Taylor-Green problem of SURVEY.md 8(d): box [-1,1]^3, N^3 x 6 Kuhn tetrahedra, P2-P1,
nu = 0.01, dt = 0.005, BiCGStab+Jacobi (tentative velocity), CG+multigrid (pressure, null space projected),
CG+Jacobi (mass), rtol 1e-10, initial guesses extrapolated from the solution histories (KRYLOV below; the CPU arm uses
the same options except a Jacobi-preconditioned pressure CG).  Default: the reference demo's 100 steps after 3 warm-up
steps; the first ~30 steps carry the start-up transient (more Krylov iterations), so short runs report fewer steps/s.

  value   device-timed steps/s with every input (state, BC values of all timed steps) resident in HBM;
  e2e     the same steps through the public Python API (host evaluation of the callable BCs, H2D of
          the BC values, D2H of the step result), wall clock between two device synchronisations;
  roofline  the dominant kernel (CSR SpMM of the P2xP2 operator on gdim right-hand sides) timed live
          with CUDA events on the context's stream, algorithmic bytes / time vs MEASURED_PEAKS.json;
  cpu_baseline  the CPU restatement (oracle/) on the box's host cores on a bounded sample.

``--impl reference`` times the CPU restatement only (the DOLFINx/PETSc reference cannot be installed
here: DESIGN.md).  No torch: ranks are launched by torchrun but only RANK/WORLD_SIZE are read.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "IPCS steps/s, 3D Taylor-Green P2-P1 box"
DT, NU = 0.005, 0.01
KRYLOV = {
    "tentative": {"ksp_type": "bcgs", "pc_type": "jacobi", "ksp_rtol": 1e-10, "ksp_initial_guess_nonzero": True,
                  "b200_guess": "extrapolate2", "b200_block_rtol": True},
    "pressure": {"ksp_type": "cg", "pc_type": "mg", "ksp_rtol": 1e-10, "ksp_initial_guess_nonzero": True,
                 "b200_guess": "extrapolate"},
    "scalar": {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-10, "ksp_initial_guess_nonzero": True,
               "b200_guess": "extrapolate2", "b200_block_rtol": True},
}


def ncu_traffic(mesh: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_spmm launch from the committed `ncu --set full`
    capture (profiles/r01_ncu_spmm_final_96cube.txt); only valid for the mesh it was taken on."""
    if mesh != 96:
        return None
    try:
        tot = 0.0
        for line in open(os.path.join(ROOT, "profiles", "r01_ncu_spmm_final_96cube.txt")):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[2]]
        return tot or None
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def step_algorithmic_bytes(n2, n1, nnz22, nnz21, nnz11, gdim, its, bytes_assemble, bytes_spmm, bytes_spmv_q):
    """Algorithmic bytes of ONE IPCS step with the measured Krylov iteration counts its = (k_u, k_p, k_m): the
    per-kernel figures of SURVEY.md 8(d) (k-RHS SpMM, CSR-value passes, 8 bytes per vector entry read or written)
    summed over the kernels a step launches.  An accounting aid for `step_roofline`, not a
    measurement."""
    k_u, k_p, k_m = its
    vec = 8.0 * gdim * n2  # one pass over a velocity-space vector (all components)
    q = 8.0 * n1           # one pass over a pressure-space vector
    rect = nnz21 * (8.0 * gdim + 4.0)  # one pass over a rectangular operator family ([nnz][gdim] values + columns)
    tentative = (rect + 2 * vec + q) + 4 * vec + (bytes_spmm + 5 * vec) + k_u * (2 * bytes_spmm + 14 * vec) + 6 * vec
    pressure = (rect + vec + q) + 6 * q + k_p * (1.15 * (3 * bytes_spmv_q + 10 * q) + 6 * q) + 4 * q
    update = (bytes_spmm) + (rect + 2 * vec + q) + 7 * vec + (bytes_spmm + 5 * vec) + k_m * (bytes_spmm + 10 * vec) + 5 * vec
    state = 2 * vec + 2 * q  # u1 <- u, p <- ps
    return {"assemble_first": bytes_assemble, "tentative": tentative, "pressure": pressure, "update": update,
            "total": bytes_assemble + tentative + pressure + update + state}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
            sm = [float(r[0]) for r in rows if len(r) >= 7]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][1])
                out["power_w_max"] = max(float(r[2]) for r in rows if len(r) >= 7)
                out["samples"] = len(sm)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for k, nm in enumerate(names):
                    if any(r[3 + k].strip() == "Active" for r in rows if len(r) >= 7):
                        out["reasons"].append(nm)
        except Exception:
            pass
        return out


def bc_series(solver, tg, t_first: float, n_steps: int):
    """g_i at the merged BC dofs for n_steps consecutive steps, evaluated on the host ahead of time."""
    out = []
    for i, bcl in enumerate(solver._bcs_u):
        bc = bcl[0]
        own = bc._dofs < solver._nV_owned  # the device list holds the owned BC dofs only
        xT = np.ascontiguousarray(bc._xT[:, own])
        vals = np.empty((n_steps, int(own.sum())))
        for s in range(n_steps):
            tg.t_u = t_first + s * DT
            vals[s] = bc._value(xT)
        out.append(vals)
    return out


# ---- second workload: lid-driven cavity (BASELINE.json configs[4]), unit cube, Re = 1000 ---------------------
CAVITY_NU, CAVITY_DT = 1.0e-3, 0.005


def _cavity_markers():
    lid = lambda x: np.isclose(x[2], 1.0)
    walls = lambda x: (np.isclose(x[0], 0) | np.isclose(x[0], 1) | np.isclose(x[1], 0) | np.isclose(x[1], 1) | np.isclose(x[2], 0)) & ~lid(x)
    return lid, walls


def make_cavity_solver(N: int, comm, device: int):
    """Unit cube, u = (1, 0, 0) on the lid z = 1, no slip on the other walls, no pressure BC, start from rest
    (the set-up of tests/test_gpu_parity.py::test_lid_driven_cavity_matches_oracle at Re = 1000)."""
    import oasisx_b200 as oasisx
    from oasisx_b200 import mesh as bmesh

    msh = bmesh.create_unit_cube(comm, N, N, N)
    lid, walls = _cavity_markers()
    G = oasisx.LocatorMethod.GEOMETRICAL
    bcs_u = [[oasisx.DirichletBC(0.0, G, walls), oasisx.DirichletBC(1.0 if k == 0 else 0.0, G, lid)] for k in range(3)]
    s = oasisx.FractionalStep_AB_CN(msh, ("Lagrange", 2), ("Lagrange", 1), bcs_u=bcs_u, bcs_p=[], solver_options=KRYLOV,
                                    options={"low_memory_version": False}, device=device)
    return msh, s


def cpu_sample_cavity(n_cpu: int, n_steps: int, n_warm: int = 1):
    from oasisx_b200 import fem, mesh as bmesh
    from oracle import ipcs_cpu as cpu

    msh = bmesh.create_unit_cube(None, n_cpu, n_cpu, n_cpu)
    V, Q = fem.functionspace(msh, ("Lagrange", 2)), fem.functionspace(msh, ("Lagrange", 1))
    lid, walls = _cavity_markers()
    bd = fem.locate_dofs_geometrical(V, lambda x: lid(x) | walls(x))
    vals = [lambda x: np.where(lid(x), 1.0, 0.0), lambda x: np.zeros_like(x[0]), lambda x: np.zeros_like(x[0])]
    c = cpu.CpuIPCS(msh.geometry.x, msh.geometry.dofmap, 3, V.dofmap.list, Q.dofmap.list, V.tabulate_dof_coordinates(),
                    Q.tabulate_dof_coordinates(), 2, bcs_u=[[(bd, f)] for f in vals],
                    rtol=KRYLOV["tentative"]["ksp_rtol"], nonzero_guess=True, block_rtol=True,
                    extrapolate={"extrapolate": 1, "extrapolate2": 2}.get(KRYLOV["tentative"].get("b200_guess"), 0))
    if KRYLOV["pressure"].get("pc_type") == "mg":
        c.attach_pressure_multigrid(msh)
    for _ in range(n_warm):
        c.solve(CAVITY_DT, CAVITY_NU)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        c.solve(CAVITY_DT, CAVITY_NU)
    return (time.perf_counter() - t0) / n_steps, msh.num_cells, cpu.lib().ipcs_cpu_threads(), c.its.tolist()


def run_cavity(args):
    """Lid-driven cavity at Re = 1000 on an N^3 unit cube (default 128^3 = 53 M dofs on ONE GPU; with several ranks
    the same global mesh is split in z-slabs: strong scaling -- BASELINE.json's "weak scaling" would need a global mesh
    of 128 x 128 x 128 n cubes, which the z-slab provider can build but the host set-up time of this bench does not
    allow).  Same JSON contract as the Taylor-Green line; constant boundary values, so the end-to-end path moves no
    boundary data after the first step."""
    from oasisx_b200.comm import HostComm

    comm = HostComm.from_env()
    rank, world = comm.rank, comm.size
    device = int(os.environ.get("LOCAL_RANK", "0"))
    N, K, W = args.mesh, args.steps, max(args.warmup, 3)
    t_setup = time.perf_counter()
    msh, solver = make_cavity_solver(N, comm if world > 1 else None, device)
    ctx = solver._ctx
    t_setup = time.perf_counter() - t_setup
    dt, nu = CAVITY_DT, CAVITY_NU
    for s in range(W):
        solver.solve(dt, nu, max_iter=1)
    st0 = ctx.stats()
    sampler = ClockSampler(device)
    sampler.start()
    ctx.synchronize()
    comm.Barrier()
    ctx.event_record(0)
    t0 = time.perf_counter()
    its, stage_ms = [], np.zeros(4)
    for s in range(K):
        solver.solve(dt, nu, max_iter=1)  # constant BCs: after the first step this is the device path plus one scalar D2H
        st = ctx.stats()
        its.append((max(st.its_tentative), st.its_pressure, max(st.its_update)))
        stage_ms += [st.ms_assemble_first, st.ms_tentative, st.ms_pressure, st.ms_update]
    ctx.event_record(1)
    ctx.synchronize()
    e2e_s = comm.allreduce((time.perf_counter() - t0) / K, "max")
    ms_total = comm.allreduce(ctx.event_elapsed_ms(0, 1), "max")
    comm.Barrier()
    clocks = sampler.stop()
    st1 = ctx.stats()
    launches = comm.allreduce(int(st1.kernel_launches - st0.kernel_launches))
    umax = comm.allreduce(float(np.abs(solver._u[0].x.array_ro()).max()), "max")
    peak, peak_kind = measured_peaks()
    ms_k, bytes_k = ctx.bench_kernel(3, 20)
    if rank != 0:
        comm.Barrier()
        return
    cpu = None
    if not args.no_cpu and world == 1:
        n_cpu = args.cpu_mesh if args.cpu_mesh > 0 else min(N, 64)
        sec, cells, threads, cits = cpu_sample_cavity(n_cpu, 2, 1)
        sps = (1.0 / sec) * cells / msh.num_cells
        cpu = {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"C++/OpenMP restatement, 2 cavity steps after 1 warm-up on a {n_cpu}^3 cube ({sec:.2f} s/step, its u/p/m {cits})"
                         + ("" if n_cpu == N else f", scaled by cell count to {N}^3 (optimistic for the CPU: its Jacobi-PCG pressure iterations grow with N)"),
               "host_cpus": os.cpu_count()}
    nV = solver._lp.V.n_global if world > 1 else solver._nV_owned
    nQ = solver._lp.Q.n_global if world > 1 else solver._nQ_owned
    line = {
        "metric": "IPCS steps/s, 3D lid-driven cavity P2-P1 box", "value": 1000.0 * K / ms_total, "unit": "steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D lid-driven cavity P2-P1 {N}^3 unit cube, Re=1000 (nu={nu}, lid speed 1), dt={dt}, from rest, max_iter=1, rtol=1e-10",
                   "mesh": N, "cells": msh.num_cells, "dofs": 3 * nV + nQ, "krylov": KRYLOV, "setup_s": t_setup,
                   "l2": "working set per step >> 126 MB L2; no flush needed"},
        "iterations": {"tentative": int(np.median([i[0] for i in its])), "pressure": int(np.median([i[1] for i in its])),
                       "update": int(np.median([i[2] for i in its]))},
        "stage_ms": dict(zip(["assemble_first", "tentative", "pressure", "update"], (stage_ms / K).round(3).tolist())),
        "max_abs_u_x": umax,
        "roofline": {"bound": "hbm", "kernel": "k_spmm<K=3> (P2xP2 SELL-32 operator, 3 right-hand sides)", "achieved": bytes_k / (ms_k * 1e-3) / 1e9,
                     "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": bytes_k / (ms_k * 1e-3) / 1e9 / peak, "traffic": None,
                     "ms_per_launch": ms_k, "algorithmic_bytes": bytes_k},
        "cpu_baseline": cpu,
        "e2e": {"value": 1.0 / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": int(st1.bytes_h2d - st0.bytes_h2d) // K,
                "d2h_bytes_per_step": int(st1.bytes_d2h - st0.bytes_d2h) // K,
                "note": "the timed loop IS the public-API loop (constant boundary values: nothing to prefetch); value = device events, e2e = wall clock"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    comm.Barrier()


def cpu_sample(n_cpu: int, n_steps: int, n_warm: int = 1):
    """Time the CPU restatement (oracle/ipcs_cpu.cpp: C++/OpenMP, CSR, BiCGStab+Jacobi / CG+multigrid / CG+Jacobi: the
    same Krylov options, preconditioners and initial guesses as the GPU arm) on an n_cpu^3 box with all host threads.
    Returns (seconds per step, cells, threads, iterations)."""
    from oasisx_b200 import fem
    from oracle import ipcs_cpu as cpu
    from problems import TaylorGreen, boundary_facets, make_mesh

    tg = TaylorGreen(NU, 3)
    msh = make_mesh(3, n_cpu)
    V, Q = fem.functionspace(msh, ("Lagrange", 2)), fem.functionspace(msh, ("Lagrange", 1))
    bd = fem.locate_dofs_topological(V, 2, boundary_facets(msh))
    c = cpu.CpuIPCS(msh.geometry.x, msh.geometry.dofmap, 3, V.dofmap.list, Q.dofmap.list, V.tabulate_dof_coordinates(),
                    Q.tabulate_dof_coordinates(), 2, bcs_u=[[(bd, f)] for f in tg.components],
                    rtol=KRYLOV["tentative"]["ksp_rtol"], nonzero_guess=KRYLOV["tentative"]["ksp_initial_guess_nonzero"],
                    block_rtol=KRYLOV["tentative"]["b200_block_rtol"],
                    extrapolate={"extrapolate": 1, "extrapolate2": 2}.get(KRYLOV["tentative"].get("b200_guess"), 0))
    if KRYLOV["pressure"].get("pc_type") == "mg":  # the GPU arm's pressure preconditioner on the CPU arm too
        c.attach_pressure_multigrid(msh)
    xV, xQ = V.tabulate_dof_coordinates().T, Q.tabulate_dof_coordinates().T
    tg.t_u = -DT
    for i, f in enumerate(tg.components):
        c.set(cpu.U2, i, f(xV))
    tg.t_u = 0.0
    for i, f in enumerate(tg.components):
        c.set(cpu.U1, i, f(xV))
    tg.t_p = -DT / 2
    c.set(cpu.P, 0, tg.eval_p(xQ))
    for _ in range(n_warm):  # first-touch page faults and cold caches are not part of a step
        tg.t_u += DT
        c.solve(DT, NU)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        tg.t_u += DT
        c.solve(DT, NU)
    sec = (time.perf_counter() - t0) / n_steps
    return sec, msh.num_cells, cpu.lib().ipcs_cpu_threads(), c.its.tolist()


def cpu_mesh_for(args) -> int:
    """Bounded CPU sample: full IPCS steps on the benchmark mesh itself up to 96^3 (about 9 s per step on 16
    cores); larger meshes are sampled at 96^3 and scaled by the cell count (conservative: pressure
    iterations grow with N)."""
    return args.cpu_mesh if args.cpu_mesh > 0 else min(args.mesh, 96)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = cpu_mesh_for(args)
    W, K = max(args.warmup, 0), max(args.steps, 1)
    # keep the whole run within a few minutes (about 7 s per step at 96^3 during the start-up transient, less later): the
    # same 3 warm-up steps as the GPU arm, then up to 12 timed steps
    K_run, W_run = (K, W) if n_cpu < 96 else (min(K, 12), min(W, 3))
    sec, cells, threads, its = cpu_sample(n_cpu, K_run, W_run)
    target_cells = 6 * args.mesh**3
    sps = (1.0 / sec) * cells / target_cells
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "steps/s", "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1000.0 / sps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D Taylor-Green P2-P1 {args.mesh}^3 box (z-extruded exact solution), dt={DT}, nu={NU}, "
                               "max_iter=1, rtol=1e-10", "mesh": args.mesh, "krylov": KRYLOV},
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port",
                         "sample": f"C++/OpenMP restatement (oracle/ipcs_cpu.cpp, same Krylov methods, preconditioners incl. the pressure multigrid, and "
                                   f"initial guesses as the GPU arm), {K_run} full IPCS steps after {W_run} warm-up on a "
                                   f"{n_cpu}^3 box ({sec:.2f} s/step, Krylov its u/p/m {its})"
                                   + ("" if n_cpu == args.mesh else f", scaled by cell count to {args.mesh}^3"),
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference algorithm on the host cores; the DOLFINx/PETSc/MUMPS reference itself "
                "cannot be installed in this image (DESIGN.md)",
    }
    print(json.dumps(line))


def run_ours(args):
    from oasisx_b200.comm import HostComm
    from problems import TaylorGreen, make_mesh, make_solver

    comm = HostComm.from_env()
    rank, world = comm.rank, comm.size
    if world != args.gpus and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    device = int(os.environ.get("LOCAL_RANK", "0"))

    N, K, W = args.mesh, args.steps, max(args.warmup, 3)
    t_setup = time.perf_counter()
    tg = TaylorGreen(NU, 3)
    msh = make_mesh(3, N, comm if world > 1 else None)
    solver = make_solver(msh, 2, tg, DT, solver_options=KRYLOV, device=device)
    ctx = solver._ctx
    t_setup = time.perf_counter() - t_setup
    nbc = comm.allreduce(sum(len(d) for d in solver._bc_dofs))
    from oasisx_b200 import _lib as L
    sell = ctx.pattern_sell(L.PAT_VV)

    # ---- device-timed region: state and the BC values of every step already in HBM -------------
    series = bc_series(solver, tg, DT, W + K)
    for i, v in enumerate(series):
        ctx.set_velocity_bc_series(i, v)
    solver._flush()
    for s in range(W):
        ctx.select_bc_step(s)
        ctx.step(DT, NU, 1e-12, 1)
    st0 = ctx.stats()
    sampler = ClockSampler(device)
    sampler.start()
    ctx.synchronize()
    comm.Barrier()
    profile_range = os.environ.get("B200_PROFILE_RANGE") == "1"  # ncu --profile-from-start off: the timed steps only
    if profile_range:
        ctx.profiler_range(True)
    ctx.event_record(0)
    its = []
    stage_ms = np.zeros(4)
    for s in range(K):
        ctx.select_bc_step(W + s)
        ctx.step(DT, NU, 1e-12, 1)
        st = ctx.stats()
        its.append((max(st.its_tentative), st.its_pressure, max(st.its_update)))
        res0 = (st.res0_tentative, st.res0_pressure, st.res0_update)
        stage_ms += [st.ms_assemble_first, st.ms_tentative, st.ms_pressure, st.ms_update]
    ctx.event_record(1)
    ctx.synchronize()
    if profile_range:
        ctx.profiler_range(False)
    ms_total = comm.allreduce(ctx.event_elapsed_ms(0, 1), "max")  # max over ranks of the device time
    comm.Barrier()
    clocks = sampler.stop()
    st1 = ctx.stats()
    launches = comm.allreduce(int(st1.kernel_launches - st0.kernel_launches))
    ms_per_step = ms_total / K
    value = 1000.0 / ms_per_step

    # ---- end to end through the public API: callable BCs on the host, H2D, D2H ------------------
    # the SAME K steps again (state re-initialised, solution histories forgotten, W untimed steps first), so that
    # `e2e` and `value` see the same Krylov iteration counts
    solver._written(solver._u, solver._u1, solver._u2, solver._p, solver._ps, solver._dp)
    tg.t_u = -DT
    for i, f in enumerate(tg.components):
        solver._u2[i].interpolate(f)
    tg.t_u = 0.0
    for i, f in enumerate(tg.components):
        solver._u1[i].interpolate(f)
        solver._u[i].x.array[:] = 0.0
    tg.t_p = -DT / 2
    solver._p.interpolate(tg.eval_p)
    solver._flush()
    ctx.reset_time_history()
    ctx.select_bc_step(-1)
    for s in range(W):
        tg.t_u += DT
        tg.t_p += DT
        solver.solve(DT, NU, max_iter=1)
    stA = ctx.stats()
    ctx.synchronize()
    comm.Barrier()
    t0 = time.perf_counter()
    e2e_its = []
    for s in range(K):
        tg.t_u += DT
        tg.t_p += DT
        solver.solve(DT, NU, max_iter=1)
        st = ctx.stats()
        e2e_its.append((max(st.its_tentative), st.its_pressure, max(st.its_update)))
    ctx.synchronize()
    e2e_s = comm.allreduce((time.perf_counter() - t0) / K, "max")
    stB = ctx.stats()
    e2e = {"value": 1.0 / e2e_s, "unit": "steps/s",
           "h2d_bytes_per_step": comm.allreduce(int(stB.bytes_h2d - stA.bytes_h2d)) // K,
           "d2h_bytes_per_step": comm.allreduce(int(stB.bytes_d2h - stA.bytes_d2h)) // K,
           "iterations": [int(np.median([i[j] for i in e2e_its])) for j in range(3)]}

    # ---- full-size sanity of the state the timed steps produced (not part of any timing) --------
    # the discrete solution against the analytic Taylor-Green field (nodal interpolant, mass-matrix norm) and the
    # z-component that must stay at round-off (z-extruded field)
    xV = solver._Vi[0][0].tabulate_dof_coordinates().T
    exact = np.stack([f(xV) for f in tg.components], axis=1)  # blocked [n][3] at the time of the last step
    err2 = ctx.l2_diff_sq(L.VEC_U, exact)
    nrm2 = ctx.l2_diff_sq(L.VEC_U, exact * 0.0)
    wmax = comm.allreduce(float(np.abs(solver._u[2].x.array_ro()).max()), "max")
    checks = {"t_end": tg.t_u, "rel_l2_error_u_vs_exact": float(np.sqrt(err2 / nrm2)), "max_abs_w": wmax}

    # ---- roofline of the dominant kernel, measured live (rank 0's share of the rows) ------------
    peak, peak_kind = measured_peaks()
    comm.Barrier()
    ms_k, bytes_k = ctx.bench_kernel(3, 20)   # k_spmm on the P2xP2 pattern (mass operator), gdim RHS
    achieved = bytes_k / (ms_k * 1e-3) / 1e9
    ms_a, bytes_a = ctx.bench_kernel(1, 5)
    ms_q, bytes_q = ctx.bench_kernel(2, 50)
    comm.Barrier()
    roofline = {"bound": "hbm", "kernel": "k_spmm<K=3> (P2xP2 SELL-32 operator, 3 right-hand sides)",
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(N) if world == 1 else None, "ms_per_launch": ms_k, "algorithmic_bytes": bytes_k,
                "other_kernels": {
                    "assemble_first_ms": ms_a, "assemble_first_GBs": bytes_a / (ms_a * 1e-3) / 1e9,
                    "spmv_Ap_GBs": bytes_q / (ms_q * 1e-3) / 1e9, "spmv_Ap_ms": ms_q}}
    step_roofline = None
    try:
        if world == 1:
            med = (int(np.median([i[0] for i in its])), int(np.median([i[1] for i in its])), int(np.median([i[2] for i in its])))
            sb = step_algorithmic_bytes(solver._nV_owned, solver._nQ_owned, ctx.pattern_nnz(L.PAT_VV), ctx.pattern_nnz(L.PAT_VQ),
                                        ctx.pattern_nnz(L.PAT_QQ), 3, med, bytes_a, bytes_k, bytes_q)
            gbs = sb["total"] / (ms_per_step * 1e-3) / 1e9
            step_roofline = {"algorithmic_bytes_per_step": sb["total"], "iterations_assumed": list(med), "achieved": gbs, "peak": peak,
                             "unit": "GB/s", "frac": gbs / peak,
                             "stage_frac": {k: sb[k] / (v * 1e-3) / 1e9 / peak for k, v in
                                            zip(["assemble_first", "tentative", "pressure", "update"], (stage_ms / K).tolist()) if v > 0},
                             "note": "whole step and stages against the HBM roofline: algorithmic bytes (SURVEY.md 8d formulas, median "
                                     "iteration counts) / measured time / peak; the pressure stage is latency-bound by construction"}
    except Exception as exc:  # an accounting aid must never cost the bench line
        step_roofline = {"error": repr(exc)}
    halos = int(st1.halo_exchanges - st0.halo_exchanges) // K
    allred = int(st1.allreduces - st0.allreduces) // K
    if rank != 0:
        comm.Barrier()
        return

    # ---- CPU restatement on the host cores, bounded sample (rank 0, N = 1 only) -----------------
    cpu = None
    if not args.no_cpu and world == 1:
        try:
            n_cpu = cpu_mesh_for(args)
            n_cpu_steps, n_cpu_warm = (2, 1) if n_cpu < 96 else (3, 3)  # 96^3: steps 4-6 after the GPU arm's 3 warm-up steps
            sec, cells, threads, cits = cpu_sample(n_cpu, n_cpu_steps, n_cpu_warm)
            sps = (1.0 / sec) * cells / msh.num_cells
            cpu = {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port",
                   "sample": f"C++/OpenMP restatement (oracle/ipcs_cpu.cpp, same Krylov methods, preconditioners incl. the pressure "
                             f"multigrid, and initial guesses as the GPU arm), {n_cpu_steps} full IPCS step(s) after {n_cpu_warm} warm-up on "
                             f"a {n_cpu}^3 box ({sec:.2f} s/step, Krylov its u/p/m {cits})"
                             + ("" if n_cpu == N else f", scaled by cell count to {N}^3"),
                   "note": "bounded sample: the CPU steps are the first steps after the warm-up (start-up transient, more Krylov iterations per "
                           "step: see its u/p/m) while `value` averages all timed steps of the GPU run (`iterations`); at equal "
                           "iteration counts the CPU step would be shorter by about the ratio of the velocity iteration counts",
                   "host_cpus": os.cpu_count()}
        except Exception as exc:  # the GPU measurements above must not be lost to a failure of the CPU leg
            cpu = {"value": None, "unit": "steps/s", "kind": "port", "error": repr(exc)}

    nV = solver._lp.V.n_global if world > 1 else solver._nV_owned
    nQ = solver._lp.Q.n_global if world > 1 else solver._nQ_owned
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D Taylor-Green P2-P1 {N}^3 box (z-extruded exact solution), dt={DT}, nu={NU}, "
                               "max_iter=1, rtol=1e-10", "mesh": N, "cells": msh.num_cells,
                   "dofs": 3 * nV + nQ, "partition": f"{world} z-slab(s), NCCL halo + all-reduce" if world > 1 else "single GPU",
                   "l2": "working set per step >> 126 MB L2 (P2xP2 operators alone "
                         f"{3 * 12 * (230 * N**3) / 1e9:.2f} GB over all ranks); no flush needed",
                   "krylov": KRYLOV, "multigrid": "V(1,1) damped Jacobi 0.85, exact dense solve on the first level <= 5000 dofs",
                   "sell_P2xP2": {"slots": sell[0], "run_slice_columns": sell[1], "slice_columns": sell[0] // 32},
                   "setup_s": t_setup},
        "iterations": {"tentative": int(np.median([i[0] for i in its])), "pressure": int(np.median([i[1] for i in its])),
                       "update": int(np.median([i[2] for i in its]))},
        "initial_rel_residual": dict(zip(["tentative", "pressure", "update"], [float(f"{r:.3e}") for r in res0])),
        "stage_ms": dict(zip(["assemble_first", "tentative", "pressure", "update"], (stage_ms / K).round(3).tolist())),
        "nccl_per_step": {"halo_exchanges": halos, "allreduces": allred},
        "roofline": roofline, "step_roofline": step_roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "bc_dofs": nbc, "checks": checks,
    }
    print(json.dumps(line), flush=True)
    comm.Barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="timed steps (the reference demo's T/dt = 100, SURVEY.md 8d)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", type=int, default=0, help="cubes per direction (default 96 = BASELINE's metric, 48 = configs[2]; cavity: 128)")
    ap.add_argument("--workload", default="taylor-green", choices=["taylor-green", "cavity"],
                    help="taylor-green = BASELINE.json's metric (the line the driver reads); cavity = configs[4], an extra line")
    ap.add_argument("--cpu-mesh", type=int, default=0, help="box size of the bounded CPU sample (0: min(mesh, 96))")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--pressure-pc", default="mg", choices=["mg", "jacobi"], help="pressure preconditioner of the GPU arm")
    ap.add_argument("--scalar-ksp", default="auto", choices=["auto", "cg", "chebyshev"],
                    help="mass-solve method (auto = cg; chebyshev is reduction-free but needs ~3x the iterations once the "
                         "initial guesses are good)")
    args = ap.parse_args()
    KRYLOV["pressure"]["pc_type"] = args.pressure_pc
    world = int(os.environ.get("WORLD_SIZE", "1"))
    KRYLOV["scalar"]["ksp_type"] = args.scalar_ksp if args.scalar_ksp != "auto" else "cg"
    if args.mesh <= 0:
        args.mesh = 128 if args.workload == "cavity" else 96
    if args.workload == "cavity":
        if args.impl == "reference":
            raise SystemExit("--impl reference times the Taylor-Green metric; the cavity line carries its own cpu_baseline")
        run_cavity(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
