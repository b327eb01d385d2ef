#!/usr/bin/env python
"""bench.py -- GLS Navier-Stokes operator vmult throughput (GDoF/s), the reference's
performance.cc recipe re-created on synthetic meshes.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference
                                                           # algorithm on the host cores

A "step" is one operator application (vmult) on the configured mesh.  At N = 1 the workload
is BASELINE.json's config "performance.cc synthetic 3D hypercube operator vmult, degree 2":
Cartesian cells, no constraints, nu = 0.1, c1 = 4, c2 = 2, BDF(2) after one update_dt(0.1)
(weight 10), Newton form, cell-wise stabilization, no time-derivative term
(performance.cc:16-24, :44-62), with >= 1e8 DoFs per GPU and random vectors (seed 1234)
instead of the reference's zero vectors.  N > 1: weak scaling, one box of the same size per
rank stacked in z, ghost planes exchanged over NCCL (update_ghost_values / compress(add)).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NU, C1, C2, DT = 0.1, 4.0, 2.0, 0.1
SEED = 1234


def algorithmic_bytes_per_cell(dim, degree, number_bytes, *, newton=True, ctd=False, q_wise=False,
                               general=False):
    """SURVEY.md section 8d / BASELINE.md section 3: the contract figure."""
    n_q = (degree + 1) ** dim
    C = dim + 1
    T_U = dim * dim + 2 * dim if newton else dim
    T_t = dim if ctd else 0
    T_dq = 2 if q_wise else 0
    T_G = dim * dim + 1 if general else 0
    T_cell = (0 if q_wise else 2) + (0 if general else dim + 1)
    D_cell = C * degree ** dim
    s = number_bytes
    return s * (n_q * (T_U + T_t + T_dq + T_G) + T_cell) + 4 * C * n_q + 2 * s * D_cell


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.active = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(self.period)

    def stop(self):
        self._stop.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------
# CPU arm: the C restatement of the reference algorithm on the host cores
# --------------------------------------------------------------------------------------
def cpu_sample(cells_per_dir, degree, workload="P"):
    """The bounded sample of the workload the CPU restatement runs on: mesh, C oracle with its tables, and the
    seeded vectors (history, linearization point, src).  Same flags as the GPU workload, smaller mesh."""
    from dealii_ns_gls_b200 import mesh as gm
    from oracle import gls_oracle as go
    from oracle.gls_oracle_c import COracle

    rng = np.random.default_rng(SEED)
    if workload == "C":
        c = cells_per_dir
        mesh = gm.cylinder_shell((max(2, c // 2), 4 * c, max(2, c // 2)), degree)
        weights, nu, ctd, cell_wise = [15.0, -20.0, 5.0], 0.001, True, False
    else:
        mesh = gm.hypercube(3, cells_per_dir, degree)
        weights, nu, ctd, cell_wise = [10.0, -10.0, 0.0], NU, False, True
    ora = go.OracleOperator(dim=3, degree=degree, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs,
                            cell_points=mesh.cell_points, mapping_degree=mesh.mapping_degree,
                            constraints=mesh.constraints, nu=nu, c1=C1, c2=C2, theta=1.0, order=2,
                            consider_time_derivative=ctd, increment_form=True,
                            cell_wise_stabilization=cell_wise, path="sumfac")
    hist = [rng.uniform(-1, 1, mesh.n_dofs) for _ in range(3)] if workload == "C" else None
    if hist is not None:
        ora.set_previous_solution(hist, weights)
    lin = rng.uniform(-1, 1, mesh.n_dofs)
    ora.set_linearization_point(lin, DT)
    if workload == "C":
        co = COracle.from_numpy_oracle(ora, COracle.BR_NEWTON)
    else:
        K, basis = mesh.n_cells, ora.tb.b
        co = COracle(dim=3, degree=degree, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs, S=basis.S, D=basis.D,
                     w=basis.wq, cartesian=True, inv_jac=mesh.cart_inv_jac, jxw=mesh.cart_det, nu=NU, theta=1.0,
                     branch=COracle.BR_NEWTON, ctd=False, cell_wise=True)
        co.set_tables(ora.U, ora.H.reshape(K, 9, -1), ora.P, None, None, None,
                      ora.delta1_cell.reshape(K, 1), ora.delta2_cell.reshape(K, 1))
    del ora
    src = rng.uniform(-1, 1, mesh.n_dofs)
    return dict(mesh=mesh, co=co, hist=hist, lin=lin, src=src, weight=weights[0], nu=nu, ctd=ctd,
                cell_wise=cell_wise, weights=weights)


def cpu_vmult_gdofs(cells_per_dir, degree, steps, warmup, budget_s=25.0, sample=None):
    """Time oracle/gls_oracle_c on a bounded sample of the workload (same flags, smaller cube)."""
    from oracle.gls_oracle_c import max_threads

    sm = cpu_sample(cells_per_dir, degree) if sample is None else sample
    mesh, co, src = sm["mesh"], sm["co"], sm["src"]
    dst = np.empty_like(src)
    nt = max_threads()
    for _ in range(max(1, warmup)):
        co.apply_into(dst, src, sm["weight"], nt)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        co.apply_into(dst, src, sm["weight"], nt)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=mesh.n_dofs * done / dt / 1e9, steps=done, seconds=dt, cores=nt, n_dofs=mesh.n_dofs,
                sample=f"{cells_per_dir}^3 cells Q{degree} hypercube ({mesh.n_dofs} DoFs), {done} vmults, "
                       f"{nt} OpenMP threads, same flags as the GPU workload")


def gpu_parity_on_sample(sm, number, dev):
    """The CUDA path (through the C ABI) against the C restatement on the CPU sample: same mesh, same seeded
    vectors.  The cell loop of the oracle is plain gather/scatter; constrained rows (config C: no-slip) are
    resolved around it the way vmult does (operator_ns.cc:684-732).  Returns the relative l2 error over the
    unconstrained rows and whether the identity rows are bit-equal."""
    import torch

    from dealii_ns_gls_b200.operator import NavierStokesOperator
    from tests.util import TI
    from oracle.gls_oracle_c import max_threads
    mesh = sm["mesh"]
    tdt = torch.float64 if number == "double" else torch.float32
    ti = TI(2, sm["weights"], DT)
    op = NavierStokesOperator(mesh, None, sm["nu"], C1, C2, ti, sm["ctd"], True, sm["cell_wise"], number=number,
                              device=dev)
    to_dev = lambda a: torch.tensor(a, dtype=tdt, device=dev)  # noqa: E731
    if sm["hist"] is not None:
        op.set_previous_solution([to_dev(h) for h in sm["hist"]])
    op.set_linearization_point(to_dev(sm["lin"]))
    dst = op.initialize_dof_vector()
    op.vmult(dst, to_dev(sm["src"]))
    got = dst.double().cpu().numpy()
    variant = op.vmult_variant()
    del op, dst
    cons = np.fromiter(mesh.constraints.keys(), dtype=np.int64, count=len(mesh.constraints))
    x = sm["src"].copy()
    x[cons] = 0.0  # zero-type rows read 0 (read_dof_values)
    ref = sm["co"].apply(x, sm["weight"], max_threads())
    free = np.ones(mesh.n_dofs, dtype=bool)
    free[cons] = False
    err = float(np.linalg.norm((got - ref)[free]) / np.linalg.norm(ref[free]))
    ident = bool(np.array_equal(got[cons], sm["src"].astype(np.float64 if number == "double" else np.float32)[cons]))
    tol = 1e-12 if number == "double" else 2e-5
    return {"rel_l2_unconstrained_rows": err, "identity_rows_bit_equal": ident, "tol": tol,
            "ok": bool(err < tol and ident), "n_dofs": int(mesh.n_dofs), "n_cells": int(mesh.n_cells),
            "n_constrained": int(len(cons)), "number": number, "kernel_variant": variant,
            "checker": "oracle/gls_oracle_c.c (CPU restatement: its quadrature-point physics is pinned to the reference's own object code, tests/test_reference_qpoint.py; deal.II's parts are restated, unpinned)"}


def gpu_parity_full_size(chk, op, dst, tdt, number):
    """The `parity_full_size` object of the bench line: the operator bench.py has just timed, on fields that repeat
    every 4 cells, against the numpy oracle on a 12^3-cell block of the same mesh size -- every entry of the
    result compared (tests/full_size.py has the argument).  Changes the operator's linearization point."""
    from tests.full_size import oracle_on_small
    (lin_s, lin_b), (src_s, src_b) = chk.field(), chk.field()
    op.set_linearization_point(lin_b.to(tdt))
    op.vmult(dst, src_b.to(tdt))
    ref = oracle_on_small(chk, lin=lin_s, src=src_s, hist=None, nu=NU, c1=C1, c2=C2, weights=[10.0, -10.0, 0.0],
                          dt=DT, ctd=False, cell_wise=True)
    r = chk.compare(dst, ref)
    tol = 1e-12 if number == "double" else 2e-5
    r.update(tol=tol, ok=bool(r["rel_l2_all_rows"] < tol), number=number, kernel_variant=op.vmult_variant(),
             checker="oracle/gls_oracle.py (numpy restatement: quadrature-point physics pinned to the reference's own object code, deal.II's parts unpinned) on the small block, "
                     "mapped to every dof of the big block by periodicity")
    return r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_vmult_gdofs(args.cpu_cells, args.degree, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": "GLS NS operator vmult throughput", "value": r["value"], "unit": "GDoF/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, args.gpus),
                       cpu_sample=f"each step = one vmult on a {args.cpu_cells}^3-cell sample ({r['n_dofs']} DoFs) "
                                  "of this workload; GDoF/s is size-independent once the tables (1.1 GB) are out of cache"),
        "cpu_baseline": {"value": r["value"], "unit": "GDoF/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference algorithm (oracle/gls_oracle_c.c), not deal.II: the "
                "reference cannot be built without deal.II/p4est/Trilinos/MPI",
    }
    emit(line)


def workload_config(args, n_gpus):
    if args.workload == "C":
        nr, nt, nz = shell_shape(args)
        return {"workload": f"config C (input_turek_3D_Re100.json flags): 3D O-grid around a cylinder, Q{args.degree}, "
                            f"{nr}x{nt}x{nz} curved cells, Q{args.degree} mapping (general geometry), no-slip rows, "
                            "Newton form, q-point-wise delta, BDF2 with time-derivative terms, random U/src/history "
                            f"seed 1234, Number = {args.number}",
                "cells_per_gpu": nr * nt * nz, "degree": args.degree, "dim": 3,
                "parallelism": "single GPU" if n_gpus == 1 else
                f"domain decomposition, {n_gpus} boxes of {nr}x{nt}x{nz} cells (axial, then radial cuts; the periodic "
                "circumferential direction is not split)",
                "l2_policy": "inputs larger than L2 (tables + vectors >> 126 MB), no flush"}
    return {"workload": f"performance.cc: 3D hypercube, Q{args.degree}, {args.cells}^3 cells per GPU, "
                        "Cartesian, no constraints, Newton form, cell-wise delta, BDF2 weight 10, "
                        "random U/src seed 1234" + ("" if args.number == "double" else ", Number = float"),
            "cells_per_gpu": args.cells ** 3, "degree": args.degree, "dim": 3,
            "parallelism": "single GPU" if n_gpus == 1 else
            (f"domain decomposition, {n_gpus} z-slabs" if getattr(args, "partition", "morton") == "slab" else
             f"domain decomposition by owner rank along the Morton curve (p4est): {n_gpus} boxes of "
             f"{args.cells}^3 cells = " + {2: "z-halves", 4: "(z,y)-quarters", 8: "octants"}.get(n_gpus, "boxes") +
             ", up to 7 neighbours per rank"),
            "l2_policy": "inputs larger than L2 (tables + vectors >> 126 MB), no flush"}


def shell_shape(args):
    """O-grid of about args.cells^3 cells: radial x circumferential x axial."""
    c = args.cells
    return max(2, c // 2), 4 * c, max(2, c // 2)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def _bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity) before any pinned host memory is
    allocated: with one rank per GPU on a two-socket box the page-locked vectors of the end-to-end leg are
    otherwise first-touched on whatever node the process happens to start on."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass



def run_gpu(args):
    import torch
    import torch.distributed as dist

    from dealii_ns_gls_b200 import mesh as gm
    from dealii_ns_gls_b200.operator import NavierStokesOperator
    from dealii_ns_gls_b200.time_integration import TimeIntegratorDataBDF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _bind_to_gpu_numa_node(local_rank)
    exchange = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from dealii_ns_gls_b200.distributed import GhostExchange
        if args.workload == "C":
            mesh = gm.cylinder_shell_box(shell_shape(args), args.degree, n_ranks=world, rank=rank)
        else:
            part = gm.hypercube_slab if args.partition == "slab" else gm.hypercube_box
            mesh = part(args.cells, args.degree, n_ranks=world, rank=rank, with_points=False)
        exchange = GhostExchange(mesh.partition, dev)
    elif args.workload == "C":
        mesh = gm.cylinder_shell(shell_shape(args), args.degree)
    else:
        mesh = gm.hypercube(3, args.cells, args.degree)

    tdt = torch.float64 if args.number == "double" else torch.float32
    ti = TimeIntegratorDataBDF(2)
    ti.update_dt(DT)  # performance.cc:44-46: weights (10, -10, 0)
    if args.workload == "C":
        ti.update_dt(DT)  # second step: full BDF2 weights (15, -20, 5)
        op = NavierStokesOperator(mesh, None, 0.001, C1, C2, ti, True, True, False, number=args.number, device=dev,
                                  exchange=exchange)
    else:
        op = NavierStokesOperator(mesh, None, NU, C1, C2, ti, False, True, True, number=args.number, device=dev,
                                  exchange=exchange)
    n_local, n_owned = mesh.n_dofs, mesh.n_owned
    n_global = mesh.n_global_dofs
    n_cells = mesh.n_cells
    full_size = None
    if world == 1 and args.workload == "P" and not args.no_extras and args.cells % 4 == 0 and args.cells >= 12:
        # dof map of the every-cell parity check at the bench size (tests/full_size.py), built while the mesh exists
        try:
            from tests.full_size import PeriodicFullSizeCheck
            full_size = PeriodicFullSizeCheck(mesh, dev, period_cells=4)
        except Exception as e:
            full_size = {"error": f"{type(e).__name__}: {e}"}
    del mesh

    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    def rand_vec():
        return (torch.rand(n_local, dtype=torch.float64, device=dev, generator=g) * 2 - 1).to(tdt)

    if args.workload == "C":
        hist = [rand_vec() for _ in range(3)]
    else:
        hist = [torch.zeros(n_local, dtype=tdt, device=dev) for _ in range(3)]  # performance.cc:66-69
    op.set_previous_solution(hist)
    del hist
    lin = rand_vec()
    op.set_linearization_point(lin)
    del lin
    src = rand_vec()
    if n_local > n_owned:
        src[n_owned:] = 0
    dst = op.initialize_dof_vector()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident timed region -------------------------------------------------
    for _ in range(max(3, args.warmup)):
        op.vmult(dst, src)
    barrier()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = op.launch_count()
    sampler.active.set()
    e0.record()
    for i in range(args.steps):
        op.vmult(dst, src, kernel_events=k_ev[i])
    e1.record()
    barrier()
    sampler.active.clear()
    launches = op.launch_count() - l0
    t_ms = e0.elapsed_time(e1)
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    checksum = float(dst[:n_owned].double().abs().sum())

    # ---- end to end: host vectors in, host vector out, every step ----------------------
    h_src = torch.empty(n_local, dtype=tdt, pin_memory=True)
    h_dst = torch.empty(n_local, dtype=tdt, pin_memory=True)
    h_src.copy_(src)
    e2e_steps = max(2, min(args.steps, 10))
    for _ in range(2):
        op.vmult_host(h_dst, h_src)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    f0.record()
    for _ in range(e2e_steps):
        op.vmult_host(h_dst, h_src)
    f1.record()
    barrier()
    sampler.active.clear()
    sampler.stop()
    e2e_ms = f0.elapsed_time(f1)

    if world > 1:
        t = torch.tensor([t_ms, e2e_ms, k_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms, e2e_ms, k_ms = [float(x) for x in t]
    variant = op.vmult_variant()
    # wall time per time step on N > 1 GPUs: a collective piece of work, every rank takes part
    time_step_multi = None
    if world > 1 and args.workload == "P" and args.time_step_refinements >= 0:
        del op, src, dst, h_src, h_dst
        torch.cuda.empty_cache()
        try:
            time_step_multi = time_step_wall(args.time_step_refinements, dev, n_ranks=world, rank=rank)
        except Exception as e:
            time_step_multi = {"error": f"{type(e).__name__}: {e}"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    peaks_src = "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peaks_src = "measured"
    except Exception:
        peaks = {"hbm_gbs": 6650.0}
    nb = 8 if args.number == "double" else 4
    if args.workload == "C":
        bpc = algorithmic_bytes_per_cell(3, args.degree, nb, ctd=True, q_wise=True, general=True)
    else:
        bpc = algorithmic_bytes_per_cell(3, args.degree, nb)
    achieved = n_cells * bpc / (k_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        try:
            with open(tr_path) as f:
                tj = json.load(f)
            if tj.get("cells") == n_cells:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    value = n_global * args.steps / (t_ms * 1e-3) / 1e9
    line = {
        "metric": "GLS NS operator vmult throughput", "value": value, "unit": "GDoF/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": t_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.number == "double" else "f32", "data": "synthetic",
        "config": dict(workload_config(args, world), n_dofs_global=n_global, kernel_variant=variant,
                       checksum_abs_sum=checksum),
        "e2e": {"value": n_global * e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": "GDoF/s",
                "h2d_bytes_per_step": n_local * nb, "d2h_bytes_per_step": n_local * nb, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                     "frac": achieved / peaks.get("hbm_gbs"), "traffic": traffic,
                     "peak_source": peaks_src, "kernel": variant, "kernel_ms": k_ms,
                     "algorithmic_bytes_per_cell": bpc, "cells_per_launch": n_cells},
        "clocks": sampler.summary(),
    }
    if traffic is not None:
        line["roofline"]["traffic_source"] = "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of " \
                                             "one ncu --set full capture of this kernel at this size, not measured in this run"
    headline = world == 1 and args.workload == "P" and args.number == "double"
    if full_size is not None:
        # every cell of THIS operator (the timed one, at the bench size) against the CPU oracle
        try:
            line["parity_full_size"] = full_size if isinstance(full_size, dict) else \
                gpu_parity_full_size(full_size, op, dst, tdt, args.number)
        except Exception as e:
            line["parity_full_size"] = {"error": f"{type(e).__name__}: {e}"}
        full_size = None
    if headline and not args.no_extras:
        # the reference's own protocol next to the random-vector one: zero vectors (performance.cc:66-79)
        try:
            zero = torch.zeros(n_local, dtype=tdt, device=dev)
            op.set_linearization_point(zero)
            zr = side_measure(op, zero, dst, min(args.steps, 10), n_cells, n_global, bpc, peaks.get("hbm_gbs"))
            zr["note"] = "performance.cc:66-79 as written: linearization point, history and src all zero"
            zr["dst_abs_max"] = float(dst.abs().max())
            line["zero_vector_protocol"] = zr
            del zero
        except Exception as e:
            line["zero_vector_protocol"] = {"error": f"{type(e).__name__}: {e}"}
    if time_step_multi is not None:
        line["time_step"] = time_step_multi
    if world == 1:
        del op, src, dst, h_src, h_dst
        torch.cuda.empty_cache()
    if headline and not args.no_cpu_baseline:
        sm = cpu_sample(args.cpu_cells, args.degree)
        r = cpu_vmult_gdofs(args.cpu_cells, args.degree, 10, 2, sample=sm)
        line["cpu_baseline"] = {"value": r["value"], "unit": "GDoF/s", "cores": r["cores"], "kind": "port",
                                "sample": r["sample"]}
        if not args.no_extras:
            # parity of the CUDA path against the C restatement on that same sample, and on a config-C sample
            try:
                line["parity"] = dict(gpu_parity_on_sample(sm, "double", dev), sample=f"{args.cpu_cells}^3-cell hypercube")
                del sm
                smc = cpu_sample(max(8, args.cpu_cells // 2), args.degree, workload="C")
                line["parity_config_C"] = dict(gpu_parity_on_sample(smc, "double", dev),
                                               sample="O-grid shell, Turek-3D flags, no-slip rows")
                del smc
            except Exception as e:
                line["parity"] = {"error": f"{type(e).__name__}: {e}"}
    if headline and not args.no_extras:
        # config C (curved cells, q-point-wise delta, BDF2 terms) at the bench size, same protocol
        try:
            line["workload_C"] = config_c_measure(args, dev, peaks.get("hbm_gbs"))
        except Exception as e:
            line["workload_C"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    if world == 1 and args.workload == "P" and args.time_step_refinements >= 0:
        try:
            line["time_step"] = time_step_wall(args.time_step_refinements, dev)
            line["time_step"]["dofs_per_s"] = line["time_step"]["n_dofs"] / line["time_step"]["wall_s_per_step"]
        except Exception as e:  # the vmult line above stays valid on its own
            line["time_step"] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_cpu_baseline and "error" not in line["time_step"]:
            try:
                line["time_step"]["cpu_baseline"] = time_step_cpu(min(args.time_step_cpu_refinements,
                                                                      args.time_step_refinements))
            except Exception as e:
                line["time_step"]["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def side_measure(op, src, dst, steps, n_cells, n_global, bpc, peak):
    """3 warm-ups + `steps` timed vmults with CUDA events (whole step and cell kernel)"""
    import torch
    for _ in range(3):
        op.vmult(dst, src)
    torch.cuda.synchronize()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    s_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        s_ev[i][0].record()
        op.vmult(dst, src, kernel_events=k_ev[i])
        s_ev[i][1].record()
    e1.record()
    torch.cuda.synchronize()
    loop_ms = e0.elapsed_time(e1) / steps
    # device time of one whole vmult (memset, cell kernel, constrained rows); the loop figure next to it also holds
    # whatever the host needed between two steps in this late phase of the run (CPU-baseline threads still around)
    t_ms = float(np.mean([a.elapsed_time(b) for a, b in s_ev])) * steps
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    achieved = n_cells * bpc / (k_ms * 1e-3) / 1e9
    return {"value": n_global * steps / (t_ms * 1e-3) / 1e9, "unit": "GDoF/s", "steps": steps, "ms_per_step": t_ms / steps,
            "loop_ms_per_step": loop_ms, "kernel_ms": k_ms, "roofline_achieved_gbs": achieved, "roofline_frac": achieved / peak,
            "algorithmic_bytes_per_cell": bpc, "kernel_variant": op.vmult_variant()}


def config_c_measure(args, dev, peak):
    """bench.py --workload C as an extra key of the default line"""
    import torch

    from dealii_ns_gls_b200 import mesh as gm
    from dealii_ns_gls_b200.operator import NavierStokesOperator
    from dealii_ns_gls_b200.time_integration import TimeIntegratorDataBDF
    mesh = gm.cylinder_shell(shell_shape(args), args.degree)
    ti = TimeIntegratorDataBDF(2)
    ti.update_dt(DT)
    ti.update_dt(DT)
    op = NavierStokesOperator(mesh, None, 0.001, C1, C2, ti, True, True, False, number="double", device=dev)
    n, n_cells = mesh.n_dofs, mesh.n_cells
    del mesh
    g = torch.Generator(device=dev).manual_seed(SEED)
    rv = lambda: torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1  # noqa: E731
    op.set_previous_solution([rv() for _ in range(3)])
    op.set_linearization_point(rv())
    src, dst = rv(), op.initialize_dof_vector()
    bpc = algorithmic_bytes_per_cell(3, args.degree, 8, ctd=True, q_wise=True, general=True)
    r = side_measure(op, src, dst, min(args.steps, 10), n_cells, n, bpc, peak)
    cargs = argparse.Namespace(**dict(vars(args), workload="C"))
    r["workload"] = workload_config(cargs, 1)["workload"]
    r["n_dofs"] = n
    return r


def time_step_wall(refinements, dev, n_steps=4, n_ranks=1, rank=0):
    """The second half of BASELINE.json's metric, "wall time per time step": the time loop of main.cc:908-990
    (get_max_u, set_previous_solution on all levels, Newton with GMRES + geometric multigrid: relaxation
    smoothers, device transfers, dense coarse solve) on the 3-D channel of input_channel.json at Q2, with the
    fine operator in double and the level operators in float (config.h:6-7), everything resident on the device.
    Wall clock around whole steps, device synchronised on both sides; the first two steps are warm-up."""
    import torch

    from dealii_ns_gls_b200.driver import ChannelParameters, Driver
    t0 = time.perf_counter()
    # mg_min_level = 1: the coarsest level every rank of an 8-GPU run still holds cells on (the same hierarchy at
    # every N, so that the iteration counts can be compared across N)
    d = Driver(ChannelParameters(dim=3, fe_degree=2, n_global_refinements=refinements, mg_min_level=1), device=dev,
               n_ranks=n_ranks, rank=rank)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    def sync():
        torch.cuda.synchronize()
        if n_ranks > 1:
            import torch.distributed as dist
            dist.barrier()

    walls, recs = [], []
    for i in range(2 + n_steps):
        sync()
        t0 = time.perf_counter()
        rec = d.step()
        sync()
        if i >= 2:
            walls.append(time.perf_counter() - t0)
            recs.append(rec)
    fine = d.meshes[d.maxlevel]
    return {"wall_s_per_step": float(np.mean(walls)), "unit": "s", "steps": n_steps, "warmup_steps": 2,
            "n_gpus": n_ranks, "scaling": "strong (the same mesh cut into x-slabs of the channel)",
            "n_dofs": int(fine.n_global_dofs), "n_cells": int(fine.n_cells) * n_ranks, "levels": d.maxlevel + 1 - d.minlevel,
            "newton_iterations": [r["newton_iterations"] for r in recs],
            "gmres_iterations": [r["linear_iterations"] for r in recs],
            "setup_s": setup_s,
            "workload": f"3D channel (simulation.cc:143-189), Q2, {int(fine.n_cells) * n_ranks} cells, BDF1, CFL 0.1, Newton + "
                        "GMRES(rel 1e-2) + GMG V-cycle (5 relaxation sweeps, coarse direct), fine operator f64, "
                        "level operators f32"}


def time_step_cpu(refinements, n_steps=1, budget_s=20.0):
    """`time_step.cpu_baseline`: the same time loop (same channel, same flags, same hierarchy with mg_min_level = 1,
    float level vectors) on the host cores -- the CPU restatement of the solver stack (oracle/gls_solver.py: Newton,
    GMRES(28), V-cycle with 5-sweep relaxation smoothers and their 20-step power iteration, unit-vector diagonals,
    dense coarse solve) with every cell loop in the C restatement (oracle/gls_fast.py, all host threads) -- on a
    bounded sample: a coarser refinement of the mesh the device runs.  Two untimed steps, then up to n_steps timed
    steps inside budget_s.  `cell_loops_in_c_s` is the part of a step spent inside the C loops; the rest is numpy /
    scipy glue (table evaluation, transfers as sparse matrices) that a compiled CPU code would not pay."""
    from dealii_ns_gls_b200 import mesh as gm
    from dealii_ns_gls_b200.driver import ChannelParameters
    from oracle import gls_solver as gs
    from oracle.gls_fast import FastOracleOperator
    from oracle.gls_oracle_c import max_threads

    p = ChannelParameters(dim=3, fe_degree=2, n_global_refinements=refinements, mg_min_level=1)
    n_levels = p.n_levels()
    meshes = {l: p.level_mesh(l) for l in range(p.mg_min_level, n_levels + 1)}
    children = {l: gm.child_cells(meshes[l - 1], meshes[l]) for l in range(p.mg_min_level + 1, n_levels + 1)}
    fine = meshes[n_levels]
    ci = p.inhomogeneous_constraints(fine)
    t0 = time.perf_counter()
    d = gs.OracleChannelDriver(dim=p.dim, degree=p.fe_degree, meshes=meshes, children=children,
                               constraints_inhomogeneous=ci.rows, inhomogeneities=ci.inhomogeneities,
                               min_dx=p.minimal_cell_diameter(fine), nu=p.nu, c1=p.c_1, c2=p.c_2, cfl=p.cfl,
                               bdf_order=p.bdf_order, consider_time_derivative=p.consider_time_derivative,
                               cell_wise_stabilization=p.cell_wise_stabilization, rel_tol=p.lin_relative_tolerance,
                               abs_tol=p.lin_absolute_tolerance, newton_inexact=p.newton_inexact,
                               level_dtype=np.float32, operator_class=FastOracleOperator)
    setup_s = time.perf_counter() - t0
    for _ in range(2):      # like time_step_wall: the first two steps (impulsive start, more Newton steps) are warm-up
        d.step()
    walls, recs, c_s = [], [], []
    t_start = time.perf_counter()
    for _ in range(n_steps):
        c0, t0 = FastOracleOperator.c_seconds, time.perf_counter()
        recs.append(d.step())
        walls.append(time.perf_counter() - t0)
        c_s.append(FastOracleOperator.c_seconds - c0)
        if time.perf_counter() - t_start > budget_s:
            break
    wall = float(np.mean(walls))
    return {"wall_s_per_step": wall, "unit": "s", "steps": len(walls), "warmup_steps": 2, "kind": "port",
            "cores": max_threads(), "n_dofs": int(fine.n_dofs), "n_cells": int(fine.n_cells),
            "levels": n_levels + 1 - p.mg_min_level, "dofs_per_s": fine.n_dofs / wall,
            "cell_loops_in_c_s": float(np.mean(c_s)), "setup_s": setup_s,
            "newton_iterations": [r["newton_iterations"] for r in recs],
            "gmres_iterations": [r["linear_iterations"] for r in recs],
            "sample": f"the same 3-D Q2 channel time loop at n global refinements = {refinements} "
                      f"({int(fine.n_cells)} cells, {int(fine.n_dofs)} DoFs), CPU restatement of the solver stack "
                      "(oracle/gls_solver.py) with the cell loops in C (oracle/gls_fast.py), not deal.II"}


_REAL_STDOUT = None


def _protect_stdout():
    """Libraries (NCCL prints its version banner) must not write into the one-JSON-line stdout:
    point fd 1 at stderr for the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=160, help="cells per direction per GPU (160 -> 1.32e8 DoFs)")
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--cpu-cells", type=int, default=64, help="cells per direction of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="P", choices=["P", "C"],
                    help="P = performance.cc hypercube (the headline); C = curved O-grid with the Turek-3D flags")
    ap.add_argument("--number", default="double", choices=["double", "float"],
                    help="double = Krylov operator (headline); float = multigrid level operator (config.h:7)")
    ap.add_argument("--partition", default="morton", choices=["morton", "slab"],
                    help="N > 1: morton = the reference's p4est owner ranks (halves / quarters / octants); slab = z-slabs")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys (config C, zero vectors, parity)")
    ap.add_argument("--time-step-refinements", type=int, default=4,
                    help="n global refinements of the 3-D Q2 channel whose wall time per time step is reported "
                         "next to the vmult metric (3 -> 4.3e6 DoFs, 4 -> 3.4e7); -1 = skip")
    ap.add_argument("--time-step-cpu-refinements", type=int, default=2,
                    help="refinements of the bounded sample the CPU restatement of the time step runs on "
                         "(2 -> 5.6e5 DoFs, a few seconds per step)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
