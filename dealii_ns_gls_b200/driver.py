"""The wiring of the reference's Driver<dim>::run (main.cc:220-1000) around the device operator, for the
synthetic structured meshes of mesh.py: constraints (main.cc:258-310), fine operator (:326-348), the
global-coarsening hierarchy with level operators and the two transfer objects (:396-568), the solver hooks
(:772-869) and the time loop (:908-990).  Used by the tests (iteration counts against the CPU oracle) and by
``bench.py --workload step`` (wall time per time step); it is the *caller* of the hot path, kept as close to
the reference's control flow as the synthetic setting allows.  Only the channel simulation
(simulation.cc:143-189, input/input_channel.json) is described here.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .mesh import Mesh, dof_components, dof_coordinates, structured_box, structured_mesh
from .multigrid import (DeviceVectorOps, MGTransferGlobalCoarsening, MGTwoLevelTransfer, PreconditionerGMG,
                        PreconditionerGMGAdditionalData)
from .operator import AffineConstraints, NavierStokesOperator
from .solvers import LinearSolverGMRES, NonLinearSolverNewton
from .time_integration import SolutionHistory, TimeIntegratorDataBDF, TimeIntegratorDataNone


class ChannelParameters:
    """input/input_channel.json + the defaults of main.cc:66-192 that the device path reads."""

    def __init__(self, **kw):
        self.dim = 2
        self.fe_degree = 1
        self.n_global_refinements = 2
        self.cfl = 0.1
        self.dt = 0.0
        self.bdf_order = 1
        self.time_integration = "bdf"
        self.c_1, self.c_2, self.nu = 2.0, 1.0, 0.0
        self.consider_time_derivative = True
        self.cell_wise_stabilization = True
        self.lin_n_max_iterations = 10000
        self.lin_absolute_tolerance = 1e-12
        self.lin_relative_tolerance = 1e-2
        self.newton_inexact = False
        self.mg_number = "float"      # config.h:7
        self.n_stretching = 4         # simulation.cc:143-145
        self.mg_min_level = 0         # coarsest multigrid level ("mg min level"); a partitioned run needs every rank
                                      # to hold cells on it
        self.gmg = PreconditionerGMGAdditionalData()
        self.simulation_name = "channel"
        self._more_defaults()
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(self, k, v)

    def _more_defaults(self):
        pass

    # ---- what Driver needs from a simulation (SimulationBase of include/simulation.h) ----
    def n_levels(self):
        return 2 + self.n_global_refinements  # refine_global(2) + n (simulation.cc:166-169)

    def level_mesh(self, level, n_ranks=1, rank=0):
        return channel_level_mesh(self, level, n_ranks, rank)

    def inhomogeneous_constraints(self, mesh):
        return channel_inhomogeneous_constraints(self, mesh)

    def minimal_cell_diameter(self, fine):
        return math.sqrt(self.dim) / 2 ** self.n_levels()  # GridTools::minimal_cell_diameter of a cube


class CylinderParameters(ChannelParameters):
    """input/input_turek_2D_Re20_stat.json, input_turek_3D_Re100.json and input_hoffmann_3D_Re3900.json on the
    synthetic O-grid (mesh.cylinder_shell's map): Q2, q-point-wise delta, nu = 0.001, BDF2 or "none", inexact
    Newton, direct coarse solver; no-slip rows on the cylinder, u = (u_max, 0, 0) on the upstream half of the outer
    boundary (inhomogeneous, simulation.cc:379-431), p = 0 on the downstream half ("homogeneous nbc",
    main.cc:279-283), and on the two z-planes either no-slip or slip walls (w = 0: what
    compute_no_normal_flux_constraints gives on axis-aligned walls, main.cc:285-287)."""

    def _more_defaults(self):
        self.simulation_name = "cylinder"
        self.dim, self.fe_degree, self.n_global_refinements = 2, 2, 1
        self.cfl, self.bdf_order, self.time_integration = 1.0, 2, "bdf"
        self.c_1, self.c_2, self.nu = 2.0, 1.0, 0.001
        self.consider_time_derivative, self.cell_wise_stabilization = True, False
        self.u_max = 0.3
        self.no_slip_wall = False            # z-planes: False = slip walls (Hoffmann), True = no-slip (Turek)
        self.base_shape = (1, 4, 1)          # coarse level: radial x circumferential (x axial) cells
        self.r_inner, self.r_outer, self.length = 0.05, 0.5, 0.41

    def n_levels(self):
        return self.n_global_refinements

    def _deform(self, x):
        r = self.r_inner + (self.r_outer - self.r_inner) * x[..., 0] ** 1.5
        th = 2.0 * np.pi * x[..., 1]
        out = np.empty_like(x)
        out[..., 0], out[..., 1] = r * np.cos(th), r * np.sin(th)
        if self.dim == 3:
            out[..., 2] = self.length * x[..., 2]
        return out

    def _masks(self, ref):
        eps = 1e-12
        inner, outer = np.abs(ref[:, 0]) < eps, np.abs(ref[:, 0] - 1.0) < eps
        upstream = np.cos(2.0 * np.pi * ref[:, 1]) < -1e-9
        wall = np.zeros(len(ref), dtype=bool)
        if self.dim == 3:
            wall = (np.abs(ref[:, 2]) < eps) | (np.abs(ref[:, 2] - 1.0) < eps)
        return inner, outer & upstream, outer & ~upstream, wall

    def _zero_constrained(self, ref, c):
        inner, inflow, outflow, wall = self._masks(ref)
        if c == self.dim:
            return outflow                              # homogeneous nbc: pressure row
        m = inner | inflow
        if self.dim == 3:
            m |= wall if (self.no_slip_wall or c == 2) else np.zeros(len(ref), dtype=bool)
        return m

    def level_mesh(self, level):
        shape = tuple(s * 2 ** level for s in self.base_shape[:self.dim])
        periodic = (False, True) + ((False,) if self.dim == 3 else ())
        return structured_mesh(self.dim, shape, self.fe_degree, deform=self._deform, mapping_degree=self.fe_degree,
                               periodic=periodic, dirichlet=self._zero_constrained)

    def inhomogeneous_constraints(self, mesh):
        ref, comp = dof_coordinates(mesh), dof_components(mesh)
        inner, inflow, _, wall = self._masks(ref)
        rows, inhom = {}, {}
        for d in mesh.constraints:
            rows[d] = []
            if comp[d] == 0 and inflow[d] and not inner[d] and not (self.no_slip_wall and wall[d]):
                inhom[d] = self.u_max
        return AffineConstraints(rows, inhom)

    def minimal_cell_diameter(self, fine):
        from .mesh import cell_diameters
        return float(cell_diameters(fine).min())


def channel_level_mesh(params: ChannelParameters, level: int, n_ranks: int = 1, rank: int = 0) -> Mesh:
    """SimulationChannel::create_triangulation (simulation.cc:150-170): n_stretching x 1 (x 1) unit blocks,
    refined `level` times; boundary ids 0/1 = x faces, 2/3 = y, 4/5 = z (colorize = true).  The zero
    constraints are constraints_homogeneous of main.cc:258-305: no-slip walls (ids >= 2), the inflow face
    (id 0) for the velocity, pressure = 0 on the outflow face (id 1)."""
    dim, ns = params.dim, params.n_stretching
    shape = (ns * 2 ** level,) + (2 ** level,) * (dim - 1)
    extent = (float(ns),) + (1.0,) * (dim - 1)
    eps = 1e-12

    def zero_constrained(x, c):
        if c == dim:
            return np.abs(x[:, 0] - extent[0]) < eps
        m = np.abs(x[:, 0]) < eps
        for e in range(1, dim):
            m |= (np.abs(x[:, e]) < eps) | (np.abs(x[:, e] - 1.0) < eps)
        return m

    if n_ranks == 1:
        return structured_mesh(dim, shape, params.fe_degree, extent=extent, dirichlet=zero_constrained)
    # one x-slab of the channel per rank on every level (the partitions of consecutive levels nest: children live
    # on the rank of their parent, like the reference's global-coarsening hierarchy after repartitioning by
    # the same space-filling curve)
    if shape[0] % n_ranks != 0:
        raise ValueError(f"level {level}: {shape[0]} cells in x cannot be cut into {n_ranks} slabs; raise mg_min_level")
    grid = (n_ranks,) + (1,) * (dim - 1)
    box_of = np.array([(r,) + (0,) * (dim - 1) for r in range(n_ranks)], dtype=np.int64)
    local = (shape[0] // n_ranks,) + shape[1:]
    return structured_box(local, params.fe_degree, grid=grid, box_of=box_of, rank=rank,
                          box_extent=np.asarray(extent) / np.asarray(grid), dirichlet=zero_constrained)


def channel_inhomogeneous_constraints(params: ChannelParameters, mesh: Mesh) -> AffineConstraints:
    """constraints_inhomogeneous of the time loop (main.cc:877-895): the zero constraints of the walls and the
    outflow pressure, plus u = (1, 0, 0) on the inflow face where no wall constraint exists yet
    (InflowBoundaryValues::Channel(0, 1), simulation.cc:176-177)."""
    dim = params.dim
    cdofs = np.fromiter(mesh.constraints.keys(), dtype=np.int64, count=len(mesh.constraints))
    x, comp = dof_coordinates(mesh, only=cdofs), dof_components(mesh, only=cdofs)
    eps = 1e-12
    wall = np.zeros(len(cdofs), dtype=bool)
    for e in range(1, dim):
        wall |= (np.abs(x[:, e]) < eps) | (np.abs(x[:, e] - 1.0) < eps)
    inflow = (comp == 0) & (np.abs(x[:, 0]) < eps) & ~wall
    rows = {d: [] for d in mesh.constraints}
    inhom = {int(d): 1.0 for d in cdofs[inflow]}
    return AffineConstraints(rows, inhom)


class Driver:
    """Driver<dim>::run with "preconditioner": "GMG", "nonlinear solver": "Newton" for the simulations the
    parameter object describes (ChannelParameters, CylinderParameters)."""

    def __init__(self, params: ChannelParameters, device=None, verbose=False, n_ranks=1, rank=0, group=None):
        """n_ranks > 1: one process per GPU (torch.distributed initialised by the caller); every level is
        partitioned into the same boxes, level vectors are [owned | ghost] like LinearAlgebra::distributed::Vector,
        inner products are summed over the ranks (solver_l.cc:46-74, solver_nl.cc:50,76 through deal.II)."""
        self.params, self.verbose = params, verbose
        self.timers = None
        p = params
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_ranks, self.rank = n_ranks, rank
        n_levels = p.n_levels()
        self.minlevel, self.maxlevel = p.mg_min_level, n_levels
        if n_ranks == 1:
            self.meshes = {l: p.level_mesh(l) for l in range(self.minlevel, n_levels + 1)}
            make_exchange = lambda mesh: None  # noqa: E731
            DeviceVectorOps.allreduce_sum = None
        else:
            import torch.distributed as dist
            from .distributed import GhostExchange
            self.meshes = {l: p.level_mesh(l, n_ranks, rank) for l in range(self.minlevel, n_levels + 1)}
            make_exchange = lambda mesh: GhostExchange(mesh.partition, self.device, group)  # noqa: E731
            DeviceVectorOps.allreduce_sum = staticmethod(lambda t: dist.all_reduce(t, group=group))
        fine = self.meshes[self.maxlevel]
        if p.time_integration == "bdf":
            self.time_integrator_data = TimeIntegratorDataBDF(p.bdf_order)
        elif p.time_integration == "none":
            self.time_integrator_data = TimeIntegratorDataNone()
        else:
            raise NotImplementedError(p.time_integration)
        tid = self.time_integrator_data
        ci = p.inhomogeneous_constraints(fine)
        # vector-level constraint objects act on the owned block only (ghost entries stay zero between calls)
        own = fine.n_owned
        self.constraints_inhomogeneous = AffineConstraints({d: r for d, r in ci.rows.items() if d < own},
                                                           {d: v for d, v in ci.inhomogeneities.items() if d < own})
        # `constraints` of main.cc:268-306: everything but the rows of the inhomogeneous boundary ids
        self.constraints = AffineConstraints({d: r for d, r in self.constraints_inhomogeneous.rows.items()
                                              if d not in self.constraints_inhomogeneous.inhomogeneities})
        increment_form = True  # Newton (main.cc:331)
        # main.cc:333-348
        self.ns_operator = NavierStokesOperator(fine, self.constraints_inhomogeneous, p.nu, p.c_1, p.c_2, tid,
                                                p.consider_time_derivative, increment_form,
                                                p.cell_wise_stabilization, number="double", device=self.device,
                                                exchange=make_exchange(fine))
        # main.cc:396-568: level operators (MGNumber) and transfers
        self.mg_ns_operators = {}
        for l in range(self.minlevel, self.maxlevel + 1):
            self.mg_ns_operators[l] = NavierStokesOperator(self.meshes[l], None, p.nu, p.c_1, p.c_2, tid,
                                                           p.consider_time_derivative, increment_form,
                                                           p.cell_wise_stabilization, number=p.mg_number,
                                                           device=self.device, exchange=make_exchange(self.meshes[l]))
        t_nc, t_c = {}, {}
        for l in range(self.minlevel + 1, self.maxlevel + 1):
            mf, mc = self.meshes[l], self.meshes[l - 1]
            of, oc = self.mg_ns_operators[l], self.mg_ns_operators[l - 1]
            t_nc[l] = MGTwoLevelTransfer().reinit(mf, mc, None, None, number=p.mg_number, device=self.device,
                                                  op_fine=of, op_coarse=oc)
            t_c[l] = MGTwoLevelTransfer().reinit(mf, mc, AffineConstraints(mf.constraints),
                                                 AffineConstraints(mc.constraints), number=p.mg_number,
                                                 device=self.device, op_fine=of, op_coarse=oc)
        init = lambda l: self.mg_ns_operators[l].initialize_dof_vector()  # noqa: E731
        self.mg_transfer_no_constraints = MGTransferGlobalCoarsening(t_nc, init)
        self.transfer = MGTransferGlobalCoarsening(t_c, init)
        self.preconditioner = PreconditionerGMG(self.mg_ns_operators, self.transfer, p.gmg)
        self.linear_solver = LinearSolverGMRES(self.ns_operator, self.preconditioner, p.lin_n_max_iterations,
                                               p.lin_absolute_tolerance, p.lin_relative_tolerance)
        self.nonlinear_solver = NonLinearSolverNewton(p.newton_inexact)
        self._ops = DeviceVectorOps()
        self._hom_idx = torch.tensor(sorted(fine.constraints.keys()), dtype=torch.int32, device=self.device)
        self._wire_hooks()
        self.solution = SolutionHistory(tid.get_order() + 1)
        self.solution.solutions = [self.ns_operator.initialize_dof_vector() for _ in range(tid.get_order() + 1)]
        self.constraints_inhomogeneous.distribute(self.solution.get_current_solution())
        self.t, self.counter = 0.0, 1
        self.min_dx = p.minimal_cell_diameter(fine)
        self.log = []

    # ---- main.cc:772-869 ------------------------------------------------------------------------------
    def _wire_hooks(self):
        nl, tid = self.nonlinear_solver, self.time_integrator_data

        def setup_jacobian(src):
            self.ns_operator.set_linearization_point(src)

        def setup_preconditioner(solution):
            mg_solution = {}
            self.mg_transfer_no_constraints.interpolate_to_mg(mg_solution, solution)
            for l, op in self.mg_ns_operators.items():
                op.set_linearization_point(mg_solution[l])
            self.preconditioner.initialize()
            self.linear_solver.initialize()

        def evaluate_rhs(dst):
            self.ns_operator.evaluate_rhs(dst)

        def evaluate_residual(dst, src):
            self.ns_operator.evaluate_residual(dst, src)

        def solve_with_jacobian(dst, src):
            self._ops.set_zero_indexed(src, self._hom_idx)  # constraints_homogeneous.set_zero(src)
            self.linear_solver.solve(dst, src)
            self.ns_operator.get_constraints().distribute(dst)

        t = self._timed  # the reference wraps every hook in a MyScope timer (main.cc:806-857)
        nl.setup_jacobian = t("setup_jacobian", setup_jacobian)
        nl.setup_preconditioner = t("setup_preconditioner", setup_preconditioner)
        nl.evaluate_rhs = t("evaluate_rhs", evaluate_rhs)
        nl.evaluate_residual = t("evaluate_residual", evaluate_residual)
        nl.solve_with_jacobian = t("solve_with_jacobian", solve_with_jacobian)

    def _timed(self, name, fn):
        """wall-clock scopes like the reference's TimerCollection; only when self.timers is a dict (each scope
        then synchronises the device, so leave it None for throughput runs)"""
        def wrapped(*a):
            if self.timers is None:
                return fn(*a)
            import time
            torch.cuda.synchronize(self.device)
            t0 = time.perf_counter()
            r = fn(*a)
            torch.cuda.synchronize(self.device)
            self.timers[name] = self.timers.get(name, 0.0) + time.perf_counter() - t0
            return r
        return wrapped

    def set_previous_solution(self, solution: SolutionHistory):
        """main.cc:772-803"""
        tid = self.time_integrator_data
        self.ns_operator.set_previous_solution(solution)
        order = tid.get_order()
        if order == 0:
            return
        all_mg = {l: [None] * (order + 1) for l in self.mg_ns_operators}
        for i in range(1, order + 1):
            mg_solution = {}
            self.mg_transfer_no_constraints.interpolate_to_mg(mg_solution, solution.get_vectors()[i])
            for l in self.mg_ns_operators:
                all_mg[l][i] = mg_solution[l]
        for l, op in self.mg_ns_operators.items():
            all_mg[l][0] = all_mg[l][1]  # slot 0 (the current solution) is not read (operator_ns.cc:246-258)
            op.set_previous_solution(all_mg[l])

    def step(self):
        """one pass of the time loop body, main.cc:908-978"""
        p, tid = self.params, self.time_integrator_data
        cur = self.solution.get_current_solution()
        u_max = self.ns_operator.get_max_u(cur)
        dt = p.dt if p.dt != 0.0 else self.min_dx * p.cfl / max(u_max, 1.0)
        tid.update_dt(dt)
        self.ns_operator.invalidate_system()
        for op in self.mg_ns_operators.values():
            op.invalidate_system()
        self.solution.commit_solution()
        self.set_previous_solution(self.solution)
        cur = self.solution.get_current_solution()
        n_lin_before = len(self.linear_solver.n_iterations)
        n_newton = self.nonlinear_solver.solve(cur)
        self.constraints_inhomogeneous.distribute(cur)
        self.constraints.distribute(cur)  # main.cc:963
        self.t += dt
        rec = dict(cycle=self.counter, t=self.t, dt=dt, u_max=u_max, newton_iterations=n_newton,
                   newton_residuals=list(self.nonlinear_solver.residuals),
                   linear_iterations=list(self.linear_solver.n_iterations[n_lin_before:]))
        self.log.append(rec)
        if self.verbose and self.rank == 0:
            # the reference's console lines (main.cc:921-923, solver_nl.cc:53-88, solver_l.cc:70, main.cc:971), so that
            # a deal.II run of the reference and this run can be diffed (reflog.py)
            from .reflog import format_step
            l2 = float(torch.linalg.vector_norm(cur[:self.ns_operator.n_owned])) if self.n_ranks == 1 else None
            print(format_step(dict(rec, t=rec["t"] - dt, solution_l2=l2)), flush=True)
        self.counter += 1
        return rec
