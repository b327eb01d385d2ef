"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the
CPU oracle on the same seeded inputs.  Tolerances: 1e-12 relative l2 in FP64 (the
north-star bar), 2e-5 in FP32 (multigrid level operators, config.h:7)."""
import numpy as np
import pytest

from dealii_ns_gls_b200 import mesh as gm
from tests.util import TI, make_gpu, make_oracle, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-12, "float": 2e-5}
NPDT = {"double": np.float64, "float": np.float32}


def _torch():
    import torch
    return torch


def _to_dev(a, number):
    torch = _torch()
    return torch.tensor(np.asarray(a), dtype=torch.float64 if number == "double" else torch.float32, device="cuda")


def torch_idx(a):
    return _torch().tensor(np.asarray(a, dtype=np.int64), device="cuda")


def _mesh(kind, dim, degree, **kw):
    if kind == "cube":
        return gm.hypercube(dim, 4 if dim == 2 else 3, degree, **kw)
    shape = (3, 6) if dim == 2 else (2, 5, 2)
    return gm.cylinder_shell(shape, degree, **kw)


def _setup(mesh, ti, number, seed=1234, **flags):
    rng = np.random.default_rng(seed)
    ora = make_oracle(mesh, ti, dtype=NPDT[number], **flags)
    gpu = make_gpu(mesh, ti, number=number, **flags)
    hist = [rng.uniform(-1, 1, mesh.n_dofs) for _ in range(ti.get_order() + 1)]
    lin = rng.uniform(-1, 1, mesh.n_dofs)
    src = rng.uniform(-1, 1, mesh.n_dofs)
    if ti.get_order() > 0:
        ora.set_previous_solution(hist, ti.get_weights())
        gpu.set_previous_solution([_to_dev(h, number) for h in hist])
    ora.set_linearization_point(lin, ti.get_current_dt())
    gpu.set_linearization_point(_to_dev(lin, number))
    return ora, gpu, src, rng


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("kind", ["cube", "shell"])
@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (2, 3), (2, 4), (3, 1), (3, 2), (3, 3), (3, 4)])
def test_vmult_newton(dim, degree, kind, number):
    """performance.cc flags (cell-wise delta, no time derivative) on random vectors."""
    mesh = _mesh(kind, dim, degree)
    ti = TI(2, [10.0, -10.0, 0.0], 0.1)
    ora, gpu, src, _ = _setup(mesh, ti, number)
    ref = ora.vmult(src, 10.0)
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, _to_dev(src, number))
    assert rel_l2(dst.cpu().numpy(), ref, mesh=mesh) < TOL[number]


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("dim,degree", [(2, 2), (3, 2), (3, 3)])
def test_vmult_newton_turek_flags(dim, degree, number):
    """input_turek_3D_Re100.json flags: BDF2, time derivative in the stabilization,
    q-point-wise delta, curved cells, no-slip rows."""
    mesh = _mesh("shell", dim, degree)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    ora, gpu, src, _ = _setup(mesh, ti, number, ctd=True, cell_wise=False, nu=0.001)
    ref = ora.vmult(src, 15.0)
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, _to_dev(src, number))
    assert rel_l2(dst.cpu().numpy(), ref, mesh=mesh) < TOL[number]
    cons = np.array(sorted(mesh.constraints.keys()))
    assert np.array_equal(dst.cpu().numpy()[cons], src.astype(NPDT[number])[cons])  # identity rows, bit-exact


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("name", ["u_star_value", "u_star_gradient", "p_star_gradient", "u_time_derivative_old",
                                  "delta_1", "delta_2", "delta_1_q", "delta_2_q"])
def test_tables(name, number):
    mesh = _mesh("shell", 3, 2)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    ora, gpu, _, _ = _setup(mesh, ti, number, ctd=True, cell_wise=False, nu=0.001)
    d = mesh.dim
    got = gpu.get_table(name).cpu().numpy()
    K = mesh.n_cells
    ref = {"u_star_value": lambda: ora.U.transpose(1, 0, 2),
           "u_star_gradient": lambda: ora.H.reshape(K, d * d, -1).transpose(1, 0, 2),
           "p_star_gradient": lambda: ora.P.transpose(1, 0, 2),
           "u_time_derivative_old": lambda: ora.o.transpose(1, 0, 2),
           "delta_1": lambda: ora.delta1_cell.reshape(1, K, 1),
           "delta_2": lambda: ora.delta2_cell.reshape(1, K, 1),
           "delta_1_q": lambda: ora.delta1_q.reshape(1, K, -1),
           "delta_2_q": lambda: ora.delta2_q.reshape(1, K, -1)}[name]()
    assert rel_l2(got, ref) < TOL[number]


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("dim,degree", [(2, 1), (3, 2)])
@pytest.mark.parametrize("theta", [1.0, 0.5])
def test_fixed_point_and_residual(dim, degree, theta, number):
    """!increment_form vmult and evaluate_residual (operator_ns.cc:955-1066), BDF and theta schemes."""
    mesh = _mesh("shell", dim, degree)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1) if theta == 1.0 else TI(1, [10.0, -10.0], 0.1, theta=theta)
    ora, gpu, src, _ = _setup(mesh, ti, number, ctd=(theta == 1.0), cell_wise=False, increment_form=False)
    ref = ora.vmult(src, ti.get_primary_weight())
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, _to_dev(src, number))
    assert rel_l2(dst.cpu().numpy(), ref, mesh=mesh) < TOL[number]
    # residual: src with boundary values already distributed (zero-type rows: src entries = 0)
    sb = src.copy()
    sb[list(mesh.constraints.keys())] = 0.0
    ref_r = ora.evaluate_residual(sb, ti.get_primary_weight())
    gpu.evaluate_residual(dst, _to_dev(sb, number))
    assert rel_l2(dst.cpu().numpy(), ref_r, mesh=mesh) < TOL[number]


@pytest.mark.parametrize("number", ["double", "float"])
def test_residual_increment_form(number):
    """Newton configuration: vmult = linearized branch, residual = fixed-point branch (Appendix A)."""
    mesh = _mesh("shell", 3, 2)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    ora, gpu, src, _ = _setup(mesh, ti, number, ctd=True, cell_wise=False, increment_form=True)
    ref_r = ora.evaluate_residual(src, 15.0)
    dst = gpu.initialize_dof_vector()
    gpu.evaluate_residual(dst, _to_dev(src, number))
    assert rel_l2(dst.cpu().numpy(), ref_r, mesh=mesh) < TOL[number]


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("dim,degree,kind", [(2, 2, "cube"), (3, 2, "shell"), (3, 1, "cube")])
def test_weighted_constraints_and_component_numbering(dim, degree, kind, number):
    """General affine rows (hanging-node-like) + component-major numbering (general index path)."""
    mesh = _mesh(kind, dim, degree, numbering="component")
    gm.add_random_constraints(mesh, n_weighted=12, n_zero=7, seed=11)
    ti = TI(2, [10.0, -10.0, 0.0], 0.1)
    ora, gpu, src, _ = _setup(mesh, ti, number)
    ref = ora.vmult(src, 10.0)
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, _to_dev(src, number))
    assert rel_l2(dst.cpu().numpy(), ref, mesh=mesh) < TOL[number]
    ref_r = ora.evaluate_residual(src, 10.0)
    gpu.evaluate_residual(dst, _to_dev(src, number))
    assert rel_l2(dst.cpu().numpy(), ref_r, mesh=mesh) < TOL[number]


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("dim,degree,kind,constrained", [(2, 1, "cube", False), (2, 3, "shell", True),
                                                         (3, 2, "cube", True), (3, 2, "shell", False)])
def test_inverse_diagonal(dim, degree, kind, constrained, number):
    mesh = _mesh(kind, dim, degree)
    if constrained:
        gm.add_random_constraints(mesh, n_weighted=8, n_zero=5, seed=5)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    ora, gpu, _, _ = _setup(mesh, ti, number, ctd=True, cell_wise=False)
    ref = ora.compute_inverse_diagonal(15.0)
    diag = gpu.initialize_dof_vector()
    gpu.compute_inverse_diagonal(diag)
    assert rel_l2(diag.cpu().numpy(), ref, mesh=mesh) < (1e-11 if number == "double" else 5e-5)


@pytest.mark.parametrize("number", ["double", "float"])
def test_get_max_u(number):
    mesh = _mesh("shell", 3, 2)
    ti = TI(0, [], 1.0)
    ora, gpu, src, _ = _setup(mesh, ti, number)
    got = gpu.get_max_u(_to_dev(src, number))
    assert abs(got - ora.get_max_u(src)) < TOL[number] * 10


def test_stationary_configuration():
    """'time intration: none' => order 0, weight 0, dt 1 (time_integration.cc:141-178)."""
    mesh = _mesh("shell", 2, 2)
    ti = TI(0, [], 1.0)
    ora, gpu, src, _ = _setup(mesh, ti, "double", ctd=True, cell_wise=False)
    ref = ora.vmult(src, 0.0)
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, _to_dev(src, "double"))
    assert rel_l2(dst.cpu().numpy(), ref, mesh=mesh) < 1e-12


def test_errors_are_loud():
    from dealii_ns_gls_b200._lib import GlsbError
    mesh = _mesh("cube", 2, 1)
    gpu = make_gpu(mesh, TI(2, [10.0, -10.0, 0.0], 0.1))
    v = gpu.initialize_dof_vector()
    with pytest.raises(GlsbError, match="set_linearization_point"):
        gpu.vmult(v, v.clone())
    with pytest.raises(ValueError):
        gpu.vmult(v[:-1], v)


def test_vmult_linearity_and_repeatability_large():
    """Size-independent properties at a size the oracle does not reach: linearity and
    zero -> zero (the reference's own benchmark vectors, performance.cc:66-79)."""
    torch = _torch()
    mesh = gm.hypercube(3, 24, 2)
    ti = TI(2, [10.0, -10.0, 0.0], 0.1)
    gpu = make_gpu(mesh, ti)
    g = torch.Generator(device="cuda").manual_seed(1234)
    lin = torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    gpu.set_linearization_point(lin)
    x = torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    y = torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    ax, ay, axy, z = (gpu.initialize_dof_vector() for _ in range(4))
    gpu.vmult(ax, x)
    gpu.vmult(ay, y)
    gpu.vmult(axy, 2.0 * x - 3.0 * y)
    assert float(torch.linalg.norm(axy - (2.0 * ax - 3.0 * ay)) / torch.linalg.norm(axy)) < 1e-13
    gpu.vmult(z, torch.zeros_like(x))
    assert float(z.abs().max()) == 0.0
    ax2 = gpu.initialize_dof_vector()
    gpu.vmult(ax2, x)
    assert float(torch.linalg.norm(ax2 - ax) / torch.linalg.norm(ax)) < 1e-14


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("kind,ctd,cell_wise", [("cube", False, True), ("shell", True, False), ("shell", False, True),
                                                ("cube", True, False)])
def test_q2_fast_kernel_matches_generic_and_oracle(kind, ctd, cell_wise, number):
    """The register-tiled Q2 kernel (3-D, degree 2, Newton branch) is the variant vmult picks; it
    must agree with the oracle and with the generic kernel, on meshes whose cell count is not a
    multiple of the 32-cell batch and with Dirichlet rows."""
    mesh = gm.hypercube(3, 5, 2) if kind == "cube" else gm.cylinder_shell((3, 7, 3), 2)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    ora, gpu, src, _ = _setup(mesh, ti, number, ctd=ctd, cell_wise=cell_wise, nu=0.01)
    ref = ora.vmult(src, 15.0)
    x = _to_dev(src, number)
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, x)
    assert gpu.vmult_variant() == "q2_regtile_tma"
    assert rel_l2(dst.cpu().numpy(), ref, mesh=mesh) < TOL[number]
    gpu.set_variant(1)
    dst2 = gpu.initialize_dof_vector()
    gpu.vmult(dst2, x)
    assert gpu.vmult_variant() == "generic"
    assert rel_l2(dst2.cpu().numpy(), ref, mesh=mesh) < TOL[number]
    assert rel_l2(dst.cpu().numpy(), dst2.cpu().numpy(), mesh=mesh) < TOL[number]


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("case", ["cube_24_chunked", "cube_13_chunked", "cube_dirichlet_chunked", "shell_small"])
def test_vmult_host_matches_device_vmult(case, number, monkeypatch):
    """glsb_vmult_host (host vectors in, host vector out; chunked upload / cells / download pipeline)
    must give bit-for-bit the order-independent part of the device result and agree to round-off
    (atomics reorder the sums), with and without constrained rows, pinned and pageable buffers."""
    torch = _torch()
    monkeypatch.setenv("GLSB_HOST_CHUNKS", "7")
    if case == "cube_24_chunked":
        mesh = gm.hypercube(3, 24, 2)
    elif case == "cube_13_chunked":
        mesh = gm.hypercube(3, 13, 2)  # 2 197 cells: the last 32-cell batch is partly padding
    elif case == "cube_dirichlet_chunked":
        def walls(ref, c):  # no-slip on all faces of the cube, pressure free
            on = (np.abs(ref) < 1e-12).any(axis=1) | (np.abs(ref - 1.0) < 1e-12).any(axis=1)
            return on if c < 3 else np.zeros(len(ref), dtype=bool)
        mesh = gm.hypercube(3, 20, 2, dirichlet=walls)
    else:
        mesh = gm.cylinder_shell((2, 5, 2), 2)
    ti = TI(2, [10.0, -10.0, 0.0], 0.1)
    gpu = make_gpu(mesh, ti, number=number)
    if case == "cube_dirichlet_chunked":
        gm.add_random_constraints  # noqa: B018  (weighted rows are covered by the device tests)
    dt = torch.float64 if number == "double" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(7)
    lin = (torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1).to(dt)
    gpu.set_linearization_point(lin)
    x = (torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1).to(dt)
    ref = gpu.initialize_dof_vector()
    gpu.vmult(ref, x)
    for pinned in (True, False):
        h_src = torch.empty(mesh.n_dofs, dtype=dt, pin_memory=pinned)
        h_dst = torch.full((mesh.n_dofs,), float("nan"), dtype=dt).pin_memory() if pinned else \
            torch.full((mesh.n_dofs,), float("nan"), dtype=dt)
        h_src.copy_(x)
        gpu.vmult_host(h_dst, h_src)
        torch.cuda.synchronize()
        assert not bool(torch.isnan(h_dst).any())
        assert rel_l2(h_dst.numpy(), ref.cpu().numpy()) < (1e-13 if number == "double" else 1e-5)


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("dim,degree,kind", [(2, 2, "cube"), (3, 2, "shell"), (3, 1, "cube")])
def test_gmg_ls_edge_indices_and_interface_operators(dim, degree, kind, number):
    """GMG-LS level operators (main.cc:569-732, input/rotation.json): vmult with edge-constrained
    indices (operator_ns.cc:692-700, :724-731) and vmult_interface_down / up (:734-787)."""
    mesh = _mesh(kind, dim, degree)
    rng = np.random.default_rng(99)
    free = np.setdiff1d(np.arange(mesh.n_owned), np.array(sorted(mesh.constraints.keys()), dtype=np.int64))
    edge = np.sort(rng.choice(free, size=max(3, len(free) // 9), replace=False))
    mesh.edge_constrained_indices = edge
    mesh.has_edge_constrained_indices = True
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    ora, gpu, src, _ = _setup(mesh, ti, number, ctd=True, cell_wise=False, nu=0.01)
    x = _to_dev(src, number)
    x0 = x.clone()
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, x)
    assert bool((x == x0).all()), "src must come back unchanged"
    assert rel_l2(dst.cpu().numpy(), ora.vmult(src, 15.0, edge), mesh=mesh) < TOL[number]
    gpu.vmult_interface_down(dst, x)
    assert rel_l2(dst.cpu().numpy(), ora.vmult_interface_down(src, 15.0), mesh=mesh) < TOL[number]
    gpu.vmult_interface_up(dst, x)
    assert rel_l2(dst.cpu().numpy(), ora.vmult_interface_up(src, 15.0, edge), mesh=mesh) < TOL[number]
    # inverse diagonal: exactly 1 on the refinement-edge dofs (operator_ns.cc:219-224)
    gpu.compute_inverse_diagonal(dst)
    ref_d = ora.compute_inverse_diagonal(15.0, edge)
    assert bool((dst[torch_idx(edge)] == 1).all()) and np.all(ref_d[edge] == 1)
    assert rel_l2(dst.cpu().numpy(), ref_d, mesh=mesh) < (1e-11 if number == "double" else 5e-5)
    # no edge indices anywhere: interface_up is the zero operator
    mesh.edge_constrained_indices = np.zeros(0, dtype=np.int64)
    mesh.has_edge_constrained_indices = False
    gpu2 = make_gpu(mesh, ti, number=number, ctd=True, cell_wise=False, nu=0.01)
    gpu2.set_previous_solution([x, x, x])
    gpu2.set_linearization_point(x)
    gpu2.vmult_interface_up(dst, x)
    assert float(dst.abs().max()) == 0.0


@pytest.mark.parametrize("number", ["float", "double"])
@pytest.mark.parametrize("dim,degree,kind", [(3, 2, "shell"), (2, 2, "cube"), (3, 1, "cube")])
def test_relaxation_smoother_on_device(dim, degree, kind, number):
    """PreconditionRelaxation as PreconditionerGMG configures it (multigrid.cc:290-304): omega from the
    20-step power iteration, 5 sweeps of vmult (zero guess) and of step, all on the device."""
    from dealii_ns_gls_b200.smoother import PreconditionRelaxation
    from oracle.gls_smoother import OracleRelaxation
    mesh = _mesh(kind, dim, degree)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    ora, gpu, src, rng = _setup(mesh, ti, number, ctd=True, cell_wise=False, nu=0.01)
    inv_diag = gpu.initialize_dof_vector()
    gpu.compute_inverse_diagonal(inv_diag)
    sm = PreconditionRelaxation(gpu, inv_diag)
    ref = OracleRelaxation(ora, 15.0, ora.compute_inverse_diagonal(15.0))
    tol = 5e-4 if number == "float" else 1e-10  # 20 normalised power iterations amplify round-off
    assert abs(sm.get_relaxation() - ref.get_relaxation()) < tol * ref.get_relaxation()
    assert abs(sm.estimate_eigenvalues().max_eigenvalue_estimate - ref.max_eigenvalue_estimate) \
        < tol * ref.max_eigenvalue_estimate
    # same omega on both sides for the sweep comparison
    ref.relaxation = sm.get_relaxation()
    b = _to_dev(src, number)
    x = gpu.initialize_dof_vector()
    sm.vmult(x, b)
    x_ref = ref.vmult(src)
    assert rel_l2(x.cpu().numpy(), x_ref, mesh=mesh) < 20 * TOL[number]
    sm.step(x, b)
    assert rel_l2(x.cpu().numpy(), ref.step(x_ref, src), mesh=mesh) < 20 * TOL[number]
    l0 = gpu.launch_count()
    sm.vmult(x, b)
    assert gpu.launch_count() - l0 >= 2 * sm.n_iterations - 1  # 4 x (cells + update) + first sweep


@pytest.mark.parametrize("kind,ctd,cell_wise", [("cube", False, True), ("shell", True, False)])
def test_q2_packed_float_kernel(kind, ctd, cell_wise, monkeypatch):
    """Opt-in packed FP32 variant of the Q2 kernel (two cells per lane, FFMA2; GLSB_Q2_PACK=1): same
    result as the oracle, odd and even numbers of 32-cell batches, with Dirichlet rows."""
    monkeypatch.setenv("GLSB_Q2_PACK", "1")
    for shape in ((5, 5, 5), (4, 4, 6)) if kind == "cube" else ((3, 7, 3), (2, 8, 2)):
        mesh = gm.hypercube(3, shape, 2) if kind == "cube" else gm.cylinder_shell(shape, 2)
        ti = TI(2, [15.0, -20.0, 5.0], 0.1)
        ora, gpu, src, _ = _setup(mesh, ti, "float", ctd=ctd, cell_wise=cell_wise, nu=0.01)
        dst = gpu.initialize_dof_vector()
        gpu.vmult(dst, _to_dev(src, "float"))
        assert gpu.vmult_variant() == "q2_regtile_tma"
        assert rel_l2(dst.cpu().numpy(), ora.vmult(src, 15.0), mesh=mesh) < TOL["float"]


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("workload,cells", [("P", 64), ("C", 32)])
def test_bench_sample_against_c_oracle(workload, cells, number):
    """The CUDA path against the C restatement at the size bench.py's CPU arm runs (64^3 cells = 8.6e6 DoFs for
    config P: 8 192 batches, i.e. ~28 batches per persistent CTA with ring refills and index-block wraparound;
    a 32 768-cell O-grid with no-slip rows for config C) -- the `parity` object of the bench line."""
    import bench
    torch = _torch()
    sm = bench.cpu_sample(cells, 2, workload=workload)
    r = bench.gpu_parity_on_sample(sm, number, torch.device("cuda", 0))
    assert r["kernel_variant"] == "q2_regtile_tma"
    assert r["identity_rows_bit_equal"]
    assert r["rel_l2_unconstrained_rows"] < TOL[number], r


@pytest.mark.parametrize("n_ranks,rank,n", [(2, 1, 24), (2, 1, 9), (4, 3, 24), (8, 7, 24), (8, 5, 24), (8, 0, 24), (4, 3, 8)])
def test_partitioned_vmult_host_matches_device_vmult_on_one_gpu(n_ranks, rank, n):
    """The host-vector pipeline of a PARTITIONED operator (glsb_vmult_host_begin / _finish around the ghost
    exchange) on the Morton boxes with 1, 3 and 7 owners of the ghost block, emulated on one GPU: the exchange is
    replaced by a loop-back that fills the ghost block with fixed values and drops the ghost contributions, for
    the device-vector path and the host-vector path alike; both must give the same owned entries."""
    torch = _torch()
    from dealii_ns_gls_b200.distributed import GhostExchange

    class LoopbackExchange(GhostExchange):
        def update_ghost_values(self, op, vec):
            vec[self.n_owned:] = self.ghost_values.to(vec.dtype)

        def compress_add(self, op, vec):
            vec[self.n_owned:] = 0

    mesh = gm.hypercube_box(n, 2, n_ranks=n_ranks, rank=rank, with_points=False)
    ex = LoopbackExchange(mesh.partition, torch.device("cuda", 0))
    g = torch.Generator(device="cuda").manual_seed(3 + rank)
    ex.ghost_values = torch.rand(mesh.n_dofs - mesh.n_owned, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    ti = TI(2, [10.0, -10.0, 0.0], 0.1)
    gpu = make_gpu(mesh, ti, exchange=ex)
    lin = torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    src = torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    lin[mesh.n_owned:] = 0
    src[mesh.n_owned:] = 0
    gpu.set_linearization_point(lin)
    ref = gpu.initialize_dof_vector()
    gpu.vmult(ref, src)
    h_src = torch.empty(mesh.n_dofs, dtype=torch.float64, pin_memory=True)
    h_dst = torch.full((mesh.n_dofs,), float("nan"), dtype=torch.float64).pin_memory()
    h_src.copy_(src)
    for _ in range(2):
        gpu.vmult_host(h_dst, h_src)
    torch.cuda.synchronize()
    no = mesh.n_owned
    assert not bool(torch.isnan(h_dst[:no]).any())
    err = float((h_dst[:no] - ref[:no].cpu()).abs().max() / ref.abs().max())
    assert err < 1e-13, err


@pytest.mark.parametrize("kind", ["cube_24", "shell", "hanging"])
def test_deterministic_mode_is_bit_reproducible(kind, monkeypatch):
    """GLSB_DETERMINISTIC=1: cells coloured over the vector entries they scatter to and run colour by colour, so every
    entry is summed in a fixed order (the north-star's "bit-for-bit identical iteration counts" needs a vmult that
    does not depend on the timing of atomics).  vmult, residual and inverse diagonal: five runs bit-equal, and equal
    to the oracle / the default mode to round-off."""
    torch = _torch()
    monkeypatch.setenv("GLSB_DETERMINISTIC", "1")
    if kind == "cube_24":
        mesh = gm.hypercube(3, 24, 2)
    elif kind == "shell":
        mesh = gm.cylinder_shell((4, 12, 4), 2)
    else:
        mesh = gm.hypercube_hanging(3, 4, 2)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    gpu = make_gpu(mesh, ti, ctd=True, cell_wise=False, nu=0.01)
    monkeypatch.delenv("GLSB_DETERMINISTIC")
    ref_op = make_gpu(mesh, ti, ctd=True, cell_wise=False, nu=0.01)
    g = torch.Generator(device="cuda").manual_seed(5)
    vec = lambda: torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1  # noqa: E731
    hist, lin, src = [vec() for _ in range(3)], vec(), vec()
    for op in (gpu, ref_op):
        op.set_previous_solution(hist)
        op.set_linearization_point(lin)
    outs = {"vmult": [], "residual": [], "diag": []}
    for _ in range(5):
        for name, fn in (("vmult", lambda d: gpu.vmult(d, src)), ("residual", lambda d: gpu.evaluate_residual(d, src)),
                         ("diag", lambda d: gpu.compute_inverse_diagonal(d))):
            d = gpu.initialize_dof_vector()
            fn(d)
            outs[name].append(d)
    for name, lst in outs.items():
        if name != "vmult" and kind == "hanging":
            # weighted (hanging-node) rows: in the generic kernels (residual, diagonal) the 27 point-threads of ONE
            # cell add to common masters concurrently; only the register-tiled vmult (one lane per component walks
            # its nodes in order) is bit-reproducible on such cells
            continue
        for d in lst[1:]:
            assert torch.equal(d, lst[0]), name
    d0 = ref_op.initialize_dof_vector()
    ref_op.vmult(d0, src)
    assert rel_l2(outs["vmult"][0].cpu().numpy(), d0.cpu().numpy(), mesh=mesh) < 1e-13
    ref_op.evaluate_residual(d0, src)
    assert rel_l2(outs["residual"][0].cpu().numpy(), d0.cpu().numpy(), mesh=mesh) < 1e-13
    ref_op.compute_inverse_diagonal(d0)
    assert rel_l2(outs["diag"][0].cpu().numpy(), d0.cpu().numpy(), mesh=mesh) < 1e-12
