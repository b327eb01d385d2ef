"""True 2:1 hanging-node meshes (mesh.hypercube_hanging; the reference gets them from
DoFTools::make_hanging_node_constraints, main.cc:293, simulation.cc:803-809, input/rotation.json):
known answers for the constraint rows on CPU, CUDA parity against the oracle on the GPU."""
import numpy as np
import pytest

from dealii_ns_gls_b200 import mesh as gm
from tests.util import TI, make_gpu, make_oracle, rel_l2


def _walls(x, c):
    on = (np.abs(x) < 1e-12).any(axis=1) | (np.abs(x - 1.0) < 1e-12).any(axis=1)
    return on if c < x.shape[1] else np.zeros(len(x), dtype=bool)


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2)])
def test_hanging_rows_known_answers(dim, degree):
    m = gm.hypercube_hanging(dim, 2, degree)
    C = dim + 1
    assert len(m.hanging_nodes) > 0
    # partition of unity, masters unconstrained and of the same component
    for dof, row in m.constraints.items():
        assert abs(sum(w for _, w in row) - 1.0) < 1e-13
        for mst, _ in row:
            assert mst not in m.constraints and mst % C == dof % C
    # deal.II's FE_Q interface constraints: Q1 1/2, 1/2 on lines and 1/4 x 4 on faces; Q2 lines 3/8, 3/4, -1/8
    ws = sorted({tuple(sorted(round(w, 12) for _, w in row)) for row in m.constraints.values()})
    if degree == 1:
        assert ws == ([(0.5, 0.5)] if dim == 2 else [(0.25, 0.25, 0.25, 0.25), (0.5, 0.5)])
    if degree == 2:
        assert (-0.125, 0.375, 0.75) in ws
    # a polynomial of degree <= p per variable is in the FE space on both sides of the interface: its nodal
    # values satisfy every hanging-node row exactly
    x = m.node_xyz
    f = np.prod([(0.3 + x[:, e]) ** degree + 0.5 * x[:, e] for e in range(dim)], axis=0)
    for nd in m.hanging_nodes:
        row = m.constraints[int(nd) * C]
        assert abs(f[nd] - sum(w * f[mst // C] for mst, w in row)) < 1e-12
    # 3-D block refinement has hanging nodes on faces (up to n^2 masters) and on lines (n masters)
    if dim == 3:
        sizes = {len(r) for r in m.constraints.values()}
        assert (degree + 1) in sizes and max(sizes) == (degree + 1) ** 2 - (1 if degree > 1 else 0) * 0


@pytest.mark.parametrize("dim,degree", [(2, 2), (3, 1)])
def test_oracle_paths_agree_on_hanging_mesh(dim, degree):
    m = gm.hypercube_hanging(dim, 2, degree, dirichlet=_walls)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    rng = np.random.default_rng(5)
    outs = []
    for path in ("naive", "sumfac"):
        o = make_oracle(m, ti, path=path, ctd=True, cell_wise=False, nu=0.01)
        hist = [rng.uniform(-1, 1, m.n_dofs) for _ in range(3)] if not outs else hist  # noqa: F821
        lin = rng.uniform(-1, 1, m.n_dofs) if not outs else lin  # noqa: F821
        src = rng.uniform(-1, 1, m.n_dofs) if not outs else src  # noqa: F821
        o.set_previous_solution(hist, ti.get_weights())
        o.set_linearization_point(lin, 0.1)
        outs.append(o.vmult(src, 15.0))
    assert rel_l2(outs[0], outs[1], mesh=m) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("dim,degree,walls", [(2, 1, False), (2, 2, True), (2, 3, False), (3, 1, True), (3, 2, False),
                                              (3, 2, True), (3, 3, False)])
def test_cuda_on_hanging_mesh(dim, degree, walls, number):
    """vmult, residual and inverse diagonal with real hanging-node rows (read_dof_values resolves them,
    distribute_local_to_global applies the transpose, operator_ns.cc:806-830)."""
    import torch
    m = gm.hypercube_hanging(dim, 2 if dim == 3 else 4, degree, dirichlet=_walls if walls else None)
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)
    dt = np.float64 if number == "double" else np.float32
    tol = 1e-12 if number == "double" else 2e-5
    tdt = torch.float64 if number == "double" else torch.float32
    dev = lambda a: torch.tensor(np.asarray(a), dtype=tdt, device="cuda")  # noqa: E731
    rng = np.random.default_rng(77)
    hist = [rng.uniform(-1, 1, m.n_dofs) for _ in range(3)]
    lin, src = rng.uniform(-1, 1, m.n_dofs), rng.uniform(-1, 1, m.n_dofs)
    ora = make_oracle(m, ti, dtype=dt, ctd=True, cell_wise=False, nu=0.01)
    gpu = make_gpu(m, ti, number=number, ctd=True, cell_wise=False, nu=0.01)
    ora.set_previous_solution(hist, ti.get_weights())
    gpu.set_previous_solution([dev(h) for h in hist])
    ora.set_linearization_point(lin, 0.1)
    gpu.set_linearization_point(dev(lin))
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, dev(src))
    assert rel_l2(dst.cpu().numpy(), ora.vmult(src, 15.0), mesh=m) < tol
    gpu.evaluate_residual(dst, dev(src))
    assert rel_l2(dst.cpu().numpy(), ora.evaluate_residual(src, 15.0), mesh=m) < tol
    gpu.compute_inverse_diagonal(dst)
    assert rel_l2(dst.cpu().numpy(), ora.compute_inverse_diagonal(15.0), mesh=m) < (1e-11 if number == "double" else 5e-5)
