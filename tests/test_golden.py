"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
  * not gpu: the oracle's sum-factorised numpy path and its C restatement reproduce them;
  * gpu: the CUDA path reproduces them through the C ABI, FP64 to 1e-12, FP32 to 2e-5."""
import glob
import os

import numpy as np
import pytest

from dealii_ns_gls_b200.mesh import Mesh
from tests.util import TI, make_gpu, make_oracle, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(name):
    g = dict(np.load(os.path.join(HERE, "golden", name + ".npz")))
    cons = {}
    for r, d in enumerate(g["row_dof"]):
        a, b = int(g["row_ptr"][r]), int(g["row_ptr"][r + 1])
        cons[int(d)] = [(int(m), float(w)) for m, w in zip(g["entry_col"][a:b], g["entry_val"][a:b])]
    dim = int(g["dim"])
    mesh = Mesh(dim=dim, degree=int(g["degree"]), n_cells=g["cell_dofs"].shape[0], n_dofs=int(g["n_dofs"]),
                n_owned=int(g["n_dofs"]), cell_dofs=g["cell_dofs"], geometry_type=int(g["geometry_type"]),
                cell_points=g["cell_points"], mapping_degree=int(g["mapping_degree"]), constraints=cons,
                cell_h_min=g["cell_h_min"], cell_measure=g["cell_measure"], n_global_dofs=int(g["n_dofs"]))
    if mesh.geometry_type == 0:
        mesh.cart_inv_jac, mesh.cart_det = g["inv_jac"], g["jxw"]
    ti = TI(int(g["order"]), list(g["weights"]), float(g["dt"]), theta=float(g["theta"]))
    flags = dict(nu=float(g["nu"]), ctd=bool(g["ctd"]), cell_wise=bool(g["cell_wise"]),
                 increment_form=bool(g["increment_form"]))
    return g, mesh, ti, flags


def test_fixtures_exist():
    assert len(CASES) >= 6


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    g, mesh, ti, flags = load(name)
    o = make_oracle(mesh, ti, path="sumfac", **flags)
    if ti.get_order() > 0:
        o.set_previous_solution(list(g["history"]), ti.get_weights())
    o.set_linearization_point(g["lin"], ti.get_current_dt())
    w = ti.get_primary_weight()
    assert rel_l2(o.vmult(g["src"], w), g["out_vmult"], mesh=mesh) < 1e-13
    assert rel_l2(o.evaluate_residual(g["src_bc"], w), g["out_residual"]) < 1e-13
    assert rel_l2(o.compute_inverse_diagonal(w), g["out_inv_diag"]) < 1e-12
    assert abs(o.get_max_u(g["src"]) - float(g["out_max_u"])) < 1e-14
    # C restatement of the cell loop (plain gather/scatter; constraints resolved around it)
    from oracle.gls_oracle_c import COracle
    co = COracle.from_numpy_oracle(o, 0 if flags["increment_form"] else 1)
    x = o._resolve(g["src"])
    got = np.asarray(o._distribute_transpose(co.apply(x, w)))
    if len(o.constrained):
        got[o.constrained] = g["src"][o.constrained]
    assert rel_l2(got, g["out_vmult"], mesh=mesh) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_golden(name, number):
    import torch
    g, mesh, ti, flags = load(name)
    tol = 1e-12 if number == "double" else 2e-5
    dt = torch.float64 if number == "double" else torch.float32
    dev = lambda a: torch.tensor(np.asarray(a), dtype=dt, device="cuda")  # noqa: E731
    op = make_gpu(mesh, ti, number=number, **flags)
    if ti.get_order() > 0:
        op.set_previous_solution([dev(h) for h in g["history"]])
    op.set_linearization_point(dev(g["lin"]))
    out = op.initialize_dof_vector()
    op.vmult(out, dev(g["src"]))
    assert rel_l2(out.cpu().numpy(), g["out_vmult"], mesh=mesh) < tol
    op.evaluate_residual(out, dev(g["src_bc"]))
    assert rel_l2(out.cpu().numpy(), g["out_residual"], mesh=mesh) < tol
    op.compute_inverse_diagonal(out)
    assert rel_l2(out.cpu().numpy(), g["out_inv_diag"], mesh=mesh) < (1e-11 if number == "double" else 5e-5)
    assert abs(op.get_max_u(dev(g["src"])) - float(g["out_max_u"])) < 10 * tol
    for tab, key in (("delta_1", "delta_1"), ("delta_2", "delta_2"), ("delta_1_q", "delta_1_q"), ("delta_2_q", "delta_2_q")):
        assert rel_l2(op.get_table(tab).cpu().numpy().reshape(-1), np.asarray(g[key]).reshape(-1)) < tol
