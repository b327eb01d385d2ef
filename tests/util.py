"""Shared helpers of the parity tests: build the oracle and the CUDA operator from one Mesh."""
import numpy as np

from oracle import gls_oracle as go


class TI:
    """plain time-integrator stand-in with fixed numbers"""

    def __init__(self, order, weights, dt, theta=1.0):
        self.order, self.weights, self.dt, self.theta = order, list(weights), dt, theta

    def get_primary_weight(self): return self.weights[0] if self.weights else 0.0
    def get_weights(self): return self.weights
    def get_order(self): return self.order
    def get_current_dt(self): return self.dt
    def get_theta(self): return self.theta


def make_oracle(mesh, ti, *, nu=0.1, c1=4.0, c2=2.0, ctd=False, increment_form=True, cell_wise=True,
                dtype=np.float64, path="sumfac"):
    return go.OracleOperator(dim=mesh.dim, degree=mesh.degree, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs,
                             cell_points=mesh.cell_points, mapping_degree=mesh.mapping_degree,
                             constraints=mesh.constraints, nu=nu, c1=c1, c2=c2, theta=ti.get_theta(),
                             order=ti.get_order(), consider_time_derivative=ctd,
                             increment_form=increment_form, cell_wise_stabilization=cell_wise,
                             dtype=dtype, path=path)


def make_gpu(mesh, ti, *, nu=0.1, c1=4.0, c2=2.0, ctd=False, increment_form=True, cell_wise=True,
             number="double", inhom=None, exchange=None):
    from dealii_ns_gls_b200.operator import NavierStokesOperator
    return NavierStokesOperator(mesh, inhom, nu, c1, c2, ti, ctd, increment_form, cell_wise, number=number,
                                exchange=exchange)


def rel_l2(a, b, mesh=None):
    """Relative l2 error of a against b.  With `mesh`, a and b are dof vectors and the error is the LARGER of
    the error over the unconstrained rows (the rows that do arithmetic: O(h^dim) entries) and the error over
    the constrained rows (identity rows carrying O(1) source values, which would otherwise dominate the
    norm and deflate the figure)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if mesh is not None and len(mesh.constraints) > 0 and a.shape == (mesh.n_dofs,):
        cons = np.zeros(mesh.n_dofs, dtype=bool)
        cons[np.fromiter(mesh.constraints.keys(), dtype=np.int64)] = True
        e_free = np.linalg.norm((a - b)[~cons]) / max(np.linalg.norm(b[~cons]), 1e-300)
        nb = np.linalg.norm(b[cons])
        e_cons = np.linalg.norm((a - b)[cons]) / nb if nb > 0 else float(np.linalg.norm(a[cons]))
        return max(e_free, e_cons)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
