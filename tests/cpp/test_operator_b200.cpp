// C++ host-mirror test: reads tests/golden/turek_3d_q2_bdf2.bin (written by make_golden.py from the
// oracle), drives glsb::NavierStokesOperator<3,double> exactly like main.cc drives the reference
// operator (set_previous_solution, set_linearization_point, vmult, evaluate_residual,
// compute_inverse_diagonal, get_max_u) and compares with the stored oracle results.
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "../../dealii_ns_gls_b200/cpp/operator_b200.h"

template <typename T>
static std::vector<T> get(std::ifstream &f)
{
  std::uint64_t n = 0;
  f.read(reinterpret_cast<char *>(&n), 8);
  std::vector<T> v(n);
  f.read(reinterpret_cast<char *>(v.data()), n * sizeof(T));
  if (!f)
    {
      std::fprintf(stderr, "short read\n");
      std::exit(2);
    }
  return v;
}

static double rel_l2(const std::vector<double> &a, const std::vector<double> &b)
{
  double d = 0, n = 0;
  for (std::size_t i = 0; i < b.size(); ++i)
    {
      d += (a[i] - b[i]) * (a[i] - b[i]);
      n += b[i] * b[i];
    }
  return std::sqrt(d / n);
}

int main(int argc, char **argv)
{
  if (argc < 2)
    return 2;
  std::ifstream f(argv[1], std::ios::binary);
  if (!f)
    return 2;
  auto hdr = get<std::int64_t>(f);
  auto par = get<double>(f);
  auto wts = get<double>(f);
  glsb::MeshDescription m;
  m.dim = (int)hdr[0], m.degree = (int)hdr[1], m.geometry_type = (int)hdr[2];
  const int  order = (int)hdr[3];
  const bool ctd = hdr[4], cell_wise = hdr[5], increment_form = hdr[6];
  m.n_cells = hdr[7], m.n_owned = hdr[8], m.n_ghost = 0, m.n_global_dofs = hdr[8];
  m.dof_indices  = get<std::uint32_t>(f);
  m.row_dof      = get<std::uint32_t>(f);
  m.row_ptr      = get<std::uint32_t>(f);
  m.entry_col    = get<std::uint32_t>(f);
  m.entry_val    = get<double>(f);
  m.inv_jac      = get<double>(f);
  m.jxw          = get<double>(f);
  m.cell_h_min   = get<double>(f);
  m.cell_measure = get<double>(f);
  auto hist = get<double>(f), lin = get<double>(f), src = get<double>(f), src_bc = get<double>(f);
  auto out_vmult = get<double>(f), out_res = get<double>(f), out_diag = get<double>(f), out_maxu = get<double>(f);
  // flag constrained entries the way the adapter does
  std::vector<std::int64_t> row_of(m.n_owned, -1);
  for (std::size_t r = 0; r < m.row_dof.size(); ++r)
    row_of[m.row_dof[r]] = (std::int64_t)r;
  for (auto &i : m.dof_indices)
    if (row_of[i] >= 0)
      i = GLSB_CONSTRAINED_BIT | (std::uint32_t)row_of[i];
  m.constrained_indices = m.row_dof;

  // BDF2 with two equal steps reproduces the stored weights (15, -20, 5) for dt = 0.1
  glsb::TimeIntegratorDataBDF ti(order);
  ti.update_dt(par[4]);
  ti.update_dt(par[4]);
  for (int i = 0; i <= order; ++i)
    if (std::abs(ti.get_weights()[i] - wts[i]) > 1e-12)
      {
        std::printf("FAIL weights\n");
        return 1;
      }
  try
    {
      glsb::NavierStokesOperator<3, double> op(m, par[0], par[1], par[2], ti, ctd, increment_form, cell_wise);
      glsb::SolutionHistory<double>         history(order + 1);
      const std::size_t                      n = m.n_owned;
      for (int i = 0; i <= order; ++i)
        history.get_vectors()[i].copy_from_host(std::vector<double>(hist.begin() + i * n, hist.begin() + (i + 1) * n));
      op.set_previous_solution(history);
      glsb::DeviceVector<double> v_lin, v_src, v_bc, dst;
      v_lin.copy_from_host(lin), v_src.copy_from_host(src), v_bc.copy_from_host(src_bc);
      op.set_linearization_point(v_lin);
      op.initialize_dof_vector(dst);
      op.vmult(dst, v_src);
      const double e1 = rel_l2(dst.to_host(), out_vmult);
      op.evaluate_residual(dst, v_bc);
      const double e2 = rel_l2(dst.to_host(), out_res);
      op.compute_inverse_diagonal(dst);
      const double e3 = rel_l2(dst.to_host(), out_diag);
      const double e4 = std::abs(op.get_max_u(v_src) - out_maxu[0]);
      std::printf("variant %s vmult %.2e residual %.2e inv_diag %.2e max_u %.1e\n", op.vmult_variant(), e1, e2, e3, e4);
      bool ok = e1 < 1e-12 && e2 < 1e-12 && e3 < 1e-11 && e4 < 1e-13;
      // error behaviour: vmult before set_linearization_point must throw
      glsb::NavierStokesOperator<3, double> op2(m, par[0], par[1], par[2], ti, ctd, increment_form, cell_wise);
      bool threw = false;
      try
        {
          op2.vmult(dst, v_src);
        }
      catch (const glsb::Error &)
        {
          threw = true;
        }
      ok = ok && threw;
      std::printf(ok ? "PASS\n" : "FAIL\n");
      return ok ? 0 : 1;
    }
  catch (const std::exception &e)
    {
      std::printf("FAIL exception: %s\n", e.what());
      return 1;
    }
}
