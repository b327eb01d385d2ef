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

// two levels of a structured 2-D Q1 block (nx x ny coarse cells, refined once), lexicographic node numbering,
// the dim + 1 components of a node consecutive: what MGTwoLevelTransfer::reinit receives
static bool check_transfer_and_vector_ops()
{
  const int nx = 3, ny = 2, C = 3;
  auto      node = [](int i, int j, int npx) { return j * npx + i; };
  const int cpx = nx + 1, fpx = 2 * nx + 1, n_coarse = cpx * (ny + 1) * C, n_fine = fpx * (2 * ny + 1) * C;
  std::vector<std::uint32_t> cidx, fidx;
  std::vector<double>        touch(n_fine, 0.0);
  auto cell_dofs = [&](int i, int j, int npx, std::vector<std::uint32_t> &out) {
    for (int c = 0; c < C; ++c)
      for (int l = 0; l < 4; ++l)
        out.push_back((std::uint32_t)(node(i + (l & 1), j + (l >> 1), npx) * C + c));
  };
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i)
      {
        cell_dofs(i, j, cpx, cidx);
        for (int ch = 0; ch < 4; ++ch)
          cell_dofs(2 * i + (ch & 1), 2 * j + (ch >> 1), fpx, fidx);
      }
  for (auto g : fidx)
    touch[g] += 1.0;
  std::vector<double> w(n_fine);
  for (int g = 0; g < n_fine; ++g)
    w[g] = 1.0 / touch[g];
  glsb::MGTwoLevelTransfer<2, double> transfer;
  transfer.reinit(1, n_fine, n_coarse, cidx, fidx, w);
  // a (bi)linear function per component is prolongated exactly and interpolated back exactly
  auto f = [](double x, double y, int c) { return (1 + c) * (0.3 + 2 * x - y + 0.5 * x * y); };
  std::vector<double> hc(n_coarse), hf(n_fine);
  for (int j = 0; j <= ny; ++j)
    for (int i = 0; i <= nx; ++i)
      for (int c = 0; c < C; ++c)
        hc[node(i, j, cpx) * C + c] = f(i, j, c);
  for (int j = 0; j <= 2 * ny; ++j)
    for (int i = 0; i <= 2 * nx; ++i)
      for (int c = 0; c < C; ++c)
        hf[node(i, j, fpx) * C + c] = f(0.5 * i, 0.5 * j, c);
  glsb::DeviceVector<double> dc, df(n_fine), back(n_coarse);
  dc.copy_from_host(hc);
  transfer.prolongate_and_add(df, dc);
  const double e1 = rel_l2(df.to_host(), hf);
  transfer.interpolate(back, df);
  const double e2 = rel_l2(back.to_host(), hc);
  // restriction is the transpose: (P c) . g == c . (P^T g)
  std::vector<double> hg(n_fine);
  for (int g = 0; g < n_fine; ++g)
    hg[g] = std::sin(0.37 * g) + 0.1;
  glsb::DeviceVector<double> dg, rc(n_coarse);
  dg.copy_from_host(hg);
  transfer.restrict_and_add(rc, dg);
  using Ops = glsb::DeviceVectorOps<double>;
  const double lhs = Ops::multi_dot(df, 1, dg)[0], rhs = Ops::multi_dot(dc, 1, rc)[0];
  const double e3  = std::abs(lhs - rhs) / std::abs(lhs);
  // batched inner products and the k-term update against host loops
  const int k = 5, n = 1001;
  std::vector<double> hV((size_t)k * n), hw(n), coef(k);
  for (size_t i = 0; i < hV.size(); ++i)
    hV[i] = std::cos(0.11 * i);
  for (int i = 0; i < n; ++i)
    hw[i] = std::sin(0.07 * i) - 0.2;
  glsb::DeviceVector<double> V, wv;
  V.copy_from_host(hV), wv.copy_from_host(hw);
  const auto dots = Ops::multi_dot(V, k, wv);
  double     e4   = 0;
  for (int j = 0; j < k; ++j)
    {
      double s = 0;
      for (int i = 0; i < n; ++i)
        s += hV[(size_t)j * n + i] * hw[i];
      e4      = std::max(e4, std::abs(s - dots[j]) / (1 + std::abs(s)));
      coef[j] = 0.5 - 0.1 * j;
    }
  Ops::multi_axpy(wv, V, coef, -1.0);
  std::vector<double> ref(hw);
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < n; ++i)
      ref[i] -= coef[j] * hV[(size_t)j * n + i];
  const double e5 = rel_l2(wv.to_host(), ref);
  glsb::DeviceVector<float> wf(n);
  glsb::DeviceVectorOps<float>::convert(wf, wv);
  std::printf("transfer: prolongation %.1e interpolation %.1e transpose %.1e; multi_dot %.1e multi_axpy %.1e\n", e1, e2, e3, e4,
              e5);
  return e1 < 1e-14 && e2 < 1e-14 && e3 < 1e-13 && e4 < 1e-13 && e5 < 1e-14 && std::abs(wf.to_host()[7] - (float)ref[7]) < 1e-6;
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
      ok = check_transfer_and_vector_ops() && ok;
      std::printf(ok ? "PASS\n" : "FAIL\n");
      return ok ? 0 : 1;
    }
  catch (const std::exception &e)
    {
      std::printf("FAIL exception: %s\n", e.what());
      return 1;
    }
}
