/* Host-side helpers of the synthetic structured-block generator (dealii_ns_gls_b200/mesh.py): cell traversal order,
 * first-touch node numbering and the assembly of the per-cell dof index array.
 *
 * In the reference these arrays come out of deal.II (p4est's Morton order of the cells, DoFHandler::distribute_dofs,
 * performance.cc:29-42, main.cc:230-256); mesh.py imitates them for structured blocks and spent ~20 s in numpy
 * passes at the bench size (160^3 cells, 1.1e8 cell-local nodes).  Same results as the numpy code paths of
 * mesh.py (tests/test_meshgen_native.py compares them), O(cells) and a few hundred ms.  Not part of the operator's
 * C ABI (include/glsb200.h): built into its own small library, libglsb_meshgen.so, no CUDA involved.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXDIM 3
#define MAXLOC 125

typedef struct
{
  int      dim, nbits;
  int64_t  shape[MAXDIM];
  int64_t *cc; /* [ncell][dim], filled in traversal order */
  int64_t  count;
} walk_t;

/* cells of the block in ascending Morton key (bits interleaved, x lowest): descend the 2^dim-tree from the top
 * bit, children in key order, boxes outside the block skipped */
static void morton_walk(walk_t *w, int bit, const int64_t *origin)
{
  const int dim = w->dim;
  if (bit < 0)
    {
      memcpy(w->cc + w->count * dim, origin, sizeof(int64_t) * dim);
      w->count++;
      return;
    }
  for (int child = 0; child < (1 << dim); ++child)
    {
      int64_t o[MAXDIM];
      int     inside = 1;
      for (int e = 0; e < dim; ++e)
        {
          o[e] = origin[e] + ((int64_t)((child >> e) & 1) << bit);
          if (o[e] >= w->shape[e])
            inside = 0;
        }
      if (inside)
        morton_walk(w, bit - 1, o);
    }
}

/* Cells in traversal order and the first-touch numbering of the nodes of a structured block.
 *   shape[dim] cells per direction, p = degree, periodic[dim] (node identification), morton != 0: Morton order of
 *   the cells, else lexicographic (x fastest).
 * Outputs (caller-allocated): cc[ncell][dim]; cell_nodes[ncell][n_loc] lexicographic grid node ids (x fastest)
 * of the local nodes (local index x fastest); node_rank[nnode] position of each grid node in the numbering "walk
 * the cells in order, inside a cell vertices, then line, quad, hex nodes, each group lexicographic, and number a
 * node when it is met first"; first_cell[nnode] the cell that met it first.  Returns 0, or 1 on bad arguments. */
int glsm_number_nodes(int dim, const int64_t *shape, int p, const uint8_t *periodic, int morton, int64_t *cc,
                      int64_t *cell_nodes, int64_t *node_rank, int64_t *first_cell)
{
  if (dim < 1 || dim > MAXDIM || p < 1 || p > 4)
    return 1;
  const int n     = p + 1;
  int       n_loc = 1;
  int64_t   ncell = 1, nnode = 1, npts[MAXDIM];
  for (int e = 0; e < dim; ++e)
    {
      if (shape[e] < 1)
        return 1;
      n_loc *= n;
      ncell *= shape[e];
      npts[e] = p * shape[e] + (periodic[e] ? 0 : 1);
      nnode *= npts[e];
    }
  /* ---- traversal order ---- */
  if (morton)
    {
      int64_t mx = 0;
      for (int e = 0; e < dim; ++e)
        if (shape[e] - 1 > mx)
          mx = shape[e] - 1;
      int nbits = 1;
      while ((mx >> nbits) != 0)
        ++nbits;
      walk_t  w;
      int64_t origin[MAXDIM] = {0, 0, 0};
      w.dim                  = dim;
      w.nbits                = nbits;
      w.cc                   = cc;
      w.count                = 0;
      for (int e = 0; e < dim; ++e)
        w.shape[e] = shape[e];
      morton_walk(&w, nbits - 1, origin);
      if (w.count != ncell)
        return 1;
    }
  else
    {
      for (int64_t k = 0; k < ncell; ++k)
        {
          int64_t r = k;
          for (int e = 0; e < dim; ++e)
            {
              cc[k * dim + e] = r % shape[e];
              r /= shape[e];
            }
        }
    }
  /* ---- local offsets (x fastest) and the order of the local nodes inside a cell ---- */
  int loc[MAXLOC][MAXDIM], ent[MAXLOC], colorder[MAXLOC];
  for (int l = 0; l < n_loc; ++l)
    {
      int r  = l;
      ent[l] = 0;
      for (int e = 0; e < dim; ++e)
        {
          loc[l][e] = r % n;
          r /= n;
          if (loc[l][e] > 0 && loc[l][e] < p)
            ent[l]++;
        }
    }
  int k = 0;
  for (int d = 0; d <= dim; ++d)
    for (int l = 0; l < n_loc; ++l)
      if (ent[l] == d)
        colorder[k++] = l;
  /* ---- grid node ids of every cell ---- */
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncell; ++c)
    for (int l = 0; l < n_loc; ++l)
      {
        int64_t id = 0, mul = 1;
        for (int e = 0; e < dim; ++e)
          {
            id += ((p * cc[c * dim + e] + loc[l][e]) % npts[e]) * mul;
            mul *= npts[e];
          }
        cell_nodes[c * n_loc + l] = id;
      }
  /* ---- first touch ---- */
  for (int64_t i = 0; i < nnode; ++i)
    node_rank[i] = -1;
  int64_t next = 0;
  for (int64_t c = 0; c < ncell; ++c)
    for (int j = 0; j < n_loc; ++j)
      {
        const int64_t id = cell_nodes[c * n_loc + colorder[j]];
        if (node_rank[id] < 0)
          {
            node_rank[id]  = next++;
            first_cell[id] = c;
          }
      }
  return next == nnode ? 0 : 1;
}

/* cell_dofs[cell][c * n_loc + l] = local_of_node[cell_nodes[cell][l]] * C + c (all components of a node
 * consecutive), and is_boundary[cell] = 1 if the cell touches an index >= n_owned (a ghost).  index_bytes = 4
 * (uint32) or 8 (int64) selects the type of cell_dofs.  Returns 0, or 1 if an index does not fit. */
int glsm_assemble_cell_dofs(int64_t ncell, int n_loc, int C, const int64_t *cell_nodes, const int64_t *local_of_node,
                            int64_t n_owned, int index_bytes, void *cell_dofs, uint8_t *is_boundary)
{
  if (index_bytes != 4 && index_bytes != 8)
    return 1;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t c = 0; c < ncell; ++c)
    {
      uint8_t ghost = 0;
      for (int l = 0; l < n_loc; ++l)
        {
          const int64_t base = local_of_node[cell_nodes[c * n_loc + l]] * C;
          if (base + C - 1 >= n_owned)
            ghost = 1;
          for (int comp = 0; comp < C; ++comp)
            {
              const int64_t v = base + comp;
              const int64_t o = c * (int64_t)(C * n_loc) + (int64_t)comp * n_loc + l;
              if (index_bytes == 4)
                {
                  if (v < 0 || v > 0xffffffffLL)
                    bad = 1;
                  ((uint32_t *)cell_dofs)[o] = (uint32_t)v;
                }
              else
                ((int64_t *)cell_dofs)[o] = v;
            }
        }
      is_boundary[c] = ghost;
    }
  return bad;
}

/* J^-1 and JxW at the quadrature points of every cell from its mapping support points (what MatrixFree's
 * MappingInfo stores for "general" cells; MappingQ(k), main.cc:253-254).
 *   T[e][q][m]: tensor-product tables, derivative in direction e of the mapping's shape function m at point q
 *   (n_q points, n_m support points per cell, both lexicographic with x fastest); wq[q] quadrature weights;
 *   points[cell][m][dim].
 * Outputs: inv_jac[cell][q][e][j] = (J^-1)_{e j} with J_{i e} = d x_i / d xi_e, jxw[cell][q] = det J * w_q.
 * Returns the number of points with det J <= 0 (0 for a valid mesh), or -1 on bad arguments. */
int64_t glsm_general_geometry(int dim, int64_t ncell, int n_q, int n_m, const double *T, const double *wq,
                              const double *points, double *inv_jac, double *jxw)
{
  if (dim < 2 || dim > 3 || n_q < 1 || n_m < 1)
    return -1;
  int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t c = 0; c < ncell; ++c)
    {
      const double *X = points + c * (int64_t)n_m * dim;
      for (int q = 0; q < n_q; ++q)
        {
          double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
          for (int e = 0; e < dim; ++e)
            {
              const double *t = T + ((int64_t)e * n_q + q) * n_m;
              double        s[3] = {0, 0, 0};
              for (int m = 0; m < n_m; ++m)
                for (int i = 0; i < dim; ++i)
                  s[i] += t[m] * X[m * dim + i];
              for (int i = 0; i < dim; ++i)
                J[i][e] = s[i];
            }
          double *out = inv_jac + (c * (int64_t)n_q + q) * dim * dim;
          double  det;
          if (dim == 2)
            {
              det               = J[0][0] * J[1][1] - J[0][1] * J[1][0];
              const double r    = 1.0 / det;
              out[0 * 2 + 0]    = J[1][1] * r;
              out[0 * 2 + 1]    = -J[0][1] * r;
              out[1 * 2 + 0]    = -J[1][0] * r;
              out[1 * 2 + 1]    = J[0][0] * r;
            }
          else
            {
              const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
              const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
              const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
              det              = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
              const double r   = 1.0 / det;
              out[0]           = c00 * r;
              out[1]           = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
              out[2]           = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
              out[3]           = c01 * r;
              out[4]           = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
              out[5]           = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
              out[6]           = c02 * r;
              out[7]           = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
              out[8]           = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
            }
          if (!(det > 0))
            bad++;
          jxw[c * (int64_t)n_q + q] = det * wq[q];
        }
    }
  return bad;
}

/* minimum_vertex_distance() and measure() of every cell from its 2^dim vertices (verts[cell][v][dim], vertex v at
 * the corner with bit e of v set in direction e): the smallest distance between two vertices, and the volume of
 * the multilinear cell by 2-point Gauss quadrature per direction (exact for it).  Same formulas as
 * mesh.py::_vertex_geometry. */
int glsm_vertex_geometry(int dim, int64_t ncell, const double *verts, double *h_min, double *measure)
{
  if (dim < 2 || dim > 3)
    return 1;
  const int    nv = 1 << dim;
  const double g[2] = {0.5 - 0.5 / 1.7320508075688772, 0.5 + 0.5 / 1.7320508075688772};
  /* d phi_v / d xi_e at the 2^dim Gauss points */
  double dphi[8][8][3];
  for (int qi = 0; qi < nv; ++qi)
    for (int v = 0; v < nv; ++v)
      for (int e = 0; e < dim; ++e)
        {
          double t = 1.0;
          for (int f = 0; f < dim; ++f)
            {
              const int    bit = (v >> f) & 1;
              const double xi  = g[(qi >> f) & 1];
              t *= (f == e) ? (bit ? 1.0 : -1.0) : (bit ? xi : 1.0 - xi);
            }
          dphi[qi][v][e] = t;
        }
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncell; ++c)
    {
      const double *X  = verts + c * (int64_t)nv * dim;
      double        h2 = 1e300;
      for (int a = 0; a < nv; ++a)
        for (int b = a + 1; b < nv; ++b)
          {
            double s = 0;
            for (int i = 0; i < dim; ++i)
              s += (X[a * dim + i] - X[b * dim + i]) * (X[a * dim + i] - X[b * dim + i]);
            if (s < h2)
              h2 = s;
          }
      double meas = 0;
      for (int qi = 0; qi < nv; ++qi)
        {
          double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
          for (int v = 0; v < nv; ++v)
            for (int i = 0; i < dim; ++i)
              for (int e = 0; e < dim; ++e)
                J[i][e] += X[v * dim + i] * dphi[qi][v][e];
          double det;
          if (dim == 2)
            det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          else
            det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                  J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
          meas += det / nv;
        }
      h_min[c]   = __builtin_sqrt(h2);
      measure[c] = meas;
    }
  return 0;
}
