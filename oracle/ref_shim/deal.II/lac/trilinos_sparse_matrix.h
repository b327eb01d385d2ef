// Stand-in for <deal.II/lac/trilinos_sparse_matrix.h> (see la_parallel_vector.h next to this file): config.h of
// the reference only names the type.
#pragma once

namespace dealii
{
  namespace TrilinosWrappers
  {
    class SparseMatrix
    {};
  } // namespace TrilinosWrappers
} // namespace dealii
