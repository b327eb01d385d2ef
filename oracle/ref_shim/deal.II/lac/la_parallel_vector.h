// Stand-in for <deal.II/lac/la_parallel_vector.h>, written for this repository (TEST INFRASTRUCTURE): just enough
// for the reference's include/config.h and include/time_integration.cc to compile UNMODIFIED from
// /root/reference (oracle/Makefile, target _ref).  deal.II itself is not available in this image; nothing of it
// is copied here.  What time_integration.cc needs: the Vector class template with copy_locally_owned_data_from
// (SolutionHistory::commit_solution), AssertThrow / ExcMessage, and the std headers deal.II pulls in.
#pragma once

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

namespace dealii
{
  struct ExcMessage : std::runtime_error
  {
    explicit ExcMessage(const std::string &m)
      : std::runtime_error(m)
    {}
  };

  namespace LinearAlgebra
  {
    namespace distributed
    {
      template <typename Number>
      class Vector
      {
      public:
        std::vector<Number> values;

        void
        copy_locally_owned_data_from(const Vector<Number> &src)
        {
          values = src.values;
        }
      };
    } // namespace distributed
  }   // namespace LinearAlgebra
} // namespace dealii

#define AssertThrow(cond, exc) \
  do                           \
    {                          \
      if (!(cond))             \
        throw exc;             \
    }                          \
  while (false)
