#!/bin/bash
# Builds oracle/_ref/libref_qpoint.so: the reference's own quadrature-point kernels (do_vmult_cell, both branches,
# and symm_scalar_product_add = lines 880-1182 of include/operator_ns.cc; the cell loop of compute_penalty_parameters,
# 348-421; do_vmult_boundary, 1195-1301), cut out of the reference tree at build
# time, checked for the statements expected at its ends, compiled UNMODIFIED inside ref_qpoint_harness.cc on the
# stand-in types of ref_shim/qpoint_shim.h.  The scratch file never leaves oracle/_ref/ and is deleted again.
# Usage: build_ref_qpoint.sh <reference root>      (called by `make -C oracle _ref`)
set -e
ref=${1:-/root/reference}
here=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$here/_ref"
inc="$here/_ref/qpoint_extract.inc"
trap 'rm -f "$inc"' EXIT
sed -n '880,1182p' "$ref/include/operator_ns.cc" > "$inc"
[ "$(sed -n '2p' "$inc")" = "namespace" ]
grep -q 'symm_scalar_product_add(Tensor<1, dim_, Tensor<1, dim, Number>> &v_gradient,' "$inc"
grep -q 'NavierStokesOperator<dim, Number>::do_vmult_cell(' "$inc"
[ "$(grep -c 'integrator.integrate(' "$inc")" = 2 ]
[ "$(tail -1 "$inc")" = "}" ]
# the body of compute_penalty_parameters' cell loop: tau / stau, cell-wise and q-point-wise delta (:348-421)
pen="$here/_ref/penalty_extract.inc"
trap 'rm -f "$inc" "$pen"' EXIT
sed -n '348,421p' "$ref/include/operator_ns.cc" > "$pen"
[ "$(sed -n '1p' "$pen")" = "  const auto tau  = this->time_integrator_data.get_current_dt();" ]
grep -q 'delta_2_q\[cell\]\[q\] = std::sqrt(u_mag_squared) \* h \* 0.5;' "$pen"
[ "$(tail -1 "$pen")" = "    }" ]
# effective_beta_face of the outflow faces (:428-457)
bet="$here/_ref/beta_extract.inc"
trap 'rm -f "$inc" "$pen" "$bet"' EXIT
sed -n '428,457p' "$ref/include/operator_ns.cc" > "$bet"
[ "$(sed -n '1p' "$bet")" = "      const double beta = 1.0; // TODO" ]
grep -q 'beta / std::pow(cell_size, static_cast<Number>(fe_degree + 1));' "$bet"
[ "$(tail -1 "$bet")" = "        }" ]
# do_vmult_boundary: the "cut" and Nitsche outflow-face terms (:1195-1301)
bnd="$here/_ref/boundary_extract.inc"
trap 'rm -f "$inc" "$pen" "$bet" "$bnd"' EXIT
sed -n '1195,1301p' "$ref/include/operator_ns.cc" > "$bnd"
[ "$(sed -n '1p' "$bnd")" = "template <int dim, typename Number>" ]
grep -q 'NavierStokesOperator<dim, Number>::do_vmult_boundary(' "$bnd"
grep -q 'normal_outflux = std::min(zero, normal_outflux);' "$bnd"
[ "$(tail -1 "$bnd")" = "}" ]
g++ -std=c++17 -O2 -fPIC -shared -I"$here/ref_shim" -I"$here/_ref" -o "$here/_ref/libref_qpoint.so" \
    "$here/ref_qpoint_harness.cc"
