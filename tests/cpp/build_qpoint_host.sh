#!/bin/bash
# Builds tests/cpp/libprod_qpoint.so: the quadrature-point physics of the CUDA sources compiled for the host
# (see qpoint_host.cpp).  The stretch is found by the two banner comments of glsb_kernels.cuh.
set -e
here=$(cd "$(dirname "$0")" && pwd)
src="$here/../../dealii_ns_gls_b200/csrc"
inc="$here/qpoint_product_extract.inc"
trap 'rm -f "$inc"' EXIT
awk '/^\/\/ quadrature-point physics \(operator_ns.cc:949-1182\), one point$/ {on = 1}
     /^\/\/ kernels$/ {on = 0}
     on {print}' "$src/glsb_kernels.cuh" > "$inc"
grep -q 'void qpoint_physics(const KParams<T> &p, const QTables<dim, T> &tb,' "$inc"
grep -q 'void symm_add(' "$inc"
g++ -std=c++17 -O2 -fPIC -shared -w -I"$src" -I"$here" -I/usr/local/cuda/include -o "$here/libprod_qpoint.so" \
    "$here/qpoint_host.cpp"
