#!/bin/sh
# Copies the reference's own golden OUTPUT files (data, not source) that pin the hot path
#  into tests/golden/.  Run in the build container where /root/reference is
# mounted; the GPU box has no /root/reference, so the copies are what the tests read.
set -e
REF=${1:-/root/reference}
HERE=$(dirname "$0")
for k in ib_4 ib_6 bspline_3 bspline_4 piecewise_linear ib_3 bspline_5 bspline_6 piecewise_cubic ib_5 piecewise_constant \
         composite_bspline_32 composite_bspline_23 composite_bspline_43 composite_bspline_34 composite_bspline_54 \
         composite_bspline_45 composite_bspline_65 composite_bspline_56 discontinuous_linear ib_4_w8; do
  for d in 2d 3d; do
    cp "$REF/tests/interpolate/interpolate_01_$d.$k.output" "$HERE/"
  done
done
for d in 2d 3d; do
  for c in cell side; do
    for v in a b; do
      cp "$REF/tests/IBTK/ghost_accumulation_01_$d.$c.spread.$v.output" "$HERE/"
    done
    cp "$REF/tests/IBTK/ghost_accumulation_01_$d.$c.spread.b.mpirun=4.output" "$HERE/ghost_accumulation_01_$d.$c.spread.b.mpirun4.output"
  done
  cp "$REF/tests/IBTK/index_utilities_$d.output" "$HERE/"
done
cp "$REF/examples/IB/explicit/ex1/curve2d_64.vertex" "$HERE/"
# sample structure files (inputs of the reference's own examples) for the reader / force tests (N1, N2)
cp "$REF/examples/IB/explicit/ex1/curve2d_64.spring" "$HERE/"
for e in vertex spring beam target; do
  cp "$REF/examples/IB/explicit/ex3/fila_256.$e" "$HERE/"
done
