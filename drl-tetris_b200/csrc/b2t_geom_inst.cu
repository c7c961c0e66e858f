// One translation unit per field size: nvcc ... -DB2T_H=22 -DB2T_W=10 (see build.py).  Instantiates every kernel for
// that geometry and exports its launch table.
#include "b2t_kernels.cuh"

#ifndef B2T_H
#error "compile with -DB2T_H=<height> -DB2T_W=<width>"
#endif
#define B2T_CAT3(a, b, c) a##b##_##c
#define B2T_FN(h, w) B2T_CAT3(b2t_geom_ops_, h, w)

const b2t::GeomOps* B2T_FN(B2T_H, B2T_W)() { return b2t::Launch<b2t::Geo<B2T_H, B2T_W> >::ops(); }
