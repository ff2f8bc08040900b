// dune-gdt_b200/csrc/fv_tma.hpp -- the TMA-staged variant of the scalar 2D FV apply (fv_tma.cu)
#pragma once
#include "kernels.hpp"

namespace gdtb {

// plain 2D apply (or fused Euler step) on an unpartitioned grid whose rows are multiples of 512 cells, no boundary
// treatments, no fused Runge-Kutta stage: everything else stays with k_fv_march
bool fv_tma_eligible(const FvParams& p, const double* u, const double* out);
int launch_fv_tma(Launch& L, const FvParams& p, const double* u, double* out);

} // namespace gdtb
