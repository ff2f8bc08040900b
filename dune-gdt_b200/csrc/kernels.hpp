// dune-gdt_b200/csrc/kernels.hpp -- host-callable launchers of the CUDA kernels (internal).
#pragma once

#include <vector>

#include "common.cuh"

namespace gdtb {

// kernel families for the optional per-kernel CUDA-event timing (gdtb_ctx_enable_timing)
enum KernelFamily
{
  KF_Q1_GATHER = 0,
  KF_FV_APPLY,
  KF_ELEMENT_MATRIX,
  KF_ELEMENT_VECTOR,
  KF_COUPLING_MATRIX,
  KF_BOUNDARY_MATRIX,
  KF_COUNT
};

struct Timing
{
  bool enabled = false;
  std::vector<cudaEvent_t> start[KF_COUNT], stop[KF_COUNT];
};

struct Launch
{
  cudaStream_t stream;
  long long count; // kernels launched through this object
  int sm_count;
  Timing* timing;
};

// CUDA events on the launching stream, directly around the kernel launch(es) of one family
inline void time_begin(Launch& L, int family)
{
  if (L.timing && L.timing->enabled) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, L.stream);
    L.timing->start[family].push_back(e);
  }
}

inline void time_end(Launch& L, int family)
{
  if (L.timing && L.timing->enabled) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, L.stream);
    L.timing->stop[family].push_back(e);
  }
}

// ---- generic, quadrature-faithful kernels (assemble_generic.cu) -----------------------------------
int launch_element_matrix(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, const long long* rowptr,
                          const int* colidx, double* values, int* error_flag);
int launch_element_vector(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, double* vec);
int launch_coupling_matrix(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, int filter,
                           const long long* rowptr, const int* colidx, double* values, int* error_flag);
int launch_boundary_matrix(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, const long long* rowptr,
                           const int* colidx, double* values, int* error_flag);

// ---- CG-Q1 row-gather assembly (assemble_q1_gather.cu) --------------------------------------------
constexpr int Q1G_MAX_ELEM_CHANNELS = 3;

// The local matrix of an axis-aligned cell with an element-constant coefficient c is c * Lref, where Lref
// (reference tensor) only depends on the cell's extents, the rule and the integrand.  A "channel" is one
// such (coefficient source, Lref) pair; all constant-coefficient summands are merged into channel 0.
struct Q1GatherParams
{
  GridDev g;
  int has_const;                             // channel with coefficient 1
  int n_elem;                                // channels with a per-element coefficient array
  double T_const[8][8];                      // T[o][s] = Lref[i(o)][s], see assemble_q1_gather.cu
  double T_elem[Q1G_MAX_ELEM_CHANNELS][8][8];
  const double* coef[Q1G_MAX_ELEM_CHANNELS]; // device arrays indexed by element
  // right-hand side: b[v] = rhs_const * sum_o valid(o) + sum_o rhs_elem_scale * f[e_o]
  //                         + rhs_sep_scale * prod_k B_k[i_k]
  int has_rhs;
  int rhs_has_const;
  int rhs_has_elem;
  int rhs_has_sep;
  double rhs_const;          // c * ie * prod_k s1 = per (vertex, element) contribution of a constant source
  double rhs_elem_scale;     // ie * prod_k s1
  const double* rhs_elem;    // per-element source values
  double rhs_sep_scale;      // p0 * w * ie
  const double* rhs_sep_tab; // 3 tables B_k[i_k], k-th table at offset k * rhs_sep_stride
  long long rhs_sep_stride;
  long long value_offset; // global CSR position of this process' first row (values points at it)
  long long row_offset;   // first vertex row of this process
  // owner-computes-rows slab along the last direction: vertex layers [row_lo, row_hi) are produced here from the
  // element layers [elem_lo, elem_hi) (the owned ones plus the ghost layer below)
  long long row_lo, row_hi;
  long long elem_lo, elem_hi;
};

int launch_q1_gather(Launch& L, const Q1GatherParams& p, double* values, double* rhs, bool accumulate);

// builds the separable right-hand-side tables B_k[i_k] for a product-separable built-in source
int launch_q1_rhs_tables(Launch& L, const GridDev& g, long long elem_lo, long long elem_hi, const FnDev& f, int m,
                         const double* qx, const double* qw, const double* phi /* [m][2] */, double* tab,
                         long long stride);

// ---- sparsity pattern (pattern.cu) ----------------------------------------------------------------
int pattern_sort_unique(Launch& L, const GridDev& g, const SpaceDev& test, const SpaceDev& ansatz, int stencil,
                        long long** d_rowptr, int** d_colidx, long long* nnz);
int pattern_structured_cg_q1(Launch& L, const GridDev& g, const SpaceDev& sp, long long** d_rowptr, int** d_colidx,
                             long long* nnz);

// ---- finite volumes (fv.cu) -----------------------------------------------------------------------
struct FvParams
{
  GridDev g;
  gdtb_flux flux;
  int ghosted;  // 1: vectors carry one ghost layer below and above the owned layers (multi-GPU slabs)
  int euler;    // 1: out = u - dt * L(u), 0: out = L(u)
  double dt;
  double lf_lambda_linear; // max_k |a_k| of a linear flux (1 / lambda of the Lax-Friedrichs flux)
  const double* ext[3];    // per-axis cell extents (device)
};
int launch_fv_apply(Launch& L, const FvParams& p, const double* u, double* out);
int launch_fv_interpolate(Launch& L, const GridDev& g, const FnDev& f, int m, const double* qx, const double* qw,
                          double* u);

} // namespace gdtb
