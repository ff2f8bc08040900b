// dune-gdt_b200/csrc/handles.hpp -- the opaque handle types of include/gdtb.h (internal; shared by capi.cu and solve.cu).
#pragma once

#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.hpp"

using namespace gdtb;

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct gdtb_ctx
{
  int device;
  cudaStream_t own_stream;
  Launch launch;
  Timing timing;
  int* d_error_flag;
  bool error_flag_pending; // an asynchronous generic assemble has not been checked yet
  // per-axis geometry tables of the grids seen so far (assemble_q1_gather.cu), keyed by the grid description
  struct AxisTables
  {
    GridDev grid;
    double* d_tab;
    long long offset[3], inv;
  };
  std::vector<AxisTables> axis_tables;
};

struct gdtb_grid
{
  gdtb_ctx* ctx;
  gdtb_grid_desc desc;
  GridDev dev;
};

struct gdtb_space
{
  gdtb_ctx* ctx;
  GridDev grid;
  SpaceDev dev;
};

struct gdtb_pattern
{
  gdtb_ctx* ctx;
  GridDev grid;
  SpaceDev test, ansatz;
  int stencil;
  long long rows, cols, nnz;
  long long* d_rowptr;
  int* d_colidx;
};

namespace gdtb {

struct LoweredForm
{
  gdtb_form form;                // cloned descriptor (data pointers replaced by device pointers)
  std::vector<double*> owned;    // device arrays cloned from host data
  // discrete functions (GDTB_FN_DOF_VECTOR): DoF array -> device-resident copy of the function's space description
  std::vector<std::pair<const double*, const SpaceDev*>> dof_spaces;
  int filter;
};

inline void free_form(LoweredForm& f)
{
  for (double* p : f.owned)
    cudaFree(p);
  f.owned.clear();
}

} // namespace gdtb

struct gdtb_matop
{
  gdtb_ctx* ctx;
  GridDev grid;
  SpaceDev test, ansatz;
  const gdtb_pattern* pattern;
  double* d_values;
  bool owns_values;
  std::vector<LoweredForm> element_forms, coupling_forms, boundary_forms;
  std::string plan;
  bool warned_generic = false; // GDTB_WARN_GENERIC: the slow-path note has been printed for this operator
  void* d_forms = nullptr; // lowered FormDev array of the DG gather path
  size_t d_forms_bytes = 0;
  std::vector<char> h_forms_cache; // what d_forms holds
  double* d_q2_tab = nullptr; // per-axis sum-factorisation tables of the CG Q2 gather path
  size_t d_q2_tab_bytes = 0;
  void* d_q1_items = nullptr; // work-item records of k_q1_gather<..., PREF> (kernels.hpp, Q1GatherParams::items)
  size_t d_q1_items_bytes = 0;
  long long q1_items_key[3] = {-1, -1, -1}; // (row_lo, row_hi, value_offset) the records were computed for
  void* d_q2_items = nullptr; // work-item records of the CG Q2 gather kernels (kernels.hpp, Q2GatherParams::items)
  size_t d_q2_items_bytes = 0;
  long long q2_items_key[3] = {-1, -1, -1}; // (layer_lo, layer_hi, items) the records were computed for
  double* d_qp_scratch = nullptr; // coefficient samples (one value / tensor per quadrature point) of the *_qp paths
  size_t d_qp_scratch_bytes = 0;
  // CSR pattern materialised on demand for the closed-form CG Q1 operator (Dirichlet constraints, SpMV, solvers)
  long long* d_own_rowptr = nullptr;
  int* d_own_colidx = nullptr;
  // owner-computes-rows slab (multi-GPU): only the rows [row_begin, row_end) live in d_values
  bool slab;
  long long row_begin, row_end;   // global row range held by this process
  long long value_offset, nnz_local;
  long long row_lo, row_hi, elem_lo, elem_hi; // vertex / element layers along the last direction
  bool halo = false; // interface-row halo partition: rows of one extra (interface) layer, own elements only
  // peer-memory halo (gdtb_halo_p2p_*): own receive buffers (two step parities) + flags, the neighbours' opened via CUDA IPC
  double* halo_recv = nullptr;
  int* halo_flags = nullptr;
  double* halo_peer_recv = nullptr; // upper neighbour's receive buffers
  int* halo_peer_flags = nullptr;   // upper neighbour's flags
  int* halo_lower_flags = nullptr;  // lower neighbour's flags
  bool halo_opened_lower = false, halo_opened_upper = false, halo_connected = false;
  long long halo_step = 0, halo_lower_layers = 0;
  // CG Q2 slabs: one row range per sub-entity group of the MCMG numbering (d_values holds them back to back)
  int n_ranges = 0;
  long long range_row_begin[8], range_row_end[8], range_value_offset[8], range_count[8];
};

struct gdtb_vecfun
{
  gdtb_ctx* ctx;
  GridDev grid;
  SpaceDev space;
  double* d_vec;
  bool owns_vec;
  std::vector<LoweredForm> forms;
  double* d_sep_tab; // separable right-hand-side tables for the gather kernel
  double* d_rule;    // qx | qw | phi for the table kernel
  bool slab;
  long long row_begin, row_end;
  long long row_lo, row_hi, elem_lo, elem_hi;
  // slabs of spaces whose owned rows are not one range (CG Q2: one range per sub-entity group) or that take the
  // quadrature-faithful kernel: global row ranges and their positions in the local vector (n_ranges == 0: [row_begin,
  // row_end) at position 0)
  int n_ranges = 0;
  long long range_row_begin[8], range_row_end[8], range_local[8];
  long long local_size = 0;
  double h_rule[4 * MAX_Q1D];
  bool rule_uploaded;
  bool halo = false;
  // what d_sep_tab was built for
  bool sep_valid = false;
  FnDev sep_fn;
  int sep_m = 0;
  long long sep_lo = 0, sep_hi = 0;
};

struct gdtb_fvop
{
  gdtb_ctx* ctx;
  GridDev grid;
  SpaceDev space;
  gdtb_flux flux;
  bool ghosted;
  double* d_tmp; // ping-pong buffer for the Euler loop
  double* d_src; // staging for the *_host entry points
  double* d_dst;
  double* d_ext; // per-axis cell extents, axis k at d_ext + ext_offset[k]; reciprocals inv_ext_shift further on
  long long ext_offset[3];
  long long inv_ext_shift;
  int rows_per_block; // tuning knob of the marching kernel (0 = automatic), GDTB_FV_ROWS in the environment
  // boundary treatments resolved per domain side (gdtb_fvop_append_boundary)
  unsigned bnd_ext_mask = 0, bnd_nf_mask = 0;
  double bnd_ext_a[6] = {0, 0, 0, 0, 0, 0}, bnd_ext_b[6] = {0, 0, 0, 0, 0, 0};
  double bnd_nf_a[6] = {0, 0, 0, 0, 0, 0}, bnd_nf_b[6] = {0, 0, 0, 0, 0, 0};
  double* d_partial = nullptr; // block partials of the dt estimate
  // peer-memory ghost exchange (gdtb_fvop_p2p_*): own slab vectors + flags, the neighbours' opened through CUDA IPC
  double* p2p_u[2] = {nullptr, nullptr};
  int* p2p_flags = nullptr;
  double* p2p_peer_u[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; // [lower / upper neighbour][buffer]
  int* p2p_peer_flags[2] = {nullptr, nullptr};
  long long p2p_peer_layers[2] = {0, 0};
  bool p2p_opened[2] = {false, false};
  long long p2p_step = 0;
};

struct gdtb_rk
{
  gdtb_fvop* op;
  int s;
  double A[GDTB_RK_MAX_STAGES * GDTB_RK_MAX_STAGES], b[GDTB_RK_MAX_STAGES], c[GDTB_RK_MAX_STAGES];
  double r, t;
  double* d_ui = nullptr;                    // stage vector u_i (Euler: ping-pong buffer)
  double* d_k[GDTB_RK_MAX_STAGES] = {};      // stages k_i
  // slab mode (operator with gdtb_fvop_set_slab): solution and stage vectors are stepper-owned, exported through CUDA
  // IPC, and their boundary layers are handed to the neighbours by peer stores (gdtb_rk_p2p_*)
  bool slab = false;
  double* p2p_un = nullptr;
  double* p2p_ui[2] = {nullptr, nullptr};
  int* p2p_flags = nullptr; // [0] lower ghost filled, [1] upper ghost filled, [2] timeout, [4] edge counter
  double* peer_un[2] = {nullptr, nullptr};      // [lower / upper neighbour]
  double* peer_ui[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  int* peer_flags[2] = {nullptr, nullptr};
  long long peer_layers[2] = {0, 0};
  bool peer_opened[2] = {false, false};
  bool has_peer[2] = {false, false};
  long long sends = 0; // hand-overs issued so far (= value the neighbours' counters must have reached)
  int ui_parity = 0;
};

namespace gdtb {
// host numerics of capi.cu: Gauss-Legendre rules on [0,1] ([EXT] dune-geometry), 1D Lagrange tables
int gauss_points_for_order(int order);
void gauss_legendre_01(int m, double* x, double* w);
void lagrange_1d(int K, double x, double* v, double* dv);
int internal_check_ctx(gdtb_ctx* ctx);
int internal_validate_function(const gdtb_function& f, const char* what);
int internal_lower_function(gdtb_ctx* ctx, const GridDev& g, gdtb_function& f, LoweredForm& owner);
FnDev internal_to_dev(const gdtb_function& f, const LoweredForm* owner = nullptr);
// the CSR pattern an operator's values follow (solve.cu): the caller's, or the closed-form one materialised on demand
int internal_matop_pattern(gdtb_matop* op, const long long** rowptr, const int** colidx);
} // namespace gdtb
