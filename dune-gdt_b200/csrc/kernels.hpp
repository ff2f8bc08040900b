// dune-gdt_b200/csrc/kernels.hpp -- host-callable launchers of the CUDA kernels (internal).
#pragma once

#include <string>
#include <vector>

#include "common.cuh"

namespace gdtb {

// kernel families for the optional per-kernel CUDA-event timing (gdtb_ctx_enable_timing)
enum KernelFamily
{
  KF_Q1_GATHER = 0,
  KF_Q2_GATHER,
  KF_DG_GATHER,
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
  // (mangled) symbol of the kernel instantiation each family launched last (gdtb_ctx_kernel_name)
  std::string last_kernel[KF_COUNT];
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

// remembers which instantiation a family launched (cudaFuncGetName: the symbol as it appears in ncu launch lists)
inline void note_kernel(Launch& L, int family, const void* kernel)
{
  if (L.timing) {
    const char* name = nullptr;
    if (cudaFuncGetName(&name, kernel) == cudaSuccess && name)
      L.timing->last_kernel[family] = name;
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
// owned global row ranges -> positions in a slab-local vector (n == 0: the vector is global)
struct RowMap
{
  int n;
  long long begin[8], end[8], local[8];
};
int launch_element_vector(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, double* vec,
                          const RowMap& rows);
int launch_coupling_matrix(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, int filter,
                           const long long* rowptr, const int* colidx, double* values, int* error_flag);
int launch_boundary_matrix(Launch& L, const GridDev& g, const SpaceDev& sp, const FormDev& f, const long long* rowptr,
                           const int* colidx, double* values, int* error_flag);

// ---- CG-Q1 row-gather assembly (assemble_q1_gather.cu) --------------------------------------------
constexpr int Q1G_MAX_GROUPS = 3;

enum
{
  Q1G_LAPLACE_SCALAR = 0, // LocalLaplaceIntegrand with kappa = c * I (constant c or one value per element)
  Q1G_LAPLACE_TENSOR = 1, // LocalLaplaceIntegrand with a full d x d kappa (constant or one tensor per element)
  Q1G_MASS = 2            // LocalElementProductIntegrand with a scalar weight (constant or per element)
};

// One "group" = one integrand of one appended local form.  On an affine, axis-aligned cell e with an element-wise
// constant coefficient the quadrature sum of integrals.hh:116-132 factorises exactly into
//   L_e[i][j] = sum_{r,c} kappa_rc(e) * (1/h_r(e)) * (1/h_c(e)) * |det J_e| * M[r][c][i][j],
//   M[r][c][i][j] = sum_q w_q d_r phihat_i(xhat_q) d_c phihat_j(xhat_q)      (reference element, the form's own rule)
// (mass: L_e[i][j] = w(e) * |det J_e| * sum_q w_q phihat_i phihat_j).  M is tabulated at plan time; the geometry
// factors h_k(e), 1/h_k(e), |det J_e| and the coefficient are evaluated per element on the device, in FP64.
struct Q1Group
{
  int kind;
  int coef_elem;      // 1: coefficient array indexed by element (scalar, or d*d row-major for the tensor kind)
  double scale;       // MatrixOperator::scaling at append time (* the constant scalar coefficient)
  double kappa[9];    // constant tensor, row-major with leading dimension 3 (Q1G_LAPLACE_TENSOR, coef_elem == 0)
  const double* coef; // device array
  double M[9][8][8];  // M[r*3+c][o][s]: o = offset of the element around the vertex (test index i = (2^d-1) ^ o),
                      // s = ansatz index; the scalar kinds only use r == c, the mass kind only M[0]
  double K1[2][2], M1[2][2]; // 1D reference stiffness / mass tables of the form's rule (sum-factorised constant path)
};

// exact unsigned division v / d for v < 2^31: (v * magic) >> shift
struct FastDiv
{
  unsigned d, magic, shift;
};

inline FastDiv make_fast_div(unsigned d)
{
  FastDiv f;
  f.d = d;
  unsigned l = 0;
  while ((1ull << l) < d)
    ++l;
  f.shift = 31 + l;
  f.magic = (unsigned)(((1ull << f.shift) + d - 1) / d); // <= 2^32 - 1 for d >= 2; d == 1: 2^31
  return f;
}

// Interface-row halo handed over INSIDE the gather kernel (multi-GPU, one process per GPU, buffers opened through CUDA
// IPC): the slab walks its own elements only; the rows of its top vertex layer are partial sums that belong to the
// slab above -- the work items of that layer are computed first and stored straight into the neighbour's receive
// buffer (NVLink peer stores) followed by a counter increment in the neighbour's memory; the items of the bottom layer
// are computed last, wait for the lower neighbour's counter and add what arrived before their rows leave the SM.
struct Q1HaloP2p
{
  int has_lower, has_upper;
  long long layer_rows, layer_values; // one interface layer: vertices, CSR values
  double* peer_values;                // upper neighbour's receive buffer (this step's parity)
  double* peer_rhs;
  int* peer_flags;                    // upper neighbour: [0] data counter
  const double* recv_values;          // own receive buffer (this step's parity)
  const double* recv_rhs;
  int* my_flags;                      // [0] data counter (raised by the lower neighbour), [1] acknowledgements (by the upper), [2] timeout
  int* lower_flags;                   // lower neighbour's flags: [1] is raised when a bottom item has consumed its data
  int expect_data;                    // data counter value that marks this step's layer as complete
  int expect_ack;                     // acknowledgements needed before this step's parity buffer may be overwritten
};

struct Q1GatherParams
{
  GridDev g;
  FastDiv div_vx, div_vy;     // division by the number of vertices per x-line / per y-line
  const double* axis_tab[3];  // per-axis geometry tables (k_q1_axis_tables): [h | 1/h], entry i + 1 = cell i
  long long axis_tab_inv;     // offset of the 1/h half
  int n_groups;
  Q1Group group[Q1G_MAX_GROUPS];
  // right-hand side: b[v] = sum_o valid(o) * |det J_e| * (rhs_S_const[o] + rhs_S_elem[o] * f[e])
  //                         + rhs_sep_scale * prod_k B_k[i_k]
  int has_rhs;
  int rhs_has_const;
  int rhs_has_elem;
  int rhs_has_sep;
  double rhs_S_const[8];     // sum over forms of w * c * prod_k s1[a_k(o)], s1[a] = sum_q w_q phihat_a(x_q)
  double rhs_S_elem[8];      // w * prod_k s1[a_k(o)]
  const double* rhs_elem;    // per-element source values
  double rhs_sep_scale;      // w * p0
  const double* rhs_sep_tab; // 3 tables B_k[i_k] (they carry the cells' extents), k-th table at offset k * stride
  long long rhs_sep_stride;
  long long value_offset; // global CSR position of this process' first row (values points at it)
  long long row_offset;   // first vertex row of this process
  // owner-computes-rows slab along the last direction: vertex layers [row_lo, row_hi) are produced here from the
  // element layers [elem_lo, elem_hi) (the owned ones plus the ghost layer below)
  long long row_lo, row_hi;
  long long elem_lo, elem_hi;
  int halo_p2p; // the interface-row halo travels inside the kernel (halo below)
  int no_sf3;   // A/B knob (GDTB_Q1_NO_SF3): constant kappa through the per-cell sum instead of the sum-factorised stencil
  // work-item records of the kernel variant with one kappa per element (k_q1_items; 32 bytes per item of Q1G_ROWS_PREF
  // rows, q1_pref_item_capacity() items): optional caller-owned device buffer; items_ready: it already holds the
  // records of this grid / slab (set by the launcher when it wrote them)
  void* items;
  int items_ready;
  long long halo_top_value_start; // local position of the first value of the top (interface) layer
  Q1HaloP2p halo;
};

int launch_q1_gather(Launch& L, const Q1GatherParams& p, double* values, double* rhs, bool accumulate);
// work items (chunks of Q1G_ROWS rows) that touch the first / last vertex layer of a slab holding `layers` layers of
// `layer_rows` rows: the number of counter increments per step in the peer-memory halo
int q1_halo_items(long long layer_rows, long long layers, bool top);
int launch_q1_axis_tables(Launch& L, const GridDev& g, double* const* tabs, long long inv);

// builds the separable right-hand-side tables B_k[i_k] for a product-separable built-in source
// upper bound of the work items of k_q1_gather<..., PREF> for the rows [row_lo, row_hi) of the grid
long long q1_pref_item_capacity(const GridDev& g, long long row_lo, long long row_hi);
int launch_q1_rhs_tables(Launch& L, const GridDev& g, long long elem_lo, long long elem_hi, const FnDev& f, int m,
                         const double* qx, const double* qw, const double* phi /* [m][2] */, double* tab,
                         long long stride);

// ---- CG row gather with coefficients that vary inside a cell -----------------------------------------------------
// LocalLaplaceIntegrand / LocalElementProductIntegrand with an arbitrary GridFunction (laplace.hh:40-48, product.hh:
// 56-65): the coefficient arrives as one value (or one d x d tensor) per quadrature point of the form's own rule --
// caller-sampled (GDTB_FN_QP_*) or sampled on the device from an analytic / discrete function (k_sample_function).
// The quadrature sum of integrals.hh:116-132 is evaluated per element by sum factorisation over the tensor rule:
//   L_e[i][j] = scale |det J_e| sum_{r,c} (1/h_r)(1/h_c) sum_q kappa_rc(x_q) prod_k PT^{(k==r, k==c)}[q_k][i_k][j_k],
//   PT^{(ta,tb)}[q][a][b] = w_q D^ta phi_a(x_q) D^tb phi_b(x_q)   (1D point tables, a = test, b = ansatz)
// (mass: PT^{(0,0)} along every axis), one integrand per launch, rows owned and written once like in the
// constant-coefficient gather kernels.
enum
{
  QPT_MM = 0, // phi_a  phi_b
  QPT_KK = 1, // phi_a' phi_b'
  QPT_KM = 2, // phi_a' phi_b   (test derivative)
  QPT_MK = 3  // phi_a  phi_b'  (ansatz derivative)
};

struct CgQpGroup
{
  int kind;           // Q1G_LAPLACE_SCALAR, Q1G_LAPLACE_TENSOR or Q1G_MASS
  int m;              // Gauss points per direction
  double scale;       // MatrixOperator::scaling at append time
  const double* coef; // device array [e][q] (scalar kinds) or [e][q][d * d]; e = element index of the grid view
  double pt[4][MAX_Q1D][3][3];
};

struct Q1QpParams
{
  GridDev g;
  FastDiv div_vx, div_vy;
  const double* axis_tab[3];
  long long axis_tab_inv;
  CgQpGroup group;
  long long value_offset, row_offset;
  long long row_lo, row_hi, elem_lo, elem_hi;
};
bool q1_qp_supported(int d, int m, int kind);
int launch_q1_gather_qp(Launch& L, const Q1QpParams& p, double* values, bool accumulate);

// samples a grid function at the quadrature points of the tensor rule with m points per direction: out[(e - e_begin) *
// m^d + q] (scalar) or out[((e - e_begin) * m^d + q) * d * d + r * d + c] (tensor = 1: scalar functions mean c * I) for the
// elements e_begin <= e < e_end
int launch_sample_function(Launch& L, const GridDev& g, const FnDev& f, int m, const double* qx /* host, m points */,
                           int tensor, long long e_begin, long long e_end, double* out);

// ---- CG-Q2 row-gather assembly (assemble_q2_gather.cu) --------------------------------------------
constexpr int Q2G_MAX_GROUPS = 3;

// one integrand of one appended form (kinds: Q1G_LAPLACE_SCALAR, Q1G_MASS); TM / TK are the 1D reference mass and
// stiffness tables of the Q2 shape functions (nodes 0, 1/2, 1) for the form's own Gauss rule
struct Q2Group
{
  int kind;
  int coef_elem;
  double scale;
  const double* coef;
  double TM[3][3];
  double TK[3][3];
};

// rows of one sub-entity kind (parity pattern s of the lattice point): contiguous in the global numbering
struct Q2RowGroup
{
  int s;
  long long row_begin, rows, item_begin;
  // CSR placement (closed form, see q2_row_offset): the value of row `lex` of the group lives at
  // value_begin + q2_row_offset(lex) of the (slab-local) value buffer; [lex_begin, lex_end) are the rows to produce,
  // off_end = q2_row_offset(lex_end)
  long long value_begin, off_end;
  long long lex_begin, lex_end;
  // row decode and row-start arithmetic of the group: extents, their division magics floor(2^64 / e) + 1, and the
  // per-axis entry totals Tx, Tx * Ty
  unsigned ex, ey;
  unsigned long long mex, mey;
  unsigned Tx;
  long long TxTy;
};

struct Q2GatherParams
{
  GridDev g;
  int n_groups;
  Q2Group group[Q2G_MAX_GROUPS];
  int n_rowgroups;
  Q2RowGroup rg[8];
  long long n_items;
  const long long* rowptr; // device CSR row pointer of the element pattern
  // sum-factorised path (all coefficients constant): per group and axis, for every lattice point p in [0, 2 N_k] the
  // 1D row vectors K[0..5) = sum_e K1[i_e(p)][.] / h_e and M[0..5) = sum_e M1[i_e(p)][.] h_e over the <= 2 elements
  // that contain p, indexed by box offset; layout tab[group][sf_axis_off[k] + 10 p + {K: 0..4, M: 5..9}]
  int sf; // 0: off, 1: all coefficients constant (summed tables), 2: ONE integrand with one coefficient per element
  const double* sf_tab;
  long long sf_axis_off[3], sf_group_stride;
  CgQpGroup qp; // launch_q2_gather_qp: the one integrand with a coefficient per quadrature point
  // work-item records (optional, 32 bytes per item, q2_item_count() items): the uniform per-item bookkeeping (row group,
  // first row, CSR segment, the two lattice lines the item touches) computed once by k_q2_items instead of by every warp
  // of the gather kernel; items_ready: the buffer already holds the records of this grid / slab
  void* items;
  int items_ready;
};
// number of work items of the CG Q2 gather kernels for this grid / slab, 0 if the item records do not apply (a lattice
// line shorter than a work item)
long long q2_item_count(const GridDev& g, const SpaceDev& sp);

long long q2_sf_table_doubles(const GridDev& g); // doubles per group
long long q2_pe_table_doubles(const GridDev& g); // sf == 2: per-element 1D factors of a single integrand
// Row ranges a slab of element layers [g.layer_lo, g.layer_hi) owns (owner-computes-rows: a lattice layer belongs to
// the slab of the element layer above it, the top layer to the last slab): one contiguous range per sub-entity group
// of the MCMG numbering, with the global CSR offset of its first value and its position in the slab-local buffer
struct Q2SlabRange
{
  long long row_begin, row_end, value_offset, local_offset, count;
};
int q2_slab_ranges(const GridDev& g, const SpaceDev& sp, Q2SlabRange* out /* [8] */);
int launch_q2_gather(Launch& L, Q2GatherParams& p, const SpaceDev& sp, double* values, bool accumulate);
// the same row / plane decomposition with ONE integrand whose coefficient is given per quadrature point
bool q2_qp_supported(int d, int m, int kind);
int launch_q2_gather_qp(Launch& L, Q2GatherParams& p, const CgQpGroup& group, const SpaceDev& sp, double* values,
                        bool accumulate);
// 3D, scalar coefficient: the x-fused kernel (assemble_q2_qp.cu; three local-matrix rows per element and thread, the
// coefficient stream staged in shared memory by TMA bulk loads); [coef_e_begin, coef_e_end) = elements the coefficient
// array holds (indexed by the element index of the grid view)
bool q2_qp_xfused_supported(int d, int m, int kind);
int launch_q2_qp_xfused(Launch& L, const GridDev& g, const CgQpGroup& group, const SpaceDev& sp, long long coef_e_begin,
                        long long coef_e_end, double* values, bool accumulate);

// ---- DG row-gather assembly (assemble_dg_gather.cu) -----------------------------------------------
constexpr int DGG_THREADS = 128;
constexpr int DGG_MAX_FORMS = 8;

struct DgGatherParams
{
  GridDev g;
  SpaceDev sp;
  const FormDev* forms; // device array: element forms, then coupling forms, then boundary forms
  int n_elem, n_coup, n_bnd;
  const long long* rowptr; // device CSR row pointer of the element_and_intersection pattern
  // factorised path (order 1, every coefficient a constant or element-wise scalar): the quadrature sums of all forms
  // collapse into 1D tables; row starts are closed forms (rowptr is not read)
  int fast; // 0: quadrature-faithful kernel, 1: factorised, 2: factorised with constant coefficients tabulated
  // the operator is exactly the reference drivers' SWIPDG one: one element form {Laplace}, one coupling form {inner
  // coupling, inner penalty}, one boundary form {Dirichlet coupling, boundary penalty} -- lets the factorised kernel with
  // element-wise coefficients run with compile-time term loops
  int swip;
  // ... and then carries that operator's coefficients and scalars here (constant bank) instead of in the FormDev array in
  // global memory, which the general kernel re-reads for every face and term
  struct SwFn
  {
    const double* data; // element-wise scalar array, or nullptr: the constant c
    double c;
  };
  struct SwDesc
  {
    SwFn elem_kappa, coup_kappa, coup_weight, pen_weight, bnd_kappa, bnd_weight;
    double coup_prefactor, pen_prefactor, bnd_prefactor, bndpen_prefactor;
    double s_elem, s_coup, s_bnd; // the forms' scalings
    int pen_hI, bndpen_hI;
    int coup_same; // coup_kappa, coup_weight and pen_weight are the same function (kappa = omega): one load per cell
  } sw;
  unsigned long long magic[2]; // floor(2^64 / n_k) + 1 for the element-index decode (0 when n_k == 1)
  // element-owned rows: this process produces the rows of the elements [e_begin, e_end) (a slab of element layers),
  // `values` starts at the global CSR position value_offset
  long long e_begin, e_end, value_offset;
  // bit f: coupling form f also runs over the periodic wrap faces (GDTB_FILTER_INNER_AND_PERIODIC_ONCE)
  unsigned coup_on_periodic;
  // factorised kernels: per-axis geometry tables of the grid (k_q1_axis_tables: [h | 1/h], entry i + 1 = cell i)
  const double* axis_tab[3];
  long long axis_tab_inv;
};

bool dg_gather_supported(int d, int K);
bool dg_gather_fast_supported(const GridDev& g, int K);
int launch_dg_gather(Launch& L, const DgGatherParams& p, double* values, bool accumulate);

// ---- sparsity pattern (pattern.cu) ----------------------------------------------------------------
int pattern_sort_unique(Launch& L, const GridDev& g, const SpaceDev& test, const SpaceDev& ansatz, int stencil,
                        long long** d_rowptr, int** d_colidx, long long* nnz);
int pattern_structured_cg_q1(Launch& L, const GridDev& g, const SpaceDev& sp, long long** d_rowptr, int** d_colidx,
                             long long* nnz);
// closed-form generators of the other two gather paths (same CSR the sort-and-unique builder produces):
// CG Q2 element stencil on the lattice numbering (assemble_q2_gather.cu), DG element_and_intersection stencil on a
// non-periodic grid (assemble_dg_gather.cu)
int pattern_structured_cg_q2(Launch& L, const GridDev& g, const SpaceDev& sp, long long** d_rowptr, int** d_colidx,
                             long long* nnz);
int pattern_structured_dg(Launch& L, const GridDev& g, const SpaceDev& sp, long long** d_rowptr, int** d_colidx,
                          long long* nnz);
// the same closed-form row pointers evaluated on the host (rowptr[0 .. size]); no device involved
int q1_host_rowptr(const GridDev& g, const SpaceDev& sp, long long* rowptr);
int q2_host_rowptr(const GridDev& g, const SpaceDev& sp, long long* rowptr);
int dg_host_rowptr(const GridDev& g, const SpaceDev& sp, long long* rowptr);
// global CSR position of the first value of element e in the DG element_and_intersection stencil (e == ne: nnz)
long long dg_value_offset(const GridDev& g, const SpaceDev& sp, long long e);

// ---- finite volumes (fv.cu) -----------------------------------------------------------------------
struct FvParams
{
  GridDev g;
  gdtb_flux flux;
  int ghosted;  // 1: vectors carry one ghost layer below and above the owned layers (multi-GPU slabs)
  int euler;    // 1: out = u - dt * L(u), 0: out = L(u)
  double dt;
  double lf_lambda_linear; // max_k |a_k| of a linear flux (1 / lambda of the Lax-Friedrichs flux)
  const double* ext[3];     // per-axis cell extents (device)
  const double* inv_ext[3]; // 1 / ext
  int rows_per_block;       // marching kernel: layers per thread block (0 = choose)
  // layers [apply_lo, apply_hi) of the slab [g.layer_lo, g.layer_hi) to produce in this launch (overlap of the
  // interior with the ghost-layer exchange); the memory layout always follows the slab
  long long apply_lo, apply_hi;
  // boundary treatments resolved per domain side (2k+s): extrapolation v = a u + b through the numerical flux, and/or
  // numerical boundary flux g = a (f(u) . n) + b (local/operators/advection-fv.hh:188-457)
  unsigned bnd_ext_mask, bnd_nf_mask;
  double bnd_ext_a[6], bnd_ext_b[6], bnd_nf_a[6], bnd_nf_b[6];
  // peer-memory ghost exchange (multi-GPU slabs, one process per GPU, buffers opened through CUDA IPC): the kernel
  // stores the first / last owned layer of its result straight into the neighbours' ghost layers over NVLink and
  // raises a counter in the neighbour's memory; the CTAs that read a ghost layer first wait for the neighbour's
  // counter of the previous step.  NULL pointers: no neighbour on that side.
  int p2p;
  int wait_lo, wait_hi;  // a neighbour fills the lower / upper ghost layer of the source: wait for its counter
  double* peer_lo_ghost; // upper ghost layer of the lower neighbour's destination buffer
  double* peer_hi_ghost; // lower ghost layer of the upper neighbour's destination buffer
  int* peer_lo_flag;     // lower neighbour's "upper ghost filled" counter
  int* peer_hi_flag;     // upper neighbour's "lower ghost filled" counter
  const int* my_flags;   // {lower ghost filled, upper ghost filled} counters of this rank
  int expect;            // counter value that marks the ghost layers of the source buffer as complete (= step index)
  int* edge_count;       // {lower, upper}: edge blocks of this launch that have handed their layer over (last one signals)
  int* timeout_flag;     // set when a wait gave up (bounded spin)
  // Runge-Kutta stage combination fused into the apply (tools/timestepper/explicit-rungekutta.hh:248-263): the operator
  // is applied to u_i = u + sum_j stage_v[j] * stage_c[j], formed on the fly at every load (same order of additions as
  // the separate axpy pass, so the same roundings) -- the stage vector is never written.  out_mode 1: the last stage
  // writes the step's result u + sum_j out_v[j] * out_c[j] + L(u_i) * out_cL instead of k_i = L(u_i).
  int n_stage;
  const double* stage_v[3];
  double stage_c[3];
  int out_mode, n_out;
  const double* out_v[3];
  double out_c[3], out_cL;
};
int launch_fv_apply(Launch& L, const FvParams& p, const double* u, double* out);

// explicit Runge-Kutta vector updates (tools/timestepper/explicit-rungekutta.hh:248-263):
// out = base + sum_j v[j] * c[j], terms added in order; nv <= RK_MAX_TERMS per launch
constexpr int RK_MAX_TERMS = 4;
struct RkAxpyParams
{
  long long n;
  int nv;
  const double* v[RK_MAX_TERMS];
  double c[RK_MAX_TERMS];
};
int launch_rk_axpy(Launch& L, const RkAxpyParams& p, const double* base, double* out);
// estimate_dt_for_hyperbolic_system (tools/hyperbolic.hh:38-86): per block {min u, max u, max perimeter / volume}
// hands the first / last owned layer of a slab vector over to the neighbours' ghost layers (peer stores) and raises their
// step counters: the exchange of the Runge-Kutta stage vectors (tools/timestepper/explicit-rungekutta.hh:252-257)
struct P2pSendParams
{
  const double* src;   // slab vector [ghost | owned | ghost]
  long long plane;     // cells per layer
  long long layers;    // owned layers
  double* peer_lo_ghost;
  double* peer_hi_ghost;
  int* peer_lo_flag;
  int* peer_hi_flag;
  int* edge_count;
};
int launch_p2p_send_layers(Launch& L, const P2pSendParams& p);
int launch_fv_dt_reduce(Launch& L, const FvParams& p, const double* u, double* partial /* 3 * blocks */, int blocks);
int launch_fv_interpolate(Launch& L, const GridDev& g, const FnDev& f, int m, const double* qx, const double* qw,
                          double* u);

} // namespace gdtb
