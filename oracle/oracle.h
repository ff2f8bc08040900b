/*
 * oracle/oracle.h -- C interface of the CPU restatement oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dune-gdt_b200/ (the product) may include, link or load
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the timed CPU baseline.
 *
 * The oracle restates, for structured YaspGrid-style cube grids, the loops of
 *   dune/gdt/tools/sparsity-pattern.hh, dune/gdt/spaces/mapper/{continuous,discontinuous,finite-volume}.hh,
 *   dune/gdt/spaces/basis/default.hh, dune/gdt/local/bilinear-forms/integrals.hh,
 *   dune/gdt/local/functionals/integrals.hh,
 *   dune/gdt/local/integrands/{laplace,product,conversion,combined,ipdg,laplace-ipdg}.hh,
 *   dune/gdt/local/assembler/{bilinear-form-assemblers,functional-assemblers,operator-applicators}.hh,
 *   dune/gdt/operators/{matrix-based,localizable-operator,advection-fv}.hh,
 *   dune/gdt/local/operators/advection-fv.hh, dune/gdt/local/numerical-fluxes/upwind.hh
 * (all paths relative to the reference checkout).  The struct layouts below are deliberately
 * identical to the descriptor structs in include/gdtb.h so one ctypes definition feeds both.
 *
 * PARITY STATUS: values are pinned by the reference's own known-answer tests
 * (dune/gdt/test/integrands/integrands_laplace.cc:131-133, integrands_product.cc:122-124,
 * dune/gdt/test/linear-transport/linear_transport__1d__explicit__fv.mini:8-14,
 * dune/gdt/test/burgers/burgers__1d__explicit__fv.mini:9-15,
 * dune/gdt/test/stationary-heat-equation/stationary_heat_equation__ESV2007__table_1.mini:23-36),
 * see tests/test_oracle_golden.py.
 * "PARITY UNPINNED" for everything that lives in third-party code absent from the reference tree
 * (dune-grid / dune-geometry / dune-localfunctions / dune-xt): the concrete Q2 global DoF numbering
 * (MCMG mapper offsets + YaspGrid sub-entity indices), the quadrature point order, the local DoF
 * order of the Lagrange elements and the walker's traversal order.  No reference test pins those.
 */
#ifndef GDTB_ORACLE_H
#define GDTB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- descriptors (layout == include/gdtb.h) ------------------------------------------------ */

typedef struct orc_grid
{
  int32_t dim;          /* 1, 2 or 3 */
  int32_t periodic;     /* bit k set: periodic in direction k (XT::Grid::make_periodic_grid_view) */
  double lower[3];
  double upper[3];
  int64_t n[3];         /* elements per direction */
} orc_grid;

enum
{
  ORC_SPACE_CG = 0, /* ContinuousLagrangeSpace     */
  ORC_SPACE_DG = 1, /* DiscontinuousLagrangeSpace  */
  ORC_SPACE_FV = 2  /* FiniteVolumeSpace (order 0) */
};

enum
{
  ORC_STENCIL_ELEMENT = 0,
  ORC_STENCIL_INTERSECTION = 1,
  ORC_STENCIL_ELEMENT_AND_INTERSECTION = 2
};

enum
{
  ORC_FN_CONST_SCALAR = 0, /* c[0] (as a d x d function: c[0] * I, laplace.hh:41)   */
  ORC_FN_CONST_TENSOR = 1, /* c[0..d*d) row-major                                   */
  ORC_FN_ELEM_SCALAR = 2,  /* data[e]                                               */
  ORC_FN_ELEM_TENSOR = 3,  /* data[e*d*d + r*d + c]                                 */
  ORC_FN_BUILTIN = 4,      /* analytic scalar function of the global coordinate     */
  ORC_FN_QP_SCALAR = 5,    /* data[e * qp_per_element + q], sampled at the form's own rule (element forms)      */
  ORC_FN_QP_TENSOR = 6,    /* data[(e * qp_per_element + q) * d * d + r * d + c]                                */
  ORC_FN_DOF_VECTOR = 7    /* discrete function of the space (space_kind, space_order), DoF vector in data      */
};

enum
{
  ORC_BUILTIN_COS_PRODUCT = 1, /* p0 * prod_i cos(p1 * x_i)                   */
  ORC_BUILTIN_AFFINE = 2,      /* p0 + sum_i p[1+i] * x_i                     */
  ORC_BUILTIN_GAUSSIAN = 3,    /* exp(-(x_0 - p0)^2 / (2 p1^2))               */
  ORC_BUILTIN_INDICATOR = 4,   /* p0 <= x_0 <= p1 ? 1 : 0                     */
  ORC_BUILTIN_QUADRATIC = 5    /* p0 + p1 * sum_i x_i^2                       */
};

typedef struct orc_function
{
  int32_t kind;
  int32_t order; /* declared polynomial order, enters the quadrature order like GridFunction::order() */
  int32_t builtin;
  int32_t reserved;
  double c[9];
  double p[8];
  const double* data;
  int32_t qp_per_element;
  int32_t space_kind;
  int32_t space_order;
  int32_t reserved2;
} orc_function;

enum
{
  ORC_INT_LAPLACE = 0,                 /* LocalLaplaceIntegrand                         */
  ORC_INT_PRODUCT = 1,                 /* LocalElementProductIntegrand                  */
  ORC_INT_IPDG_INNER_COUPLING = 2,     /* LocalLaplaceIPDGIntegrands::InnerCoupling     */
  ORC_INT_IPDG_INNER_PENALTY = 3,      /* LocalIPDGIntegrands::InnerPenalty             */
  ORC_INT_IPDG_DIRICHLET_COUPLING = 4, /* LocalLaplaceIPDGIntegrands::DirichletCoupling */
  ORC_INT_IPDG_BOUNDARY_PENALTY = 5    /* LocalIPDGIntegrands::BoundaryPenalty          */
};

enum
{
  ORC_HI_DIAMETER = 0, /* default_intersection_diameter (ipdg.hh:27-38) */
  ORC_HI_VOLUME = 1    /* |I| (ESV2007.hh:108-112)                      */
};

typedef struct orc_integrand
{
  int32_t kind;
  int32_t hI_kind;
  double prefactor;       /* symmetry prefactor (couplings) / penalty sigma (penalties) */
  orc_function diffusion; /* kappa (Laplace, couplings) or the weight of the product integrand */
  orc_function weight;    /* omega of the IPDG integrands */
} orc_integrand;

#define ORC_MAX_TERMS 4

typedef struct orc_form
{
  int32_t n_terms; /* integrand_a + integrand_b + ... (combined.hh) */
  int32_t over_integrate;
  double scaling; /* matrixoperator.scaling captured at append (matrix-based.hh:342,365) */
  orc_integrand terms[ORC_MAX_TERMS];
} orc_form;

enum
{
  ORC_FLUX_LINEAR = 0,  /* f(u) = a * u, a = p[0..d)          */
  ORC_FLUX_BURGERS = 1, /* f(u) = 0.5 u^2 * (1,...,1)         */
  ORC_FLUX_EULER = 2    /* EulerTools<d>::flux, p[0] = gamma (tools/euler.hh), m = d + 2 */
};

enum
{
  ORC_NUMFLUX_UPWIND = 0,
  ORC_NUMFLUX_LAX_FRIEDRICHS = 1,
  ORC_NUMFLUX_VIJAYASUNDARAM = 2 /* local/numerical-fluxes/vijayasundaram.hh:111-133 (systems) */
};

typedef struct orc_flux
{
  int32_t kind;
  int32_t numflux;
  double p[4];
} orc_flux;

/* boundary treatments of the FV operator (operators/advection-fv.hh:96-123, local/operators/advection-fv.hh:188-457);
 * side_mask bit (2k+s): domain face with outer normal -e_k (s = 0) / +e_k (s = 1) */
enum
{
  ORC_FVBND_EXTRAPOLATION = 0, /* ...ByCustomExtrapolationOperator with v = a u + b           */
  ORC_FVBND_NUMERICAL_FLUX = 1 /* ...ByCustomNumericalFluxOperator with g = a (f(u) . n) + b  */
};
typedef struct orc_fv_boundary
{
  int32_t kind;
  uint32_t side_mask;
  double a, b;
} orc_fv_boundary;

/* ---- spaces / mappers ---------------------------------------------------------------------- */

int64_t orc_num_elements(const orc_grid* g);
int64_t orc_space_size(const orc_grid* g, int kind, int order);
int32_t orc_space_local_size(const orc_grid* g, int kind, int order);
void orc_space_global_indices(const orc_grid* g, int kind, int order, int64_t element, int64_t* out);

/* ---- quadrature / shape functions (exposed for the unit tests) ------------------------------ */

int32_t orc_gauss_rule(int order, double* points01, double* weights); /* returns m */
void orc_shape_values(int dim, int order, const double* xhat, double* values);
void orc_shape_gradients(int dim, int order, const double* xhat, double* grads /* n*dim */);

/* ---- sparsity pattern (tools/sparsity-pattern.hh) ------------------------------------------- */

typedef struct orc_pattern orc_pattern;
orc_pattern* orc_pattern_create(const orc_grid* g, int test_kind, int test_order, int ansatz_kind, int ansatz_order,
                                int stencil);
int64_t orc_pattern_rows(const orc_pattern* p);
int64_t orc_pattern_nnz(const orc_pattern* p);
void orc_pattern_copy(const orc_pattern* p, int64_t* rowptr, int32_t* colidx);
void orc_pattern_free(orc_pattern* p);

/* ---- assembly: one grid walk over everything that was appended ------------------------------ */

/* Element forms, coupling forms (filter InnerIntersectionsOnce [+ periodic once]) and boundary forms
 * (filter: all boundary intersections) are applied per element in append order, like
 * XT::Grid::Walker::walk.  values must hold nnz doubles and is ADDED to (not cleared).
 * rhs_forms are LocalElementIntegralFunctional(LocalProductIntegrand(weight).with_ansatz(f)):
 * terms[0].diffusion = weight (default 1), terms[0].weight = f.  rhs (size ndof) is added to.
 * num_threads > 1 mirrors walk(use_tbb=true): contiguous element ranges per thread, shared
 * containers with row-striped locks. */
int orc_assemble(const orc_grid* g, int kind, int order, const int64_t* rowptr, const int32_t* colidx, double* values,
                 int n_element_forms, const orc_form* element_forms, int n_coupling_forms,
                 const orc_form* coupling_forms, int n_boundary_forms, const orc_form* boundary_forms,
                 int n_rhs_forms, const orc_form* rhs_forms, double* rhs, int num_threads);

/* the local element matrix of one element form (for unit tests) */
void orc_local_element_matrix(const orc_grid* g, int kind, int order, const orc_form* form, int64_t element,
                              double* out /* n*n */);

/* ---- finite volumes ------------------------------------------------------------------------ */

/* AdvectionFvOperator::apply: range = 0; for each inner (and periodic) intersection once: coupling op */
int orc_fv_apply(const orc_grid* g, const orc_flux* flux, const double* u, double* out, int num_threads);
/* u <- u - L(u) * dt (examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:152-157), n_steps times */
int orc_fv_euler(const orc_grid* g, const orc_flux* flux, double* u, double dt, int64_t n_steps, int num_threads);
/* the same with boundary treatments appended (in this order) */
int orc_fv_apply_bnd(const orc_grid* g, const orc_flux* flux, int n_bnd, const orc_fv_boundary* bnd, const double* u,
                     double* out, int num_threads);
/* ExplicitRungeKuttaTimeStepper::step (tools/timestepper/explicit-rungekutta.hh:237-270) for u_t = r L(u):
 * Butcher array A (row-major s x s), b, c; *t is advanced by min(dt, max_dt) */
int orc_rk_step(const orc_grid* g, const orc_flux* flux, int n_bnd, const orc_fv_boundary* bnd, int num_stages,
                const double* A, const double* b, const double* c, double r, double* u, double* t, double dt,
                double max_dt, int num_threads);
/* TimeStepperInterface::solve (tools/timestepper/interface.hh:191-263), nothing saved or written */
int orc_rk_solve(const orc_grid* g, const orc_flux* flux, int n_bnd, const orc_fv_boundary* bnd, int num_stages,
                 const double* A, const double* b, const double* c, double r, double* u, double t0, double t_end,
                 double initial_dt, int64_t* n_steps, double* t_final, int num_threads);
/* estimate_dt_for_hyperbolic_system (tools/hyperbolic.hh:38-86), m = 1, FV state; boundary_data_range: {min, max} or NULL */
double orc_fv_estimate_dt(const orc_grid* g, const orc_flux* flux, const double* u, const double* boundary_data_range);
/* default_interpolation into the FV space: cell average by a Gauss rule of the declared order
 * (spaces/basis/finite-volume.hh:244-252) */
void orc_fv_interpolate(const orc_grid* g, const orc_function* f, double* u);

/* systems of conservation laws: the Euler equations (flux kind ORC_FLUX_EULER, p[0] = gamma; Lax-Friedrichs: lambda =
 * p[1]), m = d + 2 components per cell, DoF m * element + i; tools/euler.hh, local/operators/advection-fv.hh:127-153,
 * examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:141-159, tools/hyperbolic.hh:38-86 */
int orc_fvsys_apply(const orc_grid* g, const orc_flux* flux, const double* u, double* out);
int orc_fvsys_apply_walls(const orc_grid* g, const orc_flux* flux, uint32_t wall_mask, uint32_t mirror_mask, const double* u,
                          double* out);
int orc_fvsys_euler(const orc_grid* g, const orc_flux* flux, double* u, double dt, int64_t n_steps);
double orc_fvsys_estimate_dt(const orc_grid* g, const orc_flux* flux, const double* u);
void orc_euler_flux(int d, double gamma, const double* w, double* f);
void orc_euler_jacobian(int d, double gamma, const double* w, double* J);
void orc_euler_eigen(int d, double gamma, const double* w, const double* n, double* ev, double* T, double* Ti);

/* evaluate a function descriptor at a global point (scalar view) */
double orc_function_eval(const orc_function* f, int dim, const double* x, int64_t element);

/* ---- SURVEY.md section 8f "next" rows ------------------------------------------------------------ */

/* DirichletConstraints (tools/dirichlet-constraints.hh:85-110): number of Dirichlet DoFs; if out != NULL the DoFs in
 * ascending order (std::set).  boundary_mask bit (2k+s): domain face with normal -e_k / +e_k is Dirichlet. */
int64_t orc_dirichlet_dofs(const orc_grid* g, int kind, int order, uint32_t boundary_mask, int64_t* out);
/* DirichletConstraints::apply (dirichlet-constraints.hh:122-184) on a CSR matrix and/or a vector (either may be NULL) */
int orc_dirichlet_apply(int64_t rows, const int64_t* rowptr, const int32_t* colidx, double* values, double* vector,
                        int64_t n_dofs, const int64_t* dofs, int only_clear, int ensure_symmetry);
/* ConstMatrixOperator::apply = matrix.mv (operators/matrix-based.hh:121-129) */
void orc_csr_mv(int64_t rows, const int64_t* rowptr, const int32_t* colidx, const double* values, const double* x,
                double* y);
/* BilinearForm::apply2 with source = range = u_h - f (operators/bilinear-form.hh:340-452); dofs or f may be NULL */
double orc_bilinear_form_apply2(const orc_grid* g, int kind, int order, const double* dofs, const orc_function* f,
                                const orc_form* form);
/* default_interpolation into a Lagrange space (interpolations/default.hh:40-83) */
void orc_lagrange_interpolate(const orc_grid* g, int kind, int order, const orc_function* f, double* dofs);

/* pointwise integrand evaluation on caller-supplied bases (the reference's evaluates_correctly_* tests,
 * dune/gdt/test/integrands/integrands_laplace.cc:71-91, integrands_product.cc:58-76) */
int orc_element_integrand_evaluate(const orc_integrand* integrand, int dim, int n_test, const double* test_values,
                                   const double* test_grads, int n_ansatz, const double* ansatz_values,
                                   const double* ansatz_grads, const double* x, double* result);

const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
