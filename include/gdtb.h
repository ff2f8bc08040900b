/*
 * include/gdtb.h -- C ABI of libgdtb.so: the B200 (sm_100a) implementation of dune-gdt's assembly and
 * finite-volume operator-apply hot path.
 *
 * dune-gdt is a header-only C++ template library: it has no FFI of its own.  The boundary a maintainer
 * binds is therefore the set of calls the C++ facade (dune-gdt_b200/include/dune/gdt/...) makes from
 * MatrixOperator::assemble / VectorBasedFunctional::assemble / AdvectionFvOperator::apply.  Every entry
 * point below names the reference interface it replaces (paths relative to the dune-gdt checkout).
 *
 * Conventions
 *   - plain C: opaque handles, POD descriptors, pointers and sizes only; no C++/torch types.
 *   - every function returns an int status: 0 = ok, otherwise one of the GDTB_ERR_* codes that mirror the
 *     reference's exception classes (dune/gdt/exceptions.hh:24-75); gdtb_last_error() returns the
 *     thread-local message.  Nothing throws across the boundary.
 *   - ownership: descriptors (forms, integrands, functions, fluxes) are CLONED at append/create time, like
 *     copy()/copy_as_*_integrand() in the reference (operators/matrix-based.hh:346-408), including any
 *     host array a gdtb_function points to; handles are owned by the caller until the matching *_destroy.
 *     Device buffers are owned by the library unless a *_set_*_device call lends one.
 *   - "host" entry points take host pointers and copy; "device" entry points take/return device pointers
 *     valid on the context's device.
 *   - all work is issued on the context's CUDA stream; calls block until complete unless documented
 *     otherwise (*_async).  A handle must not be used from two threads at once (the reference's
 *     operators are not re-entrant either).
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails with
 *     GDTB_ERR_CUDA.
 *
 * Canonical matrix format (XT::LA::IstlRowMajorSparseMatrix / EigenRowMajorSparseMatrix compatible CSR):
 * rowptr int64[rows+1], colidx int32[nnz] (ascending, unique per row), values double[nnz].
 */
#ifndef GDTB_H
#define GDTB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDTB_VERSION 1

/* ---- status codes --------------------------------------------------------------------------- */
enum
{
  GDTB_OK = 0,
  GDTB_ERR_INVALID_ARGUMENT = 1, /* XT::Common::Exceptions::wrong_input_given                         */
  GDTB_ERR_SHAPES_DO_NOT_MATCH = 2, /* XT::Common::Exceptions::shapes_do_not_match (matrix-based.hh:73-80) */
  GDTB_ERR_INTEGRAND = 3,        /* Exceptions::integrand_error (ipdg.hh:93-94)                       */
  GDTB_ERR_FINITE_ELEMENT = 4,   /* Exceptions::finite_element_error (lagrange.hh:107-136)            */
  GDTB_ERR_SPACE = 5,            /* Exceptions::space_error / mapper_error                            */
  GDTB_ERR_OPERATOR = 6,         /* Exceptions::operator_error                                        */
  GDTB_ERR_NOT_IMPLEMENTED = 7,  /* Dune::NotImplemented                                              */
  GDTB_ERR_CUDA = 8,             /* CUDA runtime failure / no device                                  */
  GDTB_ERR_OUT_OF_MEMORY = 9
};

/* ---- descriptors ---------------------------------------------------------------------------- */

/* XT::Grid::make_cube_grid<YaspGrid<d, EquidistantOffsetCoordinates<double,d>>>(lower, upper, n)
 * (examples/stationary-heat-equation.cc:87) [+ make_periodic_grid_view, examples/mpi_2019_02...cc:261].
 * slab_begin/slab_end select the element layers [begin, end) along the LAST direction that this
 * process owns (multi-GPU element-block partition); 0/0 means the whole grid. */
typedef struct gdtb_grid_desc
{
  int32_t dim;      /* 1, 2 or 3 */
  int32_t periodic; /* bit k: periodic in direction k */
  double lower[3];
  double upper[3];
  int64_t n[3];
} gdtb_grid_desc;

enum
{
  GDTB_SPACE_CG = 0, /* make_continuous_lagrange_space    (spaces/h1/continuous-lagrange.hh:189-203)     */
  GDTB_SPACE_DG = 1, /* make_discontinuous_lagrange_space (spaces/l2/discontinuous-lagrange.hh:191-205)  */
  GDTB_SPACE_FV = 2  /* make_finite_volume_space          (spaces/l2/finite-volume.hh:208-230)           */
};

/* Dune::GDT::Stencil (type_traits.hh:55-60) */
enum
{
  GDTB_STENCIL_ELEMENT = 0,
  GDTB_STENCIL_INTERSECTION = 1,
  GDTB_STENCIL_ELEMENT_AND_INTERSECTION = 2
};

/* XT::Functions::GridFunction<E, r, rC> stand-ins.  User lambdas cannot cross a C ABI as code, so a
 * grid function is a constant, a per-element array, one of a few analytic built-ins evaluated at the
 * global coordinate, an array the binder sampled at the quadrature points of the form (this is how an
 * XT::Functions::GenericFunction lambda crosses the boundary: the C++ facade samples it, see
 * dune-gdt_b200/include/dune/gdt/b200.hh), or a discrete function given by its DoF vector.  `order` is what
 * GridFunction::order() would return: it enters the quadrature order exactly like in the reference
 * (laplace.hh:74-79, product.hh:89-100, conversion.hh:92). */
enum
{
  GDTB_FN_CONST_SCALAR = 0, /* c[0]; used as a d x d function it means c[0] * I (laplace.hh:41) */
  GDTB_FN_CONST_TENSOR = 1, /* c[0 .. d*d) row-major                                             */
  GDTB_FN_ELEM_SCALAR = 2,  /* data[e], e = element index of the grid view                        */
  GDTB_FN_ELEM_TENSOR = 3,  /* data[e*d*d + r*d + c]                                              */
  GDTB_FN_BUILTIN = 4,      /* see GDTB_BUILTIN_*                                                 */
  /* Arbitrary coefficient data (an XT::Functions::GenericFunction lambda, measured data, ...) sampled by the caller
   * at the quadrature points of the form it is appended to: data[e * qp_per_element + q] (scalar) or
   * data[(e * qp_per_element + q) * d * d + r * d + c] (tensor), q = q_0 + m (q_1 + m q_2) the tensor Gauss rule with
   * m = gdtb_gauss_points(gdtb_form_quadrature_order(...)) points per direction, x_q = lower_e + xhat_q * (upper_e -
   * lower_e).  Element forms and functionals only (volume rules). */
  GDTB_FN_QP_SCALAR = 5,
  GDTB_FN_QP_TENSOR = 6,
  /* A discrete function u_h = sum_i data[global_index(e, i)] phi_i of the Lagrange / FV space (space_kind,
   * space_order) on the same grid (XT::Functions::GridFunction wrapping a DiscreteFunction,
   * discretefunction/default.hh): evaluated on the device at every quadrature point; as a d x d function it means
   * u_h * I.  `order` should be space_order (DiscreteFunction::order()). */
  GDTB_FN_DOF_VECTOR = 7,
  /* A function together with its jacobian, sampled like GDTB_FN_QP_SCALAR: data[(e * qp_per_element + q) * (1 + d)] is
   * the value, the next d entries the gradient.  Accepted as the analytic function `f` of gdtb_bilinear_form_apply2
   * (the exact solution of examples/stationary-heat-equation.cc:71-85 with its jacobian lambda); the rule is the one
   * gdtb_bilinear_form_quadrature_order reports. */
  GDTB_FN_QP_VALUE_GRAD = 8
};

enum
{
  GDTB_BUILTIN_COS_PRODUCT = 1, /* p0 * prod_i cos(p1 * x_i)  (heat-equation source, ESV2007 force) */
  GDTB_BUILTIN_AFFINE = 2,      /* p0 + sum_i p[1+i] * x_i                                           */
  GDTB_BUILTIN_GAUSSIAN = 3,    /* exp(-(x_0 - p0)^2 / (2 p1^2))  (examples/mpi...cc:321-326)        */
  GDTB_BUILTIN_INDICATOR = 4,   /* p0 <= x_0 <= p1 ? 1 : 0        (examples/mpi...cc:275-283)        */
  GDTB_BUILTIN_QUADRATIC = 5    /* p0 + p1 * sum_i x_i^2                                             */
};

typedef struct gdtb_function
{
  int32_t kind;
  int32_t order;
  int32_t builtin;
  int32_t data_on_device; /* 0: `data` is a host array (cloned at append); 1: device array (borrowed) */
  double c[9];
  double p[8];
  const double* data;
  int32_t qp_per_element; /* GDTB_FN_QP_*: quadrature points per element `data` was sampled for (checked at append) */
  int32_t space_kind;     /* GDTB_FN_DOF_VECTOR: GDTB_SPACE_* and polynomial order of the discrete function's space */
  int32_t space_order;
  int32_t reserved;
} gdtb_function;

enum
{
  GDTB_INT_LAPLACE = 0,                 /* LocalLaplaceIntegrand(diffusion)                   laplace.hh:40-48     */
  GDTB_INT_PRODUCT = 1,                 /* LocalElementProductIntegrand(weight = diffusion)   product.hh:56-65     */
  GDTB_INT_IPDG_INNER_COUPLING = 2,     /* LocalLaplaceIPDGIntegrands::InnerCoupling          laplace-ipdg.hh:47-61 */
  GDTB_INT_IPDG_INNER_PENALTY = 3,      /* LocalIPDGIntegrands::InnerPenalty                  ipdg.hh:59-72        */
  GDTB_INT_IPDG_DIRICHLET_COUPLING = 4, /* LocalLaplaceIPDGIntegrands::DirichletCoupling      laplace-ipdg.hh:237-252 */
  GDTB_INT_IPDG_BOUNDARY_PENALTY = 5    /* LocalIPDGIntegrands::BoundaryPenalty               ipdg.hh:201-213      */
};

enum
{
  GDTB_HI_DIAMETER = 0, /* internal::default_intersection_diameter (ipdg.hh:27-38)          */
  GDTB_HI_VOLUME = 1    /* intersection.geometry().volume() (test/.../ESV2007.hh:108-112)   */
};

typedef struct gdtb_integrand
{
  int32_t kind;
  int32_t hI_kind;
  double prefactor;        /* symmetry_prefactor (couplings) or penalty (penalties) */
  gdtb_function diffusion; /* kappa; for GDTB_INT_PRODUCT: the weight */
  gdtb_function weight;    /* omega of the IPDG integrands; for a functional: the function f of with_ansatz(f) */
} gdtb_integrand;

#define GDTB_MAX_TERMS 4

/* Local*IntegralBilinearForm / LocalElementIntegralFunctional (integrand, over_integrate)
 * (local/bilinear-forms/integrals.hh:52-63,169-180,305-316; local/functionals/integrals.hh:41-52);
 * n_terms > 1 is integrand_a + integrand_b (local/integrands/combined.hh). */
typedef struct gdtb_form
{
  int32_t n_terms;
  int32_t over_integrate;
  double scaling; /* MatrixOperator::scaling at append time (matrix-based.hh:342,365) */
  gdtb_integrand terms[GDTB_MAX_TERMS];
} gdtb_form;

/* intersection filters [XT::Grid::ApplyOn], used at matrix-based.hh:371-408 */
enum
{
  GDTB_FILTER_INNER_ONCE = 0,              /* ApplyOn::InnerIntersectionsOnce                        */
  GDTB_FILTER_INNER_AND_PERIODIC_ONCE = 1, /* ... || PeriodicBoundaryIntersectionsOnce               */
  GDTB_FILTER_ALL_BOUNDARY = 2             /* CustomBoundaryIntersections(AllDirichletBoundaryInfo, DirichletBoundary) */
};

enum
{
  GDTB_FLUX_LINEAR = 0,  /* f(u) = a u, a = p[0..d)  (test/linear-transport/base.hh:50-57) */
  GDTB_FLUX_BURGERS = 1, /* f(u) = u^2/2 (1,...,1)   (test/burgers/base.hh:38-44)          */
  /* Euler equations of gas dynamics, m = d + 2 conservative variables w = (rho, rho v, E), d = 1, 2: EulerTools<d>::flux
   * / flux_jacobian / eigenvalues_ / eigenvectors_ / eigenvectors_inv_flux_jacobian (tools/euler.hh:212-236, 262-316,
   * 325-462); p[0] = gamma.  Needs a finite volume space with m components (gdtb_fv_space_create). */
  GDTB_FLUX_EULER = 2
};

enum
{
  GDTB_NUMFLUX_UPWIND = 0,         /* NumericalUpwindFlux<I,d,1>      upwind.hh:44-73         */
  GDTB_NUMFLUX_LAX_FRIEDRICHS = 1, /* NumericalLaxFriedrichsFlux      lax-friedrichs.hh:60-88; systems (m > 1): the
                                      caller provides lambda = p[1] (lax-friedrichs.hh:40-41)  */
  /* NumericalVijayasundaramFlux<I,d,m> (vijayasundaram.hh:111-133) with the flux's own eigendecomposition (the lambda
   * the reference's 2d_euler driver passes, examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:402-409) */
  GDTB_NUMFLUX_VIJAYASUNDARAM = 2
};

typedef struct gdtb_flux
{
  int32_t kind;
  int32_t numflux;
  double p[4];
} gdtb_flux;

/* Boundary treatments of the FV operator: AdvectionFvOperator::append(lambda, ..., filter)
 * (operators/advection-fv.hh:96-123) -> LocalAdvectionFvBoundaryTreatmentByCustomNumericalFluxOperator /
 * ...ByCustomExtrapolationOperator (local/operators/advection-fv.hh:188-457).  The reference takes arbitrary C++
 * lambdas; across a C ABI they are the two affine families below.  side_mask selects the boundary intersections
 * (the reference's intersection filter): bit (2k+s) = domain face with outer normal -e_k (s = 0) / +e_k (s = 1). */
enum
{
  GDTB_FVBND_EXTRAPOLATION = 0, /* v = a u + b, g = numerical_flux(u, v, n): a=1,b=0 absorbing; a=0 Dirichlet value b */
  GDTB_FVBND_NUMERICAL_FLUX = 1, /* g = a (f(u) . n) + b: a=1,b=0 outflow of the physical flux; a=0 prescribed flux b */
  /* the two impermeable-wall treatments of the Euler equations the reference's tests append to a system operator
   * (test/inviscid-compressible-flow/base.hh:187-241; a, b unused):                                                   */
  GDTB_FVBND_EULER_IMPERMEABLE_WALL = 2, /* numerical boundary flux g = EulerTools::flux_at_impermeable_walls(u, n)     */
                                         /* = (0, p n, 0) (tools/euler.hh:238-251, [DF2015, (8.58)])                     */
  GDTB_FVBND_EULER_INVISCID_MIRROR = 3   /* extrapolation v = conservative(rho, v - 2 (v . n) n, p), g = numerical_flux  */
                                         /* (u, v, n) ([DF2015, (8.66-8.67)])                                            */
};
typedef struct gdtb_fv_boundary
{
  int32_t kind;
  uint32_t side_mask;
  double a, b;
} gdtb_fv_boundary;

/* TimeStepperMethods (tools/timestepper/explicit-rungekutta.hh:63-141) */
enum
{
  GDTB_RK_EULER = 0,    /* explicit_euler                              */
  GDTB_RK_SSP2 = 1,     /* explicit_rungekutta_second_order_ssp        */
  GDTB_RK_SSP3 = 2,     /* explicit_rungekutta_third_order_ssp         */
  GDTB_RK_CLASSIC4 = 3, /* explicit_rungekutta_classic_fourth_order    */
  GDTB_RK_OTHER = 4     /* explicit_rungekutta_other: user Butcher array */
};
#define GDTB_RK_MAX_STAGES 8

enum
{
  GDTB_ASSEMBLE_OVERWRITE = 0, /* matrix/vector are known to be zero (fresh containers): values = assembled */
  GDTB_ASSEMBLE_ACCUMULATE = 1 /* values += assembled (add_to_entry into existing content)                  */
};

typedef struct gdtb_ctx gdtb_ctx;
typedef struct gdtb_grid gdtb_grid;
typedef struct gdtb_space gdtb_space;
typedef struct gdtb_pattern gdtb_pattern;
typedef struct gdtb_matop gdtb_matop;
typedef struct gdtb_vecfun gdtb_vecfun;
typedef struct gdtb_fvop gdtb_fvop;
typedef struct gdtb_rk gdtb_rk;

/* ---- context -------------------------------------------------------------------------------- */
const char* gdtb_last_error(void);
int gdtb_version(void);
/* one context per process and GPU; creates its own non-blocking stream */
int gdtb_ctx_create(int device, gdtb_ctx** ctx);
int gdtb_ctx_destroy(gdtb_ctx* ctx);
/* run on a caller-provided cudaStream_t (e.g. torch's current stream); NULL restores the own stream */
int gdtb_ctx_set_stream(gdtb_ctx* ctx, void* cuda_stream);
int gdtb_ctx_synchronize(gdtb_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t gdtb_ctx_launch_count(const gdtb_ctx* ctx);
/* optional per-kernel timing: CUDA events on the launching stream directly around the launches of one kernel family
 * ("q1_gather", "fv_apply", "element_matrix", "element_vector", "coupling_matrix", "boundary_matrix").
 * gdtb_ctx_kernel_time synchronises, returns the summed device time and launch count since the last query and
 * resets them. */
int gdtb_ctx_enable_timing(gdtb_ctx* ctx, int enabled);
int gdtb_ctx_kernel_time(gdtb_ctx* ctx, const char* family, double* total_ms, int64_t* launches);
/* (mangled) symbol of the kernel instantiation the family launched last, as it shows up in an ncu launch list; "" if the
 * family has not launched yet (gather families only) */
const char* gdtb_ctx_kernel_name(const gdtb_ctx* ctx, const char* family);

/* ---- grid / spaces -------------------------------------------------------------------------- */
/* replaces XT::Grid::make_cube_grid + leaf_view (examples/stationary-heat-equation.cc:87-88) */
int gdtb_grid_create_cube(gdtb_ctx* ctx, const gdtb_grid_desc* desc, gdtb_grid** grid);
int gdtb_grid_destroy(gdtb_grid* grid);
int64_t gdtb_grid_num_elements(const gdtb_grid* grid);

/* replaces make_{continuous_lagrange,discontinuous_lagrange,finite_volume}_space */
int gdtb_space_create(gdtb_ctx* ctx, const gdtb_grid* grid, int kind, int order, gdtb_space** space);
/* make_finite_volume_space<m>(grid_view) (spaces/l2/finite-volume.hh:208-230): m DoFs per element, global index
 * m * element + i (spaces/mapper/finite-volume.hh:92-97); 1 <= range_dim <= 4.  An advection operator on such a space
 * (gdtb_fvop_create with GDTB_FLUX_EULER) supports gdtb_fvop_apply[_host], gdtb_fvop_euler[_host] and
 * gdtb_fv_estimate_dt[_host] on periodic or unpartitioned grids. */
int gdtb_fv_space_create(gdtb_ctx* ctx, const gdtb_grid* grid, int range_dim, gdtb_space** space);
int gdtb_space_destroy(gdtb_space* space);
/* MapperInterface::size / max_local_size / global_indices (spaces/mapper/interfaces.hh) */
int64_t gdtb_space_size(const gdtb_space* space);
int32_t gdtb_space_max_local_size(const gdtb_space* space);
int gdtb_space_global_indices(const gdtb_space* space, int64_t element, int64_t* out);

/* ---- quadrature (what a binder needs to pre-sample a user lambda into GDTB_FN_QP_*) ------------------------------ */
enum
{
  GDTB_ROLE_ELEMENT = 0,  /* LocalElementIntegralBilinearForm         (integrals.hh:115: integrand.order + over_integrate) */
  GDTB_ROLE_FUNCTIONAL = 1, /* LocalElementIntegralFunctional           (local/functionals/integrals.hh:85)              */
  GDTB_ROLE_COUPLING = 2, /* LocalCouplingIntersectionIntegralBilinearForm (integrals.hh:225-229)                        */
  GDTB_ROLE_BOUNDARY = 3  /* LocalIntersectionIntegralBilinearForm    (integrals.hh:354-355)                             */
};
/* polynomial order the reference integrates `form` with on `space` (laplace.hh:74-79, product.hh:89-100,
 * conversion.hh:92, combined.hh:293-299, laplace-ipdg.hh:95-105, ipdg.hh:100-111) */
int gdtb_form_quadrature_order(const gdtb_space* space, const gdtb_form* form, int role, int32_t* order);
/* Dune::QuadratureRules<double, 1>::rule(cube, order) [EXT dune-geometry]: m = order / 2 + 1 Gauss-Legendre points on
 * [0, 1] (ascending) and their weights; the d-dimensional rule is the tensor product, first coordinate fastest.
 * points01 / weights may be NULL (query m only). */
int gdtb_gauss_rule(int order, int32_t* m, double* points01, double* weights);

/* ---- sparsity pattern ----------------------------------------------------------------------- */
enum
{
  GDTB_PATTERN_AUTO = 0,        /* structured closed form when available, else sort-unique */
  GDTB_PATTERN_SORT_UNIQUE = 1, /* emit (row,col) keys per element/intersection, radix sort, unique, CSR */
  GDTB_PATTERN_STRUCTURED = 2   /* closed-form generators: CG Q1 / Q2 element stencil, DG element_and_intersection */
};
/* replaces make_sparsity_pattern(test, ansatz, view, stencil) (tools/sparsity-pattern.hh:163-178) */
int gdtb_pattern_create(gdtb_ctx* ctx, const gdtb_space* test, const gdtb_space* ansatz, int stencil, int method,
                        gdtb_pattern** pattern);
int gdtb_pattern_destroy(gdtb_pattern* pattern);
int64_t gdtb_pattern_rows(const gdtb_pattern* pattern);
int64_t gdtb_pattern_cols(const gdtb_pattern* pattern);
int64_t gdtb_pattern_nnz(const gdtb_pattern* pattern);
int gdtb_pattern_download(const gdtb_pattern* pattern, int64_t* rowptr, int32_t* colidx);
int gdtb_pattern_device(const gdtb_pattern* pattern, const int64_t** d_rowptr, const int32_t** d_colidx);

/* Host-side view of the closed-form CSR geometry (no device, no context needed): the row pointers the structured
 * pattern generators / gather kernels use for the CG Q1 and CG Q2 element stencils and the DG element_and_intersection
 * stencil on a non-periodic grid (rowptr: gdtb_space_size + 1 entries), and the row ranges a slab of element layers
 * [layer_begin, layer_end) owns (as gdtb_matop_local_row_ranges reports them after gdtb_matop_set_slab).  For sizing
 * containers before any device work (make_*_sparsity_pattern(...).size() in the reference, tools/sparsity-pattern.hh). */
int gdtb_host_closed_form_rowptr(const gdtb_grid_desc* grid, int kind, int order, int64_t* rowptr);
int gdtb_host_slab_row_ranges(const gdtb_grid_desc* grid, int kind, int order, int64_t layer_begin, int64_t layer_end,
                              int32_t max_ranges, int64_t* row_begin, int64_t* row_end, int64_t* value_offset,
                              int64_t* value_count, int32_t* n_ranges);

/* ---- MatrixOperator (operators/matrix-based.hh:245-508) -------------------------------------- */
/* make_matrix_operator<M>(view, source_space, range_space, pattern) (matrix-based.hh:514-598):
 * rows = range/test space, cols = source/ansatz space.  `pattern` may be NULL for continuous Q1 and Q2 spaces on
 * non-periodic grids: their element stencils have closed forms, the values then follow the layout
 * gdtb_pattern_create would produce (the pattern is materialised only if a later call needs colidx). */
int gdtb_matop_create(gdtb_ctx* ctx, const gdtb_space* test, const gdtb_space* ansatz, const gdtb_pattern* pattern,
                      gdtb_matop** op);
int gdtb_matop_destroy(gdtb_matop* op);
/* MatrixOperator::append(LocalElementBilinearFormInterface) (matrix-based.hh:346-369), filter AllElements */
int gdtb_matop_append_element(gdtb_matop* op, const gdtb_form* form);
/* MatrixOperator::append(LocalCouplingIntersectionBilinearFormInterface, param, filter) (matrix-based.hh:371-393) */
int gdtb_matop_append_coupling(gdtb_matop* op, const gdtb_form* form, int filter);
/* MatrixOperator::append(LocalIntersectionBilinearFormInterface, param, filter) (matrix-based.hh:395-408) */
int gdtb_matop_append_boundary(gdtb_matop* op, const gdtb_form* form, int filter);
/* drops all appended forms (dune-xt's walker clears its functors after each walk) */
int gdtb_matop_clear_forms(gdtb_matop* op);
int gdtb_matop_num_forms(const gdtb_matop* op);
/* name of the kernel family the next assemble will use: "q1_gather", "generic_coloured", ... (diagnostics) */
const char* gdtb_matop_plan(gdtb_matop* op);
/* Why the next assemble takes that family: for "generic_coloured" (the quadrature-faithful any-configuration kernels,
 * several times the compulsory memory traffic) the first property of the operator that rules out the row-gather
 * kernels, "" otherwise.  With GDTB_WARN_GENERIC=1 in the environment gdtb_assemble prints it once per operator. */
const char* gdtb_matop_plan_reason(gdtb_matop* op);
int gdtb_matop_set_zero(gdtb_matop* op);
int gdtb_matop_values_download(const gdtb_matop* op, double* values);
int gdtb_matop_values_upload(gdtb_matop* op, const double* values);
int gdtb_matop_values_device(const gdtb_matop* op, double** d_values);
/* lend a device buffer of nnz doubles (MatrixOperator's "borrow a caller matrix" ctor, matrix-based.hh:280-293) */
int gdtb_matop_set_values_device(gdtb_matop* op, double* d_values);

/* ---- VectorBasedFunctional (functionals/vector-based.hh:133-286) ------------------------------ */
int gdtb_vecfun_create(gdtb_ctx* ctx, const gdtb_space* space, gdtb_vecfun** fun);
int gdtb_vecfun_destroy(gdtb_vecfun* fun);
/* append(LocalElementIntegralFunctional(LocalProductIntegrand(w).with_ansatz(f))) (vector-based.hh:214-222):
 * form->terms[0].kind = GDTB_INT_PRODUCT, .diffusion = w, .weight = f */
int gdtb_vecfun_append_element(gdtb_vecfun* fun, const gdtb_form* form);
int gdtb_vecfun_clear_forms(gdtb_vecfun* fun);
int gdtb_vecfun_set_zero(gdtb_vecfun* fun);
int gdtb_vecfun_download(const gdtb_vecfun* fun, double* vector);
int gdtb_vecfun_device(const gdtb_vecfun* fun, double** d_vector);
int gdtb_vecfun_set_device(gdtb_vecfun* fun, double* d_vector);

/* ---- the grid walk --------------------------------------------------------------------------- */
/* MatrixOperator::assemble / VectorBasedFunctional::assemble / walker.walk() with both appended
 * (matrix-based.hh:496-500, vector-based.hh:276-279, examples/stationary-heat-equation.cc:102-106):
 * one pass over the grid for everything appended to `op` and `fun` (either may be NULL). */
int gdtb_assemble(gdtb_matop* op, gdtb_vecfun* fun, int mode);
/* same, but only enqueues the kernels on the context's stream; errors detected on the device (entries outside the
 * pattern) surface at the next gdtb_ctx_synchronize */
int gdtb_assemble_async(gdtb_matop* op, gdtb_vecfun* fun, int mode);
/* convenience for host-side callers: assemble (overwrite) and copy values / vector to host buffers (may be NULL) */
int gdtb_assemble_host(gdtb_matop* op, gdtb_vecfun* fun, double* values, double* vector);

/* ---- AdvectionFvOperator (operators/advection-fv.hh:44-141) ----------------------------------- */
/* make_advection_fv_operator<M>(view, numerical_flux, source_space, range_space) */
int gdtb_fvop_create(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_flux* flux, gdtb_fvop** L);
int gdtb_fvop_destroy(gdtb_fvop* L);
/* LocalizableOperator::apply(source, range) (operators/localizable-operator.hh:352-387): range = L(source).
 * Device pointers, each gdtb_space_size doubles. */
int gdtb_fvop_apply(gdtb_fvop* L, const double* d_source, double* d_range);
int gdtb_fvop_apply_host(gdtb_fvop* L, const double* source, double* range);
/* explicit Euler u <- u - L(u) dt (examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:152-157), fused
 * apply + axpy, n_steps times, ping-ponging between d_u and an internal buffer; result ends in d_u. */
int gdtb_fvop_euler(gdtb_fvop* L, double* d_u, double dt, int64_t n_steps);
int gdtb_fvop_euler_host(gdtb_fvop* L, double* u, double dt, int64_t n_steps);
/* Multi-GPU (slab partition along the last direction): the local vector carries one ghost layer below
 * and above the owned cells: layout [ghost_lo | owned | ghost_hi], each ghost layer = n[0]*..*n[dim-2]
 * cells.  The caller fills the ghost layers (NCCL send/recv of the neighbours' boundary layers,
 * tools/timestepper/explicit-rungekutta.hh:252-257) before calling apply. */
int gdtb_fvop_set_slab(gdtb_fvop* L, int64_t layer_begin, int64_t layer_end);
/* Enqueue-only building block of the multi-GPU time loop (no synchronisation): range = L(source) (euler == 0) or
 * range = source - dt * L(source) (euler != 0) for the layers [layer_begin, layer_end) of the operator's slab
 * (0, 0 = the whole slab).  Lets the caller overlap the interior layers with the ghost-layer exchange and finish the
 * first / last layer afterwards.  Vectors use the slab layout of gdtb_fvop_set_slab. */
int gdtb_fvop_step_async(gdtb_fvop* L, const double* d_source, double* d_range, int euler, double dt,
                         int64_t layer_begin, int64_t layer_end);
int64_t gdtb_fvop_ghost_layer_size(const gdtb_fvop* L);

/* Peer-memory ghost exchange for the multi-GPU time loop (one process per GPU on one NVLink / NVSwitch box): instead of
 * a host-launched send / recv per step (the DataHandle communicate() of tools/timestepper/explicit-rungekutta.hh:252-257)
 * the apply kernel itself stores its first / last owned layer into the neighbours' ghost layers and raises a step
 * counter in their memory; the blocks that read a ghost layer wait for the counter of the previous step.
 *   1. gdtb_fvop_set_slab, then gdtb_fvop_p2p_alloc: two library-owned slab vectors ([ghost | owned | ghost]) and
 *      their CUDA IPC handles (3 x GDTB_IPC_HANDLE_BYTES: u0, u1, counters) to be sent to the neighbour processes;
 *   2. gdtb_fvop_p2p_connect with the handles received from the lower / upper neighbour (NULL: no neighbour; *_is_self:
 *      the periodic neighbour is this process) and the neighbours' numbers of owned layers;
 *   3. fill u0 (owned part and ghost layers) once, then gdtb_fvop_p2p_step per step (enqueue only): step s reads
 *      u[s % 2] and writes u[(s + 1) % 2]; gdtb_fvop_p2p_current gives the vector holding the current solution;
 *   4. gdtb_fvop_p2p_check synchronises and reports a timed-out wait (GDTB_ERR_OPERATOR).
 * All processes must have finished step 2 before any of them starts step 3, and destroy their operators only after a
 * common barrier. */
#define GDTB_IPC_HANDLE_BYTES 64
int gdtb_fvop_p2p_alloc(gdtb_fvop* L, double** d_u0, double** d_u1, void* handles);
int gdtb_fvop_p2p_connect(gdtb_fvop* L, const void* lower_handles, int64_t lower_layers, int lower_is_self,
                          const void* upper_handles, int64_t upper_layers, int upper_is_self);
int gdtb_fvop_p2p_step(gdtb_fvop* L, int euler, double dt);
int gdtb_fvop_p2p_current(gdtb_fvop* L, double** d_u, int64_t* step);
int gdtb_fvop_p2p_check(gdtb_fvop* L);

/* AdvectionFvOperator::append(boundary treatment lambda, param_type, filter) (operators/advection-fv.hh:96-123):
 * applies on the non-periodic domain boundary faces selected by side_mask.  Treatments add up like appended local
 * operators do; two different extrapolations on the same side are GDTB_ERR_NOT_IMPLEMENTED. */
int gdtb_fvop_append_boundary(gdtb_fvop* L, const gdtb_fv_boundary* treatment);

/* estimate_dt_for_hyperbolic_system(grid_view, state, flux, boundary_data_range) (tools/hyperbolic.hh:38-86) for the
 * operator's grid and flux and a finite-volume state d_u (device): min / max of the state and max perimeter / volume
 * are reduced on the device.  boundary_data_range = {min, max} or NULL for the reference's defaults. */
int gdtb_fv_estimate_dt(gdtb_fvop* L, const double* d_u, const double* boundary_data_range, double* dt);
int gdtb_fv_estimate_dt_host(gdtb_fvop* L, const double* u, const double* boundary_data_range, double* dt);

/* ---- ExplicitRungeKuttaTimeStepper (tools/timestepper/explicit-rungekutta.hh:158-270) ---------------------------- */
/* ExplicitRungeKuttaTimeStepper<Op, DF, method>(op, initial_values, r, t_0[, A, b, c]): solves u_t = r L(u).
 * A (row-major, num_stages^2), b, c are read only for GDTB_RK_OTHER (num_stages <= GDTB_RK_MAX_STAGES); A must be
 * strictly lower triangular (wrong_input_given -> GDTB_ERR_INVALID_ARGUMENT, :216-222).  The current solution lives
 * in the caller's device vector passed to step / solve (the reference keeps a reference to initial_values). */
int gdtb_rk_create(gdtb_fvop* L, int method, int num_stages, const double* A, const double* b, const double* c, double r,
                   double t0, gdtb_rk** ts);
int gdtb_rk_destroy(gdtb_rk* ts);
double gdtb_rk_current_time(const gdtb_rk* ts);
/* step(dt, max_dt) (:237-270): one step of length min(dt, max_dt) on d_u; *returned_dt = dt (may be NULL) */
int gdtb_rk_step(gdtb_rk* ts, double* d_u, double dt, double max_dt, double* returned_dt);
int gdtb_rk_step_host(gdtb_rk* ts, double* u, double dt, double max_dt, double* returned_dt);
/* TimeStepperInterface::solve(t_end, initial_dt) (tools/timestepper/interface.hh:191-263) with nothing saved, written
 * or printed: steps of length initial_dt, the last one shortened to hit t_end.  The steps run back to back on the
 * device (CUDA graph replay of one step, no host synchronisation in between). */
int gdtb_rk_solve(gdtb_rk* ts, double* d_u, double t_end, double initial_dt, int64_t* n_steps, double* next_dt);
int gdtb_rk_solve_host(gdtb_rk* ts, double* u, double t_end, double initial_dt, int64_t* n_steps, double* next_dt);

/* Runge-Kutta on slabs (operator with gdtb_fvop_set_slab; >= 2 stages, 2D / 3D): the reference exchanges every stage
 * vector k_i with a DataHandle communicate() (explicit-rungekutta.hh:252-257).  Here the stepper owns the solution vector
 * and two alternating stage vectors in memory the neighbour processes open through CUDA IPC; after every stage vector
 * (and after the update of u_n) a small kernel stores the first / last owned layer into the neighbours' ghost layers and
 * raises their counters, and every operator apply waits inside the kernel for the counter of its source.
 *   gdtb_rk_create on the slab operator, gdtb_rk_p2p_handles (solution vector in slab layout + 4 IPC handles),
 *   gdtb_rk_p2p_connect (as gdtb_fvop_p2p_connect), fill the solution vector incl. ghost layers once, then
 *   gdtb_rk_step / gdtb_rk_solve with d_u = that vector; gdtb_rk_p2p_check reports a timed-out wait. */
int gdtb_rk_p2p_handles(gdtb_rk* ts, double** d_un, void* handles /* 4 * GDTB_IPC_HANDLE_BYTES */);
int gdtb_rk_p2p_connect(gdtb_rk* ts, const void* lower_handles, int64_t lower_layers, int lower_is_self,
                        const void* upper_handles, int64_t upper_layers, int upper_is_self);
int gdtb_rk_p2p_check(gdtb_rk* ts);

/* default_interpolation(order, f, fv_space) (interpolations/default.hh:76-83, spaces/basis/finite-volume.hh:244-252) */
int gdtb_fv_interpolate(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_function* f, double* d_u);
/* same with a host result buffer of gdtb_space_size doubles (per-element function data may live on the host) */
int gdtb_fv_interpolate_host(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_function* f, double* u);

/* ---- multi-GPU assembly ----------------------------------------------------------------------- */
/* Element-block partition: this process owns the element layers [begin, end) along the last direction and, owner-
 * computes-rows, the DoF rows of the vertex layers [begin, end) (the last slab also owns the top layer).  Owned
 * rows are complete: the one ghost element layer below the slab is recomputed locally (what the reference does
 * with YaspGrid's overlap), so assembly needs no communication; per-element coefficient arrays stay indexed by
 * the global element index.  After this call the value / vector buffers hold the owned rows only
 * (gdtb_matop_local_rows gives the global row range and the global CSR offset of the first owned value). */
int gdtb_matop_set_slab(gdtb_matop* op, int64_t layer_begin, int64_t layer_end);
int gdtb_vecfun_set_slab(gdtb_vecfun* fun, int64_t layer_begin, int64_t layer_end);
/* Interface-row halo partition (the other legal scheme): this process walks only its OWN element
 * layers [begin, end) and holds the rows of the vertex layers [begin, end]; the top layer is the interface owned by
 * the slab above and carries partial sums that must travel there (ncclSend / ncclRecv of one layer of rows:
 * gdtb_*_halo_layout gives the position of the layer to send up and of the layer that receives from below, -1 where
 * there is no neighbour), followed by gdtb_vector_add of the received layer.  CG Q1 only. */
int gdtb_matop_set_slab_halo(gdtb_matop* op, int64_t layer_begin, int64_t layer_end);
int gdtb_vecfun_set_slab_halo(gdtb_vecfun* fun, int64_t layer_begin, int64_t layer_end);
int gdtb_matop_halo_layout(const gdtb_matop* op, int64_t* recv_offset, int64_t* send_offset, int64_t* count);
int gdtb_vecfun_halo_layout(const gdtb_vecfun* fun, int64_t* recv_offset, int64_t* send_offset, int64_t* count);
/* The same partition with the halo handed over INSIDE the gather kernel (one process per GPU on one NVLink / NVSwitch
 * box) instead of a host-launched send / recv + add: the work items of the top (interface) layer are computed first and
 * stored straight into the receive buffer of the rank above (peer stores into memory opened through CUDA IPC), each
 * followed by a counter increment in that rank's memory; the items of the bottom layer are computed last, wait for the
 * counter of the rank below and add what arrived before their rows leave the SM.  Receive buffers alternate between
 * two step parities and the consumer acknowledges every step, so a rank may run ahead by at most one assembly.
 *   1. gdtb_matop_set_slab_halo (+ gdtb_vecfun_set_slab_halo), then gdtb_halo_p2p_alloc: the rank's receive buffers and
 *      counters + 2 IPC handles to be sent to both neighbours;
 *   2. gdtb_halo_p2p_connect with the handles of the rank below / above (NULL: none) and the number of element layers of
 *      the rank below; all ranks must have connected before any of them assembles;
 *   3. gdtb_assemble[_async](op, fun, OVERWRITE) as usual -- no exchange call follows; gdtb_halo_p2p_check
 *      synchronises and reports a timed-out wait (GDTB_ERR_OPERATOR). */
int gdtb_halo_p2p_alloc(gdtb_matop* op, void* handles /* 2 * GDTB_IPC_HANDLE_BYTES */);
int gdtb_halo_p2p_connect(gdtb_matop* op, const void* lower_handles, int64_t lower_layers, const void* upper_handles);
int gdtb_halo_p2p_check(gdtb_matop* op);
/* d_y[0..n) += d_x[0..n) on the context's stream (enqueue only) */
int gdtb_vector_add(gdtb_ctx* ctx, double* d_y, const double* d_x, int64_t n);
/* Plain vector transfers between host and device memory on the context's stream, synchronised on return: for hosts
 * without a CUDA runtime of their own (the C++ facade fills the slab vectors of gdtb_rk_p2p_handles /
 * gdtb_fvop_p2p_alloc with these, where the Python mirror uses torch tensors). */
int gdtb_vector_upload(gdtb_ctx* ctx, double* d_dst, const double* src, int64_t n);
int gdtb_vector_download(gdtb_ctx* ctx, double* dst, const double* d_src, int64_t n);
int64_t gdtb_matop_local_nnz(const gdtb_matop* op);
int gdtb_matop_local_rows(const gdtb_matop* op, int64_t* row_begin, int64_t* row_end, int64_t* value_offset);
/* CG Q2: the MCMG-based ContinuousMapper (spaces/mapper/continuous.hh:117-150) numbers DoFs [cells | faces | edges |
 * vertices], each group lexicographically, so a slab owns one contiguous global row range PER GROUP (2^dim ranges);
 * the local value buffer holds the ranges back to back, value_offset[r] is the global CSR position of range r's first
 * value, value_count[r] its number of values.  CG Q1 operators report their single range.  Call with max_ranges = 0 and NULL arrays to query *n_ranges. */
int gdtb_matop_local_row_ranges(const gdtb_matop* op, int32_t max_ranges, int64_t* row_begin, int64_t* row_end,
                                int64_t* value_offset, int64_t* value_count, int32_t* n_ranges);


/* ==== callers on either side of the hot path (n1, n2, n4) ====================== */

/* ---- DirichletConstraints (tools/dirichlet-constraints.hh:44-230) ------------------------------------ */
typedef struct gdtb_dirichlet gdtb_dirichlet;
#define GDTB_BOUNDARY_ALL 0x3f
/* make_dirichlet_constraints(space, boundary_info) + the element walk that collects dirichlet_DoFs()
 * (dirichlet-constraints.hh:85-110): the local DoFs whose local key sits on a sub-entity of a boundary
 * intersection of Dirichlet type.  boundary_mask: bit (2 k + s) set <=> the domain face with outer normal
 * -e_k (s = 0) / +e_k (s = 1) is a Dirichlet boundary; GDTB_BOUNDARY_ALL = XT::Grid::AllDirichletBoundaryInfo
 * (examples/stationary-heat-equation.cc:90).  Periodic directions have no boundary intersections. */
int gdtb_dirichlet_create(gdtb_ctx* ctx, const gdtb_space* space, uint32_t boundary_mask, gdtb_dirichlet** dc);
int gdtb_dirichlet_destroy(gdtb_dirichlet* dc);
/* dirichlet_DoFs().size() and the DoFs in ascending order (std::set iteration order) */
int64_t gdtb_dirichlet_size(const gdtb_dirichlet* dc);
int gdtb_dirichlet_dofs_download(const gdtb_dirichlet* dc, int64_t* dofs);
/* device views: the ascending DoF list and one flag byte per DoF of the space */
int gdtb_dirichlet_device(const gdtb_dirichlet* dc, const int64_t** d_dofs, const uint8_t** d_flags);
/* DirichletConstraints::apply(matrix, vector, only_clear, ensure_symmetry) (dirichlet-constraints.hh:122-184):
 * unit_row (+ unit_col) / clear_row (+ clear_col) for every Dirichlet DoF, vector[DoF] = 0; op or fun may be NULL
 * (the one-argument overloads).  One pass over the CSR rows, no host round trip. */
int gdtb_dirichlet_apply(gdtb_dirichlet* dc, gdtb_matop* op, gdtb_vecfun* fun, int only_clear, int ensure_symmetry);
/* the same on caller-owned device arrays: a CSR matrix given by `pattern` + d_values and/or a vector */
int gdtb_dirichlet_apply_device(gdtb_dirichlet* dc, const gdtb_pattern* pattern, double* d_values, double* d_vector,
                                int only_clear, int ensure_symmetry);

/* ---- ConstMatrixOperator::apply / apply_inverse (operators/matrix-based.hh:121-159) ------------------ */
/* range = matrix * source (matrix_.mv, matrix-based.hh:121-129); device vectors of cols / rows doubles */
int gdtb_matop_apply(gdtb_matop* op, const double* d_source, double* d_range);
int gdtb_matop_apply_host(gdtb_matop* op, const double* source, double* range);
/* the same for a caller-owned CSR matrix */
int gdtb_csr_apply_device(gdtb_ctx* ctx, const gdtb_pattern* pattern, const double* d_values, const double* d_source,
                          double* d_range);

enum
{
  GDTB_SOLVER_CG = 0,      /* conjugate gradients (symmetric positive definite systems)  */
  GDTB_SOLVER_BICGSTAB = 1 /* BiCGStab (unsymmetric systems, e.g. NIPDG)                 */
};
enum
{
  GDTB_PRECOND_NONE = 0,
  GDTB_PRECOND_JACOBI = 1 /* diagonal scaling */
};
/* XT::LA::make_solver(matrix).apply(rhs, solution, opts) [EXT dune-xt; "type", "precision", "max_iter" keys] */
typedef struct gdtb_solver_opts
{
  int32_t type;
  int32_t preconditioner;
  int32_t max_iter;    /* <= 0: 10 * rows */
  int32_t check_every; /* iterations per CUDA-graph replay between two host-side convergence checks; <= 0: 25 */
  double precision;    /* stop when |r| <= precision * |r_0| (defect reduction); <= 0: 1e-10 */
} gdtb_solver_opts;
typedef struct gdtb_solver_info
{
  int32_t iterations;
  int32_t converged;
  double initial_residual; /* |b - A x_0|_2 */
  double residual;         /* |b - A x|_2 of the recursively updated residual at exit */
} gdtb_solver_info;
/* apply_inverse(range, source, opts) (matrix-based.hh:148-159): solves matrix * d_x = d_rhs on the device, d_x
 * holds the initial guess on entry.  Fails with GDTB_ERR_OPERATOR when the solver does not converge
 * (XT::LA::Exceptions::linear_solver_failed -> Exceptions::operator_error); opts / info may be NULL. */
int gdtb_matop_apply_inverse(gdtb_matop* op, const double* d_rhs, double* d_x, const gdtb_solver_opts* opts,
                             gdtb_solver_info* info);
int gdtb_matop_apply_inverse_host(gdtb_matop* op, const double* rhs, double* x, const gdtb_solver_opts* opts,
                                  gdtb_solver_info* info);
int gdtb_csr_apply_inverse_device(gdtb_ctx* ctx, const gdtb_pattern* pattern, const double* d_values,
                                  const double* d_rhs, double* d_x, const gdtb_solver_opts* opts,
                                  gdtb_solver_info* info);
/* the CSR pattern the operator's values follow (materialised on first use for the closed-form CG Q1 operator) */
int gdtb_matop_pattern_device(gdtb_matop* op, const int64_t** d_rowptr, const int32_t** d_colidx);

/* ---- BilinearForm::apply2 for norms (operators/bilinear-form.hh:35-470, examples/stationary-heat-equation.cc:
 * 116-127) ------------------------------------------------------------------------------------------------ */
/* result = sum over elements of the element form applied to (e, e), e = u_h - f, u_h the discrete function with
 * DoF vector d_dofs in `space` (NULL: e = -f) and f an analytic function (NULL: e = u_h): the function is used as a
 * one-function test and ansatz "basis" (bilinear-form.hh:340-352).  form: LocalLaplaceIntegrand (H^1 semi-norm^2) /
 * LocalElementProductIntegrand (L^2 norm^2) or sums; quadrature order = integrand order with test = ansatz order
 * = max(space order, f->order) plus over_integrate.  Deterministic (fixed-order two-stage reduction). */
/* order of the rule gdtb_bilinear_form_apply2 integrates with (f_order: declared order of f, -1 without f) */
int gdtb_bilinear_form_quadrature_order(const gdtb_space* space, int has_dofs, int f_order, const gdtb_form* form,
                                        int32_t* order);
int gdtb_bilinear_form_apply2(gdtb_ctx* ctx, const gdtb_space* space, const double* d_dofs, const gdtb_function* f,
                              const gdtb_form* form, double* result);
int gdtb_bilinear_form_apply2_host(gdtb_ctx* ctx, const gdtb_space* space, const double* dofs, const gdtb_function* f,
                                   const gdtb_form* form, double* result);
/* default_interpolation(f, lagrange_space) (interpolations/default.hh:40-83): DoFs = f at the Lagrange points */
int gdtb_lagrange_interpolate(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_function* f, double* d_dofs);
int gdtb_lagrange_interpolate_host(gdtb_ctx* ctx, const gdtb_space* space, const gdtb_function* f, double* dofs);

#ifdef __cplusplus
}
#endif
#endif /* GDTB_H */
