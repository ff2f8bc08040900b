"""Host-side Python mirror of the dune-gdt interface for the assembly / FV-apply hot path.

Same names, argument meaning and error behaviour as the reference classes (cited per class), lowered
through the C ABI (include/gdtb.h) to the CUDA kernels in csrc/.  The C++ facade with the same names lives in
include/dune/gdt/; this module exists so that tests and bench.py read like the reference's own drivers
(the reference ships pybind11 bindings of the same classes: python/dune/gdt/__init__.py:17-82).
"""
import ctypes as C
import enum

import numpy as np

from . import capi, descriptors as D
from .capi import check, dptr, lib


class Stencil(enum.IntEnum):
    """Dune::GDT::Stencil (dune/gdt/type_traits.hh:55-60)"""

    element = D.STENCIL_ELEMENT
    intersection = D.STENCIL_INTERSECTION
    element_and_intersection = D.STENCIL_ELEMENT_AND_INTERSECTION


class ApplyOn(enum.IntEnum):
    """XT::Grid::ApplyOn intersection filters used by the hot path"""

    InnerIntersectionsOnce = D.FILTER_INNER_ONCE
    InnerAndPeriodicIntersectionsOnce = D.FILTER_INNER_AND_PERIODIC_ONCE
    AllDirichletBoundaryIntersections = D.FILTER_ALL_BOUNDARY


class Context:
    """One per process and GPU: CUDA device + stream all handles of this module work on."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().gdtb_ctx_create(device, C.byref(self._h)))
        self.device = device
        self.stream_handle = None  # the cudaStream_t handed to set_stream (None: the context's own stream)

    def synchronize(self):
        check(lib().gdtb_ctx_synchronize(self._h))

    def set_stream(self, cuda_stream):
        """run the library on a caller-provided cudaStream_t (0 / None: back to the context's own stream)"""
        check(lib().gdtb_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))
        self.stream_handle = cuda_stream or None

    @property
    def launch_count(self):
        return lib().gdtb_ctx_launch_count(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_ctx_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass
            self._h = C.c_void_p()


class Grid:
    def __init__(self, ctx, desc):
        self.ctx, self.desc = ctx, desc
        self._h = C.c_void_p()
        check(lib().gdtb_grid_create_cube(ctx._h, C.byref(desc), C.byref(self._h)))

    @property
    def dim(self):
        return self.desc.dim

    @property
    def num_elements(self):
        return lib().gdtb_grid_num_elements(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_grid_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass


def make_cube_grid(ctx, lower, upper, num_elements, periodic=0):
    """XT::Grid::make_cube_grid<YASP_dD_EQUIDISTANT_OFFSET>(lower, upper, n) (+ make_periodic_grid_view)
    -- examples/stationary-heat-equation.cc:87, examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:260-261"""
    return Grid(ctx, D.grid_desc(lower, upper, num_elements, periodic))


class _Mapper:
    """MapperInterface (dune/gdt/spaces/mapper/interfaces.hh)"""

    def __init__(self, space):
        self._s = space

    @property
    def size(self):
        return lib().gdtb_space_size(self._s._h)

    @property
    def max_local_size(self):
        return lib().gdtb_space_max_local_size(self._s._h)

    def global_indices(self, element):
        out = np.zeros(self.max_local_size, dtype=np.int64)
        check(lib().gdtb_space_global_indices(self._s._h, element, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out


class Space:
    def __init__(self, grid, kind, order, range_dim=1):
        self.grid, self.kind, self.order, self.range_dim = grid, kind, order, range_dim
        self._h = C.c_void_p()
        if range_dim != 1:
            if kind != D.SPACE_FV:
                raise capi.NotImplementedGdt("vector-valued spaces: finite volume spaces only")
            check(lib().gdtb_fv_space_create(grid.ctx._h, grid._h, int(range_dim), C.byref(self._h)))
        else:
            check(lib().gdtb_space_create(grid.ctx._h, grid._h, kind, order, C.byref(self._h)))
        self.mapper = _Mapper(self)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_space_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass


def make_continuous_lagrange_space(grid, order):
    """dune/gdt/spaces/h1/continuous-lagrange.hh:189-203"""
    return Space(grid, D.SPACE_CG, order)


def make_discontinuous_lagrange_space(grid, order):
    """dune/gdt/spaces/l2/discontinuous-lagrange.hh:191-205"""
    return Space(grid, D.SPACE_DG, order)


def make_finite_volume_space(grid, range_dim=1):
    """make_finite_volume_space<m>(grid_view) (dune/gdt/spaces/l2/finite-volume.hh:208-230): m = range_dim DoFs per element"""
    return Space(grid, D.SPACE_FV, 0, range_dim)


class SparsityPattern:
    """XT::LA::SparsityPatternDefault as produced by make_sparsity_pattern (dune/gdt/tools/sparsity-pattern.hh:163-178)"""

    def __init__(self, test_space, ansatz_space, stencil, method=D.PATTERN_AUTO):
        self.test_space, self.ansatz_space = test_space, ansatz_space
        self._h = C.c_void_p()
        check(
            lib().gdtb_pattern_create(
                test_space.grid.ctx._h, test_space._h, ansatz_space._h, int(stencil), method, C.byref(self._h)
            )
        )

    @property
    def rows(self):
        return lib().gdtb_pattern_rows(self._h)

    @property
    def cols(self):
        return lib().gdtb_pattern_cols(self._h)

    @property
    def nnz(self):
        return lib().gdtb_pattern_nnz(self._h)

    def download(self):
        rowptr = np.empty(self.rows + 1, dtype=np.int64)
        colidx = np.empty(self.nnz, dtype=np.int32)
        check(
            lib().gdtb_pattern_download(
                self._h, rowptr.ctypes.data_as(C.POINTER(C.c_int64)), colidx.ctypes.data_as(C.POINTER(C.c_int32))
            )
        )
        return rowptr, colidx

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_pattern_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass


def make_sparsity_pattern(test_space, ansatz_space=None, stencil=Stencil.element, method=D.PATTERN_AUTO):
    return SparsityPattern(test_space, ansatz_space or test_space, stencil, method)


# ---- grid functions -----------------------------------------------------------------------------
def GridFunction(value, order=0):
    """XT::Functions::GridFunction: constant scalar / tensor, per-element numpy array, or a descriptors.Function"""
    if isinstance(value, D.Function):
        return value
    if isinstance(value, np.ndarray) and value.ndim >= 1 and value.size > 9:
        return D.fn_elem(value, order=order)
    return D.fn_const(value, order=order)


# ---- integrands (value types, summable) -------------------------------------------------------------
class _Integrand:
    role = None  # "element" | "coupling" | "boundary"

    def __init__(self, terms):
        self.terms = terms

    def __add__(self, other):
        """integrand_a + integrand_b (dune/gdt/local/integrands/interfaces.hh:233-236,480-483,609-612)"""
        if self.role != other.role:
            raise capi.IntegrandError("cannot add integrands of different kinds")
        s = _Integrand(self.terms + other.terms)
        s.role = self.role
        return s


class LocalLaplaceIntegrand(_Integrand):
    """dune/gdt/local/integrands/laplace.hh:40-48"""

    role = "element"

    def __init__(self, diffusion=1.0):
        super().__init__([D.integrand(D.INT_LAPLACE, diffusion=GridFunction(diffusion))])


class LocalProductIntegrand(_Integrand):
    """dune/gdt/local/integrands/product.hh:56-65, 389-404"""

    role = "element"

    def __init__(self, weight=1.0):
        self._weight = GridFunction(weight)
        super().__init__([D.integrand(D.INT_PRODUCT, diffusion=self._weight)])

    def with_ansatz(self, function, order=0):
        """LocalBinaryToUnaryElementIntegrand: bi(f, .) (integrands/interfaces.hh:225-229, conversion.hh:42-124)"""
        u = _Integrand([D.integrand(D.INT_PRODUCT, diffusion=self._weight, weight=GridFunction(function, order))])
        u.role = "unary_element"
        return u


LocalElementProductIntegrand = LocalProductIntegrand


class LocalLaplaceIPDGIntegrands:
    """dune/gdt/local/integrands/laplace-ipdg.hh"""

    class InnerCoupling(_Integrand):
        role = "coupling"

        def __init__(self, symmetry_prefactor, diffusion, weight=1.0):
            super().__init__(
                [
                    D.integrand(
                        D.INT_IPDG_INNER_COUPLING,
                        diffusion=GridFunction(diffusion),
                        weight=GridFunction(weight),
                        prefactor=symmetry_prefactor,
                    )
                ]
            )

    class DirichletCoupling(_Integrand):
        role = "boundary"

        def __init__(self, symmetry_prefactor, diffusion):
            super().__init__(
                [
                    D.integrand(
                        D.INT_IPDG_DIRICHLET_COUPLING, diffusion=GridFunction(diffusion), prefactor=symmetry_prefactor
                    )
                ]
            )


class LocalIPDGIntegrands:
    """dune/gdt/local/integrands/ipdg.hh"""

    class InnerPenalty(_Integrand):
        role = "coupling"

        def __init__(self, penalty, weight=1.0, intersection_diameter=D.HI_DIAMETER):
            super().__init__(
                [
                    D.integrand(
                        D.INT_IPDG_INNER_PENALTY,
                        weight=GridFunction(weight),
                        prefactor=penalty,
                        hI_kind=intersection_diameter,
                    )
                ]
            )

    class BoundaryPenalty(_Integrand):
        role = "boundary"

        def __init__(self, penalty, weight=1.0, intersection_diameter=D.HI_DIAMETER):
            super().__init__(
                [
                    D.integrand(
                        D.INT_IPDG_BOUNDARY_PENALTY,
                        weight=GridFunction(weight),
                        prefactor=penalty,
                        hI_kind=intersection_diameter,
                    )
                ]
            )


# ---- local forms ------------------------------------------------------------------------------------
class _LocalForm:
    role = None

    def __init__(self, integrand, over_integrate=0):
        if integrand.role != self.role:
            raise capi.IntegrandError(f"{type(self).__name__} needs a {self.role} integrand, got {integrand.role}")
        self.integrand, self.over_integrate = integrand, over_integrate

    def descriptor(self, scaling=1.0):
        return D.form(self.integrand.terms, self.over_integrate, scaling)


class LocalElementIntegralBilinearForm(_LocalForm):
    """dune/gdt/local/bilinear-forms/integrals.hh:38-140"""

    role = "element"


class LocalCouplingIntersectionIntegralBilinearForm(_LocalForm):
    """dune/gdt/local/bilinear-forms/integrals.hh:154-276"""

    role = "coupling"


class LocalIntersectionIntegralBilinearForm(_LocalForm):
    """dune/gdt/local/bilinear-forms/integrals.hh:290-375"""

    role = "boundary"


class LocalElementIntegralFunctional(_LocalForm):
    """dune/gdt/local/functionals/integrals.hh:27-104"""

    role = "unary_element"


# ---- containers ---------------------------------------------------------------------------------------
class CsrMatrix:
    """Host copy in the ISTL/Eigen row-major layout: rowptr int64, colidx int32, values float64"""

    def __init__(self, rows, cols, rowptr, colidx, values):
        self.rows, self.cols = rows, cols
        self.rowptr, self.colidx, self.values = rowptr, colidx, values

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.values, self.colidx, self.rowptr), shape=(self.rows, self.cols))


class VectorBasedFunctional:
    """dune/gdt/functionals/vector-based.hh:133-286"""

    def __init__(self, space):
        self.space = space
        self._h = C.c_void_p()
        check(lib().gdtb_vecfun_create(space.grid.ctx._h, space._h, C.byref(self._h)))
        self._fresh = True

    def append(self, local_functional):
        if not isinstance(local_functional, LocalElementIntegralFunctional):
            raise capi.WrongInputGiven("append() takes a LocalElementIntegralFunctional")
        desc = local_functional.descriptor()
        check(lib().gdtb_vecfun_append_element(self._h, C.byref(desc)))
        return self

    def assemble(self, use_tbb=False):
        mode = D.ASSEMBLE_OVERWRITE if self._fresh else D.ASSEMBLE_ACCUMULATE
        check(lib().gdtb_assemble(None, self._h, mode))
        self._fresh = False
        check(lib().gdtb_vecfun_clear_forms(self._h))

    def vector(self):
        out = np.empty(self.space.mapper.size, dtype=np.float64)
        check(lib().gdtb_vecfun_download(self._h, dptr(out)))
        return out

    def device_pointer(self):
        p = C.c_void_p()
        check(lib().gdtb_vecfun_device(self._h, C.byref(p)))
        return p.value

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_vecfun_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass


def make_vector_functional(space):
    return VectorBasedFunctional(space)


class MatrixOperator:
    """dune/gdt/operators/matrix-based.hh:245-508: matrix storage + operator + grid walker"""

    def __init__(self, test_space, ansatz_space, pattern):
        self.test_space, self.ansatz_space, self.pattern = test_space, ansatz_space, pattern
        self.scaling = 1.0  # captured by value at append time (matrix-based.hh:342,365)
        self._h = C.c_void_p()
        check(
            lib().gdtb_matop_create(
                test_space.grid.ctx._h, test_space._h, ansatz_space._h, pattern._h, C.byref(self._h)
            )
        )
        self._functionals = []
        self._fresh = True

    def append(self, what, param=None, filter=None):
        if isinstance(what, LocalElementIntegralBilinearForm):
            desc = what.descriptor(self.scaling)
            check(lib().gdtb_matop_append_element(self._h, C.byref(desc)))
        elif isinstance(what, LocalCouplingIntersectionIntegralBilinearForm):
            desc = what.descriptor(self.scaling)
            flt = ApplyOn.InnerIntersectionsOnce if filter is None else filter
            check(lib().gdtb_matop_append_coupling(self._h, C.byref(desc), int(flt)))
        elif isinstance(what, LocalIntersectionIntegralBilinearForm):
            desc = what.descriptor(self.scaling)
            flt = ApplyOn.AllDirichletBoundaryIntersections if filter is None else filter
            check(lib().gdtb_matop_append_boundary(self._h, C.byref(desc), int(flt)))
        elif isinstance(what, VectorBasedFunctional):
            # operator-as-walker: assemble the functional in the same grid walk
            # (examples/adaptive_elliptic_swipdg.cc:250)
            self._functionals.append(what)
        else:
            raise capi.WrongInputGiven(f"cannot append {type(what).__name__}")
        return self

    __iadd__ = append

    @property
    def plan(self):
        return lib().gdtb_matop_plan(self._h).decode()

    @property
    def plan_reason(self):
        """why the operator takes the generic kernels ("" on a row-gather path)"""
        return lib().gdtb_matop_plan_reason(self._h).decode()

    def assemble(self, use_tbb=False):
        """one grid walk (matrix-based.hh:496-500); afterwards the functor list is empty like after Walker::walk"""
        if len(self._functionals) > 1:
            raise capi.NotImplementedGdt("at most one functional can ride along with a matrix operator")
        fun = self._functionals[0] if self._functionals else None
        if fun is not None and fun._fresh != self._fresh:
            check(lib().gdtb_assemble(self._h, None, D.ASSEMBLE_OVERWRITE if self._fresh else D.ASSEMBLE_ACCUMULATE))
            fun.assemble()
        else:
            mode = D.ASSEMBLE_OVERWRITE if self._fresh else D.ASSEMBLE_ACCUMULATE
            check(lib().gdtb_assemble(self._h, fun._h if fun is not None else None, mode))
            if fun is not None:
                fun._fresh = False
                check(lib().gdtb_vecfun_clear_forms(fun._h))
        self._fresh = False
        self._functionals = []
        check(lib().gdtb_matop_clear_forms(self._h))

    def values(self):
        out = np.empty(self.pattern.nnz, dtype=np.float64)
        check(lib().gdtb_matop_values_download(self._h, dptr(out)))
        return out

    def matrix(self):
        rowptr, colidx = self.pattern.download()
        return CsrMatrix(self.pattern.rows, self.pattern.cols, rowptr, colidx, self.values())

    def device_pointer(self):
        p = C.c_void_p()
        check(lib().gdtb_matop_values_device(self._h, C.byref(p)))
        return p.value

    def apply(self, source):
        """ConstMatrixOperator::apply: range = matrix * source (matrix-based.hh:121-129); host numpy in / out"""
        src = np.ascontiguousarray(source, dtype=np.float64)
        if src.size != self.ansatz_space.mapper.size:
            raise capi.OperatorError("when applying matrix to source and range dofs: shapes do not match")
        out = np.empty(self.test_space.mapper.size, dtype=np.float64)
        check(lib().gdtb_matop_apply_host(self._h, dptr(src), dptr(out)))
        return out

    def apply_inverse(self, range_vector, opts=None, initial_guess=None):
        """ConstMatrixOperator::apply_inverse (matrix-based.hh:148-159): solves matrix * x = range on the device;
        returns (x, info) with info = descriptors.SolverInfo"""
        rhs = np.ascontiguousarray(range_vector, dtype=np.float64)
        if rhs.size != self.test_space.mapper.size:
            raise capi.OperatorError("when applying linear solver: shapes do not match")
        x = np.zeros_like(rhs) if initial_guess is None else np.array(initial_guess, dtype=np.float64, copy=True)
        info = D.SolverInfo()
        o = opts if opts is not None else D.solver_opts()
        check(lib().gdtb_matop_apply_inverse_host(self._h, dptr(rhs), dptr(x), C.byref(o), C.byref(info)))
        return x, info

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_matop_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass


def make_matrix_operator(space, stencil=Stencil.element, ansatz_space=None, pattern=None):
    """make_matrix_operator<M>(space, stencil) (dune/gdt/operators/matrix-based.hh:611-658)"""
    ansatz_space = ansatz_space or space
    if pattern is None:
        pattern = make_sparsity_pattern(space, ansatz_space, stencil)
    return MatrixOperator(space, ansatz_space, pattern)


# ---- finite volumes -----------------------------------------------------------------------------------
class NumericalUpwindFlux:
    """dune/gdt/local/numerical-fluxes/upwind.hh:31-80 with a built-in flux function"""

    numflux = D.NUMFLUX_UPWIND

    def __init__(self, flux_kind, params=()):
        self.desc = D.flux(flux_kind, self.numflux, params)


class NumericalLaxFriedrichsFlux(NumericalUpwindFlux):
    """dune/gdt/local/numerical-fluxes/lax-friedrichs.hh:60-88"""

    numflux = D.NUMFLUX_LAX_FRIEDRICHS


class NumericalVijayasundaramFlux(NumericalUpwindFlux):
    """dune/gdt/local/numerical-fluxes/vijayasundaram.hh:29-153 with the flux's own eigendecomposition (EulerTools, the
    lambda of examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:402-409); systems only"""

    numflux = D.NUMFLUX_VIJAYASUNDARAM


class EulerTools:
    """dune/gdt/tools/euler.hh: conversions between primitive and conservative variables (host side; flux, jacobian and
    eigendecomposition live in the device kernels) and the flux descriptor of the Euler equations"""

    def __init__(self, d, gamma):
        self.d, self.gamma, self.m = int(d), float(gamma), int(d) + 2

    def flux(self):
        """params of a Numerical*Flux: D.FLUX_EULER with p[0] = gamma"""
        return D.FLUX_EULER, [self.gamma]

    def conservative(self, density, velocity, pressure):
        """(rho, rho v, E), E = p / (gamma - 1) + rho |v|^2 / 2 (euler.hh:127-131, 153-157)"""
        v = np.atleast_1d(np.asarray(velocity, dtype=np.float64))
        if v.size == 1 and self.d > 1:
            v = np.full(self.d, float(v[0]))
        return np.concatenate([[density], density * v, [pressure / (self.gamma - 1.0) + 0.5 * density * float(v @ v)]])

    def primitive(self, w):
        """(rho, v, p) (euler.hh:133-147); w of shape (..., m)"""
        w = np.asarray(w, dtype=np.float64)
        rho = w[..., 0]
        v = w[..., 1:1 + self.d] / rho[..., None]
        p = (self.gamma - 1.0) * (w[..., self.m - 1] - 0.5 * rho * (v * v).sum(-1))
        return rho, v, p


class AdvectionFvOperator:
    """dune/gdt/operators/advection-fv.hh:44-141"""

    def __init__(self, numerical_flux, source_space, range_space=None):
        range_space = range_space or source_space
        if range_space is not source_space:
            raise capi.NotImplementedGdt("source and range space must be the same finite volume space")
        self.space = source_space
        self._h = C.c_void_p()
        check(
            lib().gdtb_fvop_create(
                source_space.grid.ctx._h, source_space._h, C.byref(numerical_flux.desc), C.byref(self._h)
            )
        )

    def apply(self, source):
        """V apply(const V& source) (operators/interfaces.hh:645-650); host numpy in, host numpy out"""
        src = np.ascontiguousarray(source, dtype=np.float64)
        if src.size != self.space.mapper.size:
            raise capi.ShapesDoNotMatch("source vector has the wrong size")
        out = np.empty_like(src)
        check(lib().gdtb_fvop_apply_host(self._h, dptr(src), dptr(out)))
        return out

    def apply_device(self, d_source, d_range):
        check(lib().gdtb_fvop_apply(self._h, C.c_void_p(d_source), C.c_void_p(d_range)))

    def explicit_euler(self, u, dt, n_steps):
        """u <- u - L(u) dt, n_steps times (examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:152-157)"""
        u = np.array(u, dtype=np.float64, copy=True)
        check(lib().gdtb_fvop_euler_host(self._h, dptr(u), float(dt), int(n_steps)))
        return u

    def euler_device(self, d_u, dt, n_steps):
        check(lib().gdtb_fvop_euler(self._h, C.c_void_p(d_u), float(dt), int(n_steps)))

    def append(self, treatment):
        """append(boundary treatment, param_type, filter) (operators/advection-fv.hh:96-123); `treatment` is a
        descriptors.fv_boundary(kind, side_mask, a, b)"""
        check(lib().gdtb_fvop_append_boundary(self._h, C.byref(treatment)))
        return self

    def estimate_dt(self, state, boundary_data_range=None):
        """estimate_dt_for_hyperbolic_system(grid_view, state, flux, boundary_data_range) (tools/hyperbolic.hh:38-86)"""
        u = np.ascontiguousarray(state, dtype=np.float64)
        if u.size != self.space.mapper.size:
            raise capi.ShapesDoNotMatch("state vector has the wrong size")
        rng = None if boundary_data_range is None else np.ascontiguousarray(boundary_data_range, dtype=np.float64)
        dt = C.c_double()
        check(lib().gdtb_fv_estimate_dt_host(self._h, dptr(u), None if rng is None else dptr(rng), C.byref(dt)))
        return dt.value

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_fvop_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass


def make_advection_fv_operator(numerical_flux, source_space, range_space=None):
    """make_advection_fv_operator<M>(view, numerical_flux, source_space, range_space) (advection-fv.hh:130-141)"""
    return AdvectionFvOperator(numerical_flux, source_space, range_space)


def estimate_dt_for_hyperbolic_system(op, state, boundary_data_range=None):
    """tools/hyperbolic.hh:38-86 (grid view and flux are the operator's)"""
    return op.estimate_dt(state, boundary_data_range)


class TimeStepperMethods:
    """tools/timestepper/interface.hh TimeStepperMethods (the explicit Runge-Kutta members)"""

    explicit_euler = D.RK_EULER
    explicit_rungekutta_second_order_ssp = D.RK_SSP2
    explicit_rungekutta_third_order_ssp = D.RK_SSP3
    explicit_rungekutta_classic_fourth_order = D.RK_CLASSIC4
    explicit_rungekutta_other = D.RK_OTHER


class ExplicitRungeKuttaTimeStepper:
    """tools/timestepper/explicit-rungekutta.hh:158-270: u_t = r L(u); the stepper owns the current solution (a host
    copy of initial_values, stepped on the device)"""

    def __init__(self, op, initial_values, r=1.0, t_0=0.0, method=TimeStepperMethods.explicit_euler, A=None, b=None, c=None):
        self.op = op
        self._u = np.array(initial_values, dtype=np.float64, copy=True)
        if self._u.size != op.space.mapper.size:
            raise capi.ShapesDoNotMatch("initial values have the wrong size")
        self._h = C.c_void_p()
        s = 0
        arrs = [None, None, None]
        if method == D.RK_OTHER and A is not None:
            arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (A, b, c)]
            s = arrs[1].size
            if arrs[0].shape != (s, s) or arrs[2].size != s:
                raise capi.ShapesDoNotMatch("Butcher array: A must be s x s, b and c of size s")
        self._keep = arrs
        check(lib().gdtb_rk_create(op._h, int(method), s, *[None if x is None else dptr(x) for x in arrs], float(r), float(t_0),
                                   C.byref(self._h)))
        self.dts = []

    def current_time(self):
        return lib().gdtb_rk_current_time(self._h)

    def current_solution(self):
        return self._u

    def step(self, dt, max_dt=None):
        """step(dt, max_dt) (:237-270); returns dt"""
        ret = C.c_double()
        check(lib().gdtb_rk_step_host(self._h, dptr(self._u), float(dt), float(dt if max_dt is None else max_dt), C.byref(ret)))
        self.dts.append(min(dt, dt if max_dt is None else max_dt))
        return ret.value

    def solve(self, t_end, initial_dt):
        """TimeStepperInterface::solve(t_end, initial_dt) (tools/timestepper/interface.hh:191-263); returns the dt
        for the next step; .num_steps holds the number of steps taken"""
        n = C.c_int64()
        nxt = C.c_double()
        check(lib().gdtb_rk_solve_host(self._h, dptr(self._u), float(t_end), float(initial_dt), C.byref(n), C.byref(nxt)))
        self.num_steps = n.value
        return nxt.value

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_rk_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass


# ---- callers on either side of the hot path ---------------------------------------------
class DirichletConstraints:
    """dune/gdt/tools/dirichlet-constraints.hh:44-230; boundary_mask bit (2k+s) = domain face with outer normal
    -e_k / +e_k is a Dirichlet boundary (descriptors.BOUNDARY_ALL = XT::Grid::AllDirichletBoundaryInfo)"""

    def __init__(self, space, boundary_mask=D.BOUNDARY_ALL):
        self.space = space
        self._h = C.c_void_p()
        check(lib().gdtb_dirichlet_create(space.grid.ctx._h, space._h, int(boundary_mask), C.byref(self._h)))

    def dirichlet_DoFs(self):
        out = np.empty(lib().gdtb_dirichlet_size(self._h), dtype=np.int64)
        check(lib().gdtb_dirichlet_dofs_download(self._h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def apply(self, matrix_operator=None, functional=None, only_clear=False, ensure_symmetry=True):
        """apply(matrix, vector, only_clear, ensure_symmetry) (dirichlet-constraints.hh:122-184)"""
        check(
            lib().gdtb_dirichlet_apply(
                self._h,
                matrix_operator._h if matrix_operator is not None else None,
                functional._h if functional is not None else None,
                int(only_clear),
                int(ensure_symmetry),
            )
        )

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            try:
                lib().gdtb_dirichlet_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass


def make_dirichlet_constraints(space, boundary_info=D.BOUNDARY_ALL):
    """make_dirichlet_constraints(space, boundary_info) (examples/stationary-heat-equation.cc:100)"""
    return DirichletConstraints(space, boundary_info)


class BilinearForm:
    """dune/gdt/operators/bilinear-form.hh:35-470 for (error, error) products: error = discrete function - exact"""

    def __init__(self, space, dofs=None, minus=None):
        self.space, self.dofs, self.minus = space, dofs, minus
        self._forms = []

    def append(self, local_form):
        if not isinstance(local_form, (LocalElementIntegralBilinearForm, D.Form)):
            raise capi.WrongInputGiven("only element bilinear forms are supported by apply2")
        self._forms.append(local_form)
        return self

    __iadd__ = append

    def apply2(self):
        total = 0.0
        ctx = self.space.grid.ctx
        for lf in self._forms:
            desc = lf if isinstance(lf, D.Form) else lf.descriptor()
            res = C.c_double()
            dofs = None if self.dofs is None else np.ascontiguousarray(self.dofs, dtype=np.float64)
            check(
                lib().gdtb_bilinear_form_apply2_host(
                    ctx._h, self.space._h, dptr(dofs) if dofs is not None else None,
                    C.byref(self.minus) if self.minus is not None else None, C.byref(desc), C.byref(res),
                )
            )
            total += res.value
        return total


def make_bilinear_form(space, dofs=None, minus=None):
    return BilinearForm(space, dofs, minus)


def default_interpolation(function, space, order=None):
    """default_interpolation(f, space) (interpolations/default.hh:40-83): Lagrange spaces = nodal values, finite-volume
    spaces = cell averages by a rule of the declared order"""
    f = GridFunction(function, order or 0)
    out = np.empty(space.mapper.size, dtype=np.float64)
    ctx = space.grid.ctx
    if space.kind == D.SPACE_FV:
        check(lib().gdtb_fv_interpolate_host(ctx._h, space._h, C.byref(f), dptr(out)))
    else:
        check(lib().gdtb_lagrange_interpolate_host(ctx._h, space._h, C.byref(f), dptr(out)))
    return out
