"""ctypes mirrors of the POD descriptors in include/gdtb.h.

The oracle (oracle/oracle.h, test infrastructure) deliberately uses the same struct layouts, so the
parity tests build ONE descriptor and hand it to both sides.
"""
import ctypes as C

import numpy as np

# ---- enums (include/gdtb.h) -------------------------------------------------------------------
SPACE_CG, SPACE_DG, SPACE_FV = 0, 1, 2
STENCIL_ELEMENT, STENCIL_INTERSECTION, STENCIL_ELEMENT_AND_INTERSECTION = 0, 1, 2
FN_CONST_SCALAR, FN_CONST_TENSOR, FN_ELEM_SCALAR, FN_ELEM_TENSOR, FN_BUILTIN = 0, 1, 2, 3, 4
FN_QP_SCALAR, FN_QP_TENSOR, FN_DOF_VECTOR, FN_QP_VALUE_GRAD = 5, 6, 7, 8
ROLE_ELEMENT, ROLE_FUNCTIONAL, ROLE_COUPLING, ROLE_BOUNDARY = 0, 1, 2, 3
BUILTIN_COS_PRODUCT, BUILTIN_AFFINE, BUILTIN_GAUSSIAN, BUILTIN_INDICATOR, BUILTIN_QUADRATIC = 1, 2, 3, 4, 5
INT_LAPLACE, INT_PRODUCT = 0, 1
INT_IPDG_INNER_COUPLING, INT_IPDG_INNER_PENALTY = 2, 3
INT_IPDG_DIRICHLET_COUPLING, INT_IPDG_BOUNDARY_PENALTY = 4, 5
HI_DIAMETER, HI_VOLUME = 0, 1
FILTER_INNER_ONCE, FILTER_INNER_AND_PERIODIC_ONCE, FILTER_ALL_BOUNDARY = 0, 1, 2
FLUX_LINEAR, FLUX_BURGERS, FLUX_EULER = 0, 1, 2
NUMFLUX_UPWIND, NUMFLUX_LAX_FRIEDRICHS, NUMFLUX_VIJAYASUNDARAM = 0, 1, 2
ASSEMBLE_OVERWRITE, ASSEMBLE_ACCUMULATE = 0, 1
PATTERN_AUTO, PATTERN_SORT_UNIQUE, PATTERN_STRUCTURED = 0, 1, 2
SOLVER_CG, SOLVER_BICGSTAB = 0, 1
PRECOND_NONE, PRECOND_JACOBI = 0, 1
BOUNDARY_ALL = 0x3F
MAX_TERMS = 4


class GridDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("periodic", C.c_int32),
        ("lower", C.c_double * 3),
        ("upper", C.c_double * 3),
        ("n", C.c_int64 * 3),
    ]


class Function(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("order", C.c_int32),
        ("builtin", C.c_int32),
        ("data_on_device", C.c_int32),
        ("c", C.c_double * 9),
        ("p", C.c_double * 8),
        ("data", C.POINTER(C.c_double)),
        ("qp_per_element", C.c_int32),
        ("space_kind", C.c_int32),
        ("space_order", C.c_int32),
        ("reserved", C.c_int32),
    ]


class Integrand(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("hI_kind", C.c_int32),
        ("prefactor", C.c_double),
        ("diffusion", Function),
        ("weight", Function),
    ]


class Form(C.Structure):
    _fields_ = [
        ("n_terms", C.c_int32),
        ("over_integrate", C.c_int32),
        ("scaling", C.c_double),
        ("terms", Integrand * MAX_TERMS),
    ]


class Flux(C.Structure):
    _fields_ = [("kind", C.c_int32), ("numflux", C.c_int32), ("p", C.c_double * 4)]


class FvBoundary(C.Structure):
    """gdtb_fv_boundary / orc_fv_boundary: one boundary treatment of the FV operator (operators/advection-fv.hh:96-123)"""

    _fields_ = [("kind", C.c_int32), ("side_mask", C.c_uint32), ("a", C.c_double), ("b", C.c_double)]


class SolverOpts(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("preconditioner", C.c_int32),
        ("max_iter", C.c_int32),
        ("check_every", C.c_int32),
        ("precision", C.c_double),
    ]


class SolverInfo(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32),
        ("converged", C.c_int32),
        ("initial_residual", C.c_double),
        ("residual", C.c_double),
    ]


# ---- constructors -----------------------------------------------------------------------------
def grid_desc(lower, upper, n, periodic=0):
    """XT::Grid::make_cube_grid(lower, upper, n); scalars broadcast over `dim` = len(n)."""
    n = list(np.atleast_1d(n))
    dim = len(n)
    lower = list(np.broadcast_to(np.asarray(lower, dtype=float), (dim,)))
    upper = list(np.broadcast_to(np.asarray(upper, dtype=float), (dim,)))
    g = GridDesc()
    g.dim = dim
    g.periodic = int(periodic)
    for k in range(3):
        g.lower[k] = lower[k] if k < dim else 0.0
        g.upper[k] = upper[k] if k < dim else 1.0
        g.n[k] = int(n[k]) if k < dim else 1
    return g


def _keep(obj, *refs):
    obj._refs = getattr(obj, "_refs", []) + list(refs)
    return obj


def fn_const(value, order=0):
    """A constant grid function: scalar, or a d x d tensor (nested list / array)."""
    f = Function()
    v = np.asarray(value, dtype=float)
    f.order = order
    if v.ndim == 0:
        f.kind = FN_CONST_SCALAR
        f.c[0] = float(v)
    else:
        f.kind = FN_CONST_TENSOR
        flat = v.reshape(-1)
        for i, x in enumerate(flat):
            f.c[i] = float(x)
    return f


def fn_elem(values, dim=None, order=0):
    """Element-wise constant function: array of n_elem scalars or (n_elem, d, d) tensors (host memory)."""
    a = np.ascontiguousarray(values, dtype=np.float64)
    f = Function()
    f.order = order
    f.kind = FN_ELEM_SCALAR if a.ndim == 1 else FN_ELEM_TENSOR
    f.data = a.ctypes.data_as(C.POINTER(C.c_double))
    return _keep(f, a)


def fn_builtin(builtin, order, *params):
    f = Function()
    f.kind = FN_BUILTIN
    f.builtin = builtin
    f.order = order
    for i, x in enumerate(params):
        f.p[i] = float(x)
    return f


def fn_qp(values, order):
    """Coefficient data sampled at the quadrature points of the form it will be appended to: array of shape
    (n_elem, n_qp) (scalar) or (n_elem, n_qp, d, d) (tensor), q = q_0 + m (q_1 + m q_2); `order` is the declared
    polynomial order (it enters the form's quadrature order, which in turn fixes n_qp)."""
    a = np.ascontiguousarray(values, dtype=np.float64)
    assert a.ndim in (2, 4)
    f = Function()
    f.order = int(order)
    f.kind = FN_QP_SCALAR if a.ndim == 2 else FN_QP_TENSOR
    f.qp_per_element = a.shape[1]
    f.data = a.ctypes.data_as(C.POINTER(C.c_double))
    return _keep(f, a)


def fn_dofs(dofs, space_kind, space_order):
    """A discrete function of the space (space_kind, space_order) on the form's grid, given by its DoF vector (host)."""
    a = np.ascontiguousarray(dofs, dtype=np.float64)
    f = Function()
    f.kind = FN_DOF_VECTOR
    f.order = int(space_order)
    f.space_kind, f.space_order = int(space_kind), int(space_order)
    f.data = a.ctypes.data_as(C.POINTER(C.c_double))
    return _keep(f, a)


def _as_function(x):
    return x if isinstance(x, Function) else fn_const(x)


def _copy_fn(dst, src):
    C.memmove(C.byref(dst), C.byref(src), C.sizeof(Function))


def integrand(kind, diffusion=1.0, weight=1.0, prefactor=0.0, hI_kind=HI_DIAMETER):
    it = Integrand()
    it.kind = kind
    it.hI_kind = hI_kind
    it.prefactor = float(prefactor)
    d, w = _as_function(diffusion), _as_function(weight)
    _copy_fn(it.diffusion, d)
    _copy_fn(it.weight, w)
    return _keep(it, d, w)


def form(terms, over_integrate=0, scaling=1.0):
    if isinstance(terms, Integrand):
        terms = [terms]
    f = Form()
    f.n_terms = len(terms)
    f.over_integrate = int(over_integrate)
    f.scaling = float(scaling)
    for i, t in enumerate(terms):
        C.memmove(C.byref(f.terms[i]), C.byref(t), C.sizeof(Integrand))
    return _keep(f, *terms)


def flux(kind, numflux=NUMFLUX_UPWIND, params=()):
    fl = Flux()
    fl.kind = kind
    fl.numflux = numflux
    for i, x in enumerate(params):
        fl.p[i] = float(x)
    return fl


FVBND_EXTRAPOLATION, FVBND_NUMERICAL_FLUX, FVBND_EULER_IMPERMEABLE_WALL, FVBND_EULER_INVISCID_MIRROR = 0, 1, 2, 3
RK_EULER, RK_SSP2, RK_SSP3, RK_CLASSIC4, RK_OTHER = 0, 1, 2, 3, 4

# internal::ButcherArrayProvider (tools/timestepper/explicit-rungekutta.hh:63-141): A (row-major), b, c
BUTCHER = {
    RK_EULER: ([[0.0]], [1.0], [0.0]),
    RK_SSP2: ([[0.0, 0.0], [1.0, 0.0]], [0.5, 0.5], [0.0, 1.0]),
    RK_SSP3: ([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.25, 0.25, 0.0]], [1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0], [0.0, 1.0, 0.5]),
    RK_CLASSIC4: (
        [[0.0, 0.0, 0.0, 0.0], [0.5, 0.0, 0.0, 0.0], [0.0, 0.5, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0]],
        [1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0],
        [0.0, 0.5, 0.5, 1.0],
    ),
}


def fv_boundary(kind, side_mask, a, b):
    t = FvBoundary()
    t.kind, t.side_mask, t.a, t.b = int(kind), int(side_mask), float(a), float(b)
    return t


def solver_opts(type=SOLVER_CG, preconditioner=PRECOND_JACOBI, precision=1e-10, max_iter=0, check_every=0):
    o = SolverOpts()
    o.type, o.preconditioner, o.max_iter, o.check_every, o.precision = type, preconditioner, max_iter, check_every, precision
    return o
