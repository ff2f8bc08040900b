"""Python loader of the CPU restatement oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing in dune-gdt_b200/ (the product) imports this package.
Descriptor layouts are shared with include/gdtb.h, so the ctypes classes of dune_gdt_b200.descriptors are reused.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        from dune_gdt_b200.descriptors import Flux, Form, Function, FvBoundary, GridDesc, Integrand

        if not os.path.exists(LIB_PATH):
            build()
        h = C.CDLL(LIB_PATH)
        P, DP = C.c_void_p, C.POINTER(C.c_double)
        I64P, I32P = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        G = C.POINTER(GridDesc)
        protos = {
            "orc_num_elements": (C.c_int64, [G]),
            "orc_space_size": (C.c_int64, [G, C.c_int, C.c_int]),
            "orc_space_local_size": (C.c_int32, [G, C.c_int, C.c_int]),
            "orc_space_global_indices": (None, [G, C.c_int, C.c_int, C.c_int64, I64P]),
            "orc_gauss_rule": (C.c_int32, [C.c_int, DP, DP]),
            "orc_shape_values": (None, [C.c_int, C.c_int, DP, DP]),
            "orc_shape_gradients": (None, [C.c_int, C.c_int, DP, DP]),
            "orc_pattern_create": (P, [G, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
            "orc_pattern_rows": (C.c_int64, [P]),
            "orc_pattern_nnz": (C.c_int64, [P]),
            "orc_pattern_copy": (None, [P, I64P, I32P]),
            "orc_pattern_free": (None, [P]),
            "orc_assemble": (
                C.c_int,
                [G, C.c_int, C.c_int, I64P, I32P, DP, C.c_int, C.POINTER(Form), C.c_int, C.POINTER(Form), C.c_int,
                 C.POINTER(Form), C.c_int, C.POINTER(Form), DP, C.c_int],
            ),
            "orc_local_element_matrix": (None, [G, C.c_int, C.c_int, C.POINTER(Form), C.c_int64, DP]),
            "orc_fv_apply": (C.c_int, [G, C.POINTER(Flux), DP, DP, C.c_int]),
            "orc_fv_euler": (C.c_int, [G, C.POINTER(Flux), DP, C.c_double, C.c_int64, C.c_int]),
            "orc_fv_interpolate": (None, [G, C.POINTER(Function), DP]),
            "orc_fv_apply_bnd": (C.c_int, [G, C.POINTER(Flux), C.c_int, C.POINTER(FvBoundary), DP, DP, C.c_int]),
            "orc_rk_step": (
                C.c_int,
                [G, C.POINTER(Flux), C.c_int, C.POINTER(FvBoundary), C.c_int, DP, DP, DP, C.c_double, DP, DP, C.c_double,
                 C.c_double, C.c_int],
            ),
            "orc_rk_solve": (
                C.c_int,
                [G, C.POINTER(Flux), C.c_int, C.POINTER(FvBoundary), C.c_int, DP, DP, DP, C.c_double, DP, C.c_double,
                 C.c_double, C.c_double, I64P, DP, C.c_int],
            ),
            "orc_fv_estimate_dt": (C.c_double, [G, C.POINTER(Flux), DP, DP]),
            "orc_function_eval": (C.c_double, [C.POINTER(Function), C.c_int, DP, C.c_int64]),
            "orc_last_error": (C.c_char_p, []),
            "orc_element_integrand_evaluate": (
                C.c_int, [C.POINTER(Integrand), C.c_int, C.c_int, DP, DP, C.c_int, DP, DP, DP, DP]),
            "orc_dirichlet_dofs": (C.c_int64, [G, C.c_int, C.c_int, C.c_uint32, I64P]),
            "orc_dirichlet_apply": (C.c_int, [C.c_int64, I64P, I32P, DP, DP, C.c_int64, I64P, C.c_int, C.c_int]),
            "orc_csr_mv": (None, [C.c_int64, I64P, I32P, DP, DP, DP]),
            "orc_bilinear_form_apply2": (C.c_double, [G, C.c_int, C.c_int, DP, C.POINTER(Function), C.POINTER(Form)]),
            "orc_lagrange_interpolate": (None, [G, C.c_int, C.c_int, C.POINTER(Function), DP]),
            "orc_fvsys_apply": (C.c_int, [G, C.POINTER(Flux), DP, DP]),
            "orc_fvsys_euler": (C.c_int, [G, C.POINTER(Flux), DP, C.c_double, C.c_int64]),
            "orc_fvsys_apply_walls": (C.c_int, [G, C.POINTER(Flux), C.c_uint32, C.c_uint32, DP, DP]),
            "orc_fvsys_estimate_dt": (C.c_double, [G, C.POINTER(Flux), DP]),
            "orc_euler_flux": (None, [C.c_int, C.c_double, DP, DP]),
            "orc_euler_jacobian": (None, [C.c_int, C.c_double, DP, DP]),
            "orc_euler_eigen": (None, [C.c_int, C.c_double, DP, DP, DP, DP, DP]),
        }
        for name, (res, args) in protos.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def space_size(grid, kind, order):
    return lib().orc_space_size(C.byref(grid), kind, order)


def local_size(grid, kind, order):
    return lib().orc_space_local_size(C.byref(grid), kind, order)


def global_indices(grid, kind, order, element):
    out = np.zeros(local_size(grid, kind, order), dtype=np.int64)
    lib().orc_space_global_indices(C.byref(grid), kind, order, element, out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def gauss_rule(order):
    x, w = np.zeros(8), np.zeros(8)
    m = lib().orc_gauss_rule(order, _dp(x), _dp(w))
    return x[:m], w[:m]


def pattern(grid, test=(0, 1), ansatz=None, stencil=0):
    """returns (rowptr int64, colidx int32); test/ansatz = (kind, order)"""
    ansatz = ansatz or test
    p = lib().orc_pattern_create(C.byref(grid), test[0], test[1], ansatz[0], ansatz[1], stencil)
    rows, nnz = lib().orc_pattern_rows(p), lib().orc_pattern_nnz(p)
    rowptr, colidx = np.empty(rows + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32)
    lib().orc_pattern_copy(p, rowptr.ctypes.data_as(C.POINTER(C.c_int64)), colidx.ctypes.data_as(C.POINTER(C.c_int32)))
    lib().orc_pattern_free(p)
    return rowptr, colidx


def _forms(forms):
    from dune_gdt_b200.descriptors import Form

    forms = list(forms or [])
    arr = (Form * max(len(forms), 1))()
    for i, f in enumerate(forms):
        C.memmove(C.byref(arr[i]), C.byref(f), C.sizeof(Form))
    arr._refs = forms
    return len(forms), arr


def assemble(grid, kind, order, rowptr, colidx, element_forms=(), coupling_forms=(), boundary_forms=(), rhs_forms=(),
             num_threads=1):
    """one grid walk; returns (values[nnz], rhs[ndof])"""
    values = np.zeros(len(colidx), dtype=np.float64)
    rhs = np.zeros(len(rowptr) - 1, dtype=np.float64)
    ne, ef = _forms(element_forms)
    nc, cf = _forms(coupling_forms)
    nb, bf = _forms(boundary_forms)
    nr, rf = _forms(rhs_forms)
    st = lib().orc_assemble(
        C.byref(grid), kind, order, rowptr.ctypes.data_as(C.POINTER(C.c_int64)),
        colidx.ctypes.data_as(C.POINTER(C.c_int32)), _dp(values), ne, ef, nc, cf, nb, bf, nr, rf, _dp(rhs), num_threads)
    if st != 0:
        raise RuntimeError(lib().orc_last_error().decode())
    return values, rhs


def local_element_matrix(grid, kind, order, form, element):
    n = local_size(grid, kind, order)
    out = np.zeros((n, n))
    lib().orc_local_element_matrix(C.byref(grid), kind, order, C.byref(form), element, _dp(out))
    return out


def fv_apply(grid, flux, u, num_threads=1):
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty_like(u)
    lib().orc_fv_apply(C.byref(grid), C.byref(flux), _dp(u), _dp(out), num_threads)
    return out


def fv_euler(grid, flux, u, dt, n_steps, num_threads=1):
    u = np.array(u, dtype=np.float64, copy=True)
    lib().orc_fv_euler(C.byref(grid), C.byref(flux), _dp(u), dt, n_steps, num_threads)
    return u


def _bnd_array(boundary):
    from dune_gdt_b200.descriptors import FvBoundary

    arr = (FvBoundary * max(1, len(boundary)))()
    for i, t in enumerate(boundary):
        arr[i] = t
    return arr


def fv_apply_bnd(grid, flux, boundary, u, num_threads=1):
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty_like(u)
    lib().orc_fv_apply_bnd(C.byref(grid), C.byref(flux), len(boundary), _bnd_array(boundary), _dp(u), _dp(out), num_threads)
    return out


def _butcher(A, b, c):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    assert A.shape == (b.size, b.size) and c.size == b.size
    return A, b, c


def rk_step(grid, flux, butcher, u, t, dt, max_dt=None, r=1.0, boundary=(), num_threads=1):
    """ExplicitRungeKuttaTimeStepper::step; returns (u_new, t_new)"""
    A, b, c = _butcher(*butcher)
    u = np.array(u, dtype=np.float64, copy=True)
    tt = np.array([t], dtype=np.float64)
    lib().orc_rk_step(C.byref(grid), C.byref(flux), len(boundary), _bnd_array(boundary), b.size, _dp(A), _dp(b), _dp(c),
                      r, _dp(u), _dp(tt), dt, dt if max_dt is None else max_dt, num_threads)
    return u, float(tt[0])


def rk_solve(grid, flux, butcher, u, t_end, initial_dt, t0=0.0, r=1.0, boundary=(), num_threads=1):
    """TimeStepperInterface::solve; returns (u(t_end), n_steps, t_final)"""
    A, b, c = _butcher(*butcher)
    u = np.array(u, dtype=np.float64, copy=True)
    n = np.zeros(1, dtype=np.int64)
    tf = np.zeros(1, dtype=np.float64)
    lib().orc_rk_solve(C.byref(grid), C.byref(flux), len(boundary), _bnd_array(boundary), b.size, _dp(A), _dp(b), _dp(c),
                       r, _dp(u), t0, t_end, initial_dt, n.ctypes.data_as(C.POINTER(C.c_int64)), _dp(tf), num_threads)
    return u, int(n[0]), float(tf[0])


def fv_estimate_dt(grid, flux, u, boundary_data_range=None):
    u = np.ascontiguousarray(u, dtype=np.float64)
    rng = None if boundary_data_range is None else np.ascontiguousarray(boundary_data_range, dtype=np.float64)
    return lib().orc_fv_estimate_dt(C.byref(grid), C.byref(flux), _dp(u), None if rng is None else _dp(rng))


def fv_interpolate(grid, function):
    u = np.empty(lib().orc_num_elements(C.byref(grid)), dtype=np.float64)
    lib().orc_fv_interpolate(C.byref(grid), C.byref(function), _dp(u))
    return u


def _i64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def dirichlet_dofs(grid, kind, order, boundary_mask=0x3F):
    n = lib().orc_dirichlet_dofs(C.byref(grid), kind, order, boundary_mask, None)
    out = np.empty(n, dtype=np.int64)
    lib().orc_dirichlet_dofs(C.byref(grid), kind, order, boundary_mask, _i64p(out))
    return out


def dirichlet_apply(rowptr, colidx, values, vector, dofs, only_clear=False, ensure_symmetry=True):
    """in place on copies; returns (values, vector)"""
    values = None if values is None else np.array(values, dtype=np.float64, copy=True)
    vector = None if vector is None else np.array(vector, dtype=np.float64, copy=True)
    dofs = np.ascontiguousarray(dofs, dtype=np.int64)
    st = lib().orc_dirichlet_apply(
        len(rowptr) - 1, _i64p(rowptr), colidx.ctypes.data_as(C.POINTER(C.c_int32)),
        _dp(values) if values is not None else None, _dp(vector) if vector is not None else None, len(dofs), _i64p(dofs),
        int(only_clear), int(ensure_symmetry))
    if st != 0:
        raise RuntimeError(lib().orc_last_error().decode())
    return values, vector


def csr_mv(rowptr, colidx, values, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(len(rowptr) - 1, dtype=np.float64)
    lib().orc_csr_mv(len(rowptr) - 1, _i64p(rowptr), colidx.ctypes.data_as(C.POINTER(C.c_int32)), _dp(values), _dp(x), _dp(y))
    return y


def bilinear_form_apply2(grid, kind, order, dofs, f, form):
    dofs = None if dofs is None else np.ascontiguousarray(dofs, dtype=np.float64)
    return lib().orc_bilinear_form_apply2(C.byref(grid), kind, order, _dp(dofs) if dofs is not None else None,
                                          C.byref(f) if f is not None else None, C.byref(form))


def lagrange_interpolate(grid, kind, order, f):
    out = np.empty(space_size(grid, kind, order), dtype=np.float64)
    lib().orc_lagrange_interpolate(C.byref(grid), kind, order, C.byref(f), _dp(out))
    return out


def element_integrand_evaluate(integrand, dim, test_values, test_grads, ansatz_values, ansatz_grads, x):
    """LocalLaplaceIntegrand / LocalElementProductIntegrand::evaluate on caller-supplied bases at one point; returns
    result[n_test, n_ansatz] (the reference's evaluates_correctly_for_scalar_bases tests)"""
    tv = np.ascontiguousarray(test_values, dtype=np.float64)
    tg = np.ascontiguousarray(test_grads, dtype=np.float64).reshape(len(tv), dim)
    av = np.ascontiguousarray(ansatz_values, dtype=np.float64)
    ag = np.ascontiguousarray(ansatz_grads, dtype=np.float64).reshape(len(av), dim)
    xx = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros((len(tv), len(av)))
    st = lib().orc_element_integrand_evaluate(C.byref(integrand), dim, len(tv), _dp(tv), _dp(tg), len(av), _dp(av),
                                              _dp(ag), _dp(xx), _dp(out))
    if st != 0:
        raise RuntimeError(lib().orc_last_error().decode())
    return out


# ---- systems of conservation laws: the Euler equations (tools/euler.hh), m = d + 2 components per cell ------------------
def fvsys_apply(grid, flux, u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty_like(u)
    assert lib().orc_fvsys_apply(C.byref(grid), C.byref(flux), _dp(u), _dp(out)) == 0
    return out


def fvsys_apply_walls(grid, flux, u, wall_mask=0, mirror_mask=0):
    """impermeable walls on the domain sides of the masks (bit 2k + s): wall flux / mirrored ghost state"""
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty_like(u)
    assert lib().orc_fvsys_apply_walls(C.byref(grid), C.byref(flux), wall_mask, mirror_mask, _dp(u), _dp(out)) == 0
    return out


def fvsys_euler(grid, flux, u, dt, n_steps):
    u = np.array(u, dtype=np.float64, copy=True)
    assert lib().orc_fvsys_euler(C.byref(grid), C.byref(flux), _dp(u), dt, n_steps) == 0
    return u


def fvsys_estimate_dt(grid, flux, u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    return lib().orc_fvsys_estimate_dt(C.byref(grid), C.byref(flux), _dp(u))


def euler_flux(d, gamma, w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    f = np.empty((d, d + 2))
    lib().orc_euler_flux(d, gamma, _dp(w), _dp(f))
    return f


def euler_jacobian(d, gamma, w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    J = np.empty((d, d + 2, d + 2))
    lib().orc_euler_jacobian(d, gamma, _dp(w), _dp(J))
    return J


def euler_eigen(d, gamma, w, n):
    """(eigenvalues, T, T^{-1}) of sum_s n_s A_s(w) (EulerTools::*_flux_jacobian)"""
    w = np.ascontiguousarray(w, dtype=np.float64)
    n = np.ascontiguousarray(n, dtype=np.float64)
    m = d + 2
    ev, T, Ti = np.empty(m), np.empty((m, m)), np.empty((m, m))
    lib().orc_euler_eigen(d, gamma, _dp(w), _dp(n), _dp(ev), _dp(T), _dp(Ti))
    return ev, T, Ti


def euler_conservative(gamma, rho, v, p):
    """EulerTools::conservative (tools/euler.hh:153-157)"""
    v = np.atleast_1d(np.asarray(v, dtype=np.float64))
    return np.concatenate([[rho], rho * v, [p / (gamma - 1.0) + 0.5 * rho * float(v @ v)]])
