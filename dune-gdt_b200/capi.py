"""ctypes binding of libgdtb.so (include/gdtb.h).

Loading never falls back to anything: if the shared library is missing, or a compute call is made
without a CUDA device, an exception is raised.
"""
import ctypes as C
import os

from .descriptors import Flux, Form, Function, FvBoundary, GridDesc, SolverInfo, SolverOpts

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GDTB_LIB") or os.path.join(_HERE, "lib", "libgdtb.so")  # GDTB_LIB: A/B builds

# status codes -> exception classes named after the reference's (dune/gdt/exceptions.hh:24-75)


class GdtError(RuntimeError):
    """Dune::Exception"""


class WrongInputGiven(GdtError):
    """XT::Common::Exceptions::wrong_input_given"""


class ShapesDoNotMatch(GdtError):
    """XT::Common::Exceptions::shapes_do_not_match"""


class IntegrandError(GdtError):
    """Dune::GDT::Exceptions::integrand_error"""


class FiniteElementError(GdtError):
    """Dune::GDT::Exceptions::finite_element_error"""


class SpaceError(GdtError):
    """Dune::GDT::Exceptions::space_error"""


class OperatorError(GdtError):
    """Dune::GDT::Exceptions::operator_error"""


class NotImplementedGdt(GdtError):
    """Dune::NotImplemented"""


class CudaError(GdtError):
    """CUDA runtime failure or no device; there is no CPU fallback."""


class OutOfMemory(GdtError):
    pass


_ERRORS = {
    1: WrongInputGiven,
    2: ShapesDoNotMatch,
    3: IntegrandError,
    4: FiniteElementError,
    5: SpaceError,
    6: OperatorError,
    7: NotImplementedGdt,
    8: CudaError,
    9: OutOfMemory,
}

_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_DP = C.POINTER(C.c_double)
_I64P = C.POINTER(C.c_int64)
_I32P = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/gdtb.h declares
PROTOTYPES = {
    "gdtb_last_error": (C.c_char_p, []),
    "gdtb_version": (C.c_int, []),
    "gdtb_ctx_create": (C.c_int, [C.c_int, _PP]),
    "gdtb_ctx_destroy": (C.c_int, [_P]),
    "gdtb_ctx_set_stream": (C.c_int, [_P, _P]),
    "gdtb_ctx_synchronize": (C.c_int, [_P]),
    "gdtb_ctx_kernel_name": (C.c_char_p, [_P, C.c_char_p]),
    "gdtb_ctx_launch_count": (C.c_int64, [_P]),
    "gdtb_ctx_enable_timing": (C.c_int, [_P, C.c_int]),
    "gdtb_ctx_kernel_time": (C.c_int, [_P, C.c_char_p, _DP, _I64P]),
    "gdtb_grid_create_cube": (C.c_int, [_P, C.POINTER(GridDesc), _PP]),
    "gdtb_grid_destroy": (C.c_int, [_P]),
    "gdtb_grid_num_elements": (C.c_int64, [_P]),
    "gdtb_space_create": (C.c_int, [_P, _P, C.c_int, C.c_int, _PP]),
    "gdtb_fv_space_create": (C.c_int, [_P, _P, C.c_int, _PP]),
    "gdtb_space_destroy": (C.c_int, [_P]),
    "gdtb_space_size": (C.c_int64, [_P]),
    "gdtb_space_max_local_size": (C.c_int32, [_P]),
    "gdtb_space_global_indices": (C.c_int, [_P, C.c_int64, _I64P]),
    "gdtb_form_quadrature_order": (C.c_int, [_P, C.POINTER(Form), C.c_int, _I32P]),
    "gdtb_gauss_rule": (C.c_int, [C.c_int, _I32P, _DP, _DP]),
    "gdtb_pattern_create": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _PP]),
    "gdtb_pattern_destroy": (C.c_int, [_P]),
    "gdtb_pattern_rows": (C.c_int64, [_P]),
    "gdtb_pattern_cols": (C.c_int64, [_P]),
    "gdtb_pattern_nnz": (C.c_int64, [_P]),
    "gdtb_pattern_download": (C.c_int, [_P, _I64P, _I32P]),
    "gdtb_pattern_device": (C.c_int, [_P, _PP, _PP]),
    "gdtb_matop_create": (C.c_int, [_P, _P, _P, _P, _PP]),
    "gdtb_matop_destroy": (C.c_int, [_P]),
    "gdtb_matop_append_element": (C.c_int, [_P, C.POINTER(Form)]),
    "gdtb_matop_append_coupling": (C.c_int, [_P, C.POINTER(Form), C.c_int]),
    "gdtb_matop_append_boundary": (C.c_int, [_P, C.POINTER(Form), C.c_int]),
    "gdtb_matop_clear_forms": (C.c_int, [_P]),
    "gdtb_matop_num_forms": (C.c_int, [_P]),
    "gdtb_matop_plan": (C.c_char_p, [_P]),
    "gdtb_matop_plan_reason": (C.c_char_p, [_P]),
    "gdtb_matop_set_zero": (C.c_int, [_P]),
    "gdtb_matop_values_download": (C.c_int, [_P, _DP]),
    "gdtb_matop_values_upload": (C.c_int, [_P, _DP]),
    "gdtb_matop_values_device": (C.c_int, [_P, _PP]),
    "gdtb_matop_set_values_device": (C.c_int, [_P, _P]),
    "gdtb_matop_set_slab": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "gdtb_vecfun_create": (C.c_int, [_P, _P, _PP]),
    "gdtb_vecfun_destroy": (C.c_int, [_P]),
    "gdtb_vecfun_append_element": (C.c_int, [_P, C.POINTER(Form)]),
    "gdtb_vecfun_clear_forms": (C.c_int, [_P]),
    "gdtb_vecfun_set_zero": (C.c_int, [_P]),
    "gdtb_vecfun_download": (C.c_int, [_P, _DP]),
    "gdtb_vecfun_device": (C.c_int, [_P, _PP]),
    "gdtb_vecfun_set_device": (C.c_int, [_P, _P]),
    "gdtb_vecfun_set_slab": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "gdtb_assemble": (C.c_int, [_P, _P, C.c_int]),
    "gdtb_assemble_async": (C.c_int, [_P, _P, C.c_int]),
    "gdtb_matop_local_nnz": (C.c_int64, [_P]),
    "gdtb_host_closed_form_rowptr": (C.c_int, [C.POINTER(GridDesc), C.c_int, C.c_int, _I64P]),
    "gdtb_host_slab_row_ranges": (C.c_int, [C.POINTER(GridDesc), C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int32, _I64P, _I64P,
                                  _I64P, _I64P, C.POINTER(C.c_int32)]),
    "gdtb_matop_set_slab_halo": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "gdtb_vecfun_set_slab_halo": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "gdtb_matop_halo_layout": (C.c_int, [_P, _I64P, _I64P, _I64P]),
    "gdtb_vecfun_halo_layout": (C.c_int, [_P, _I64P, _I64P, _I64P]),
    "gdtb_halo_p2p_alloc": (C.c_int, [_P, _P]),
    "gdtb_halo_p2p_connect": (C.c_int, [_P, _P, C.c_int64, _P]),
    "gdtb_halo_p2p_check": (C.c_int, [_P]),
    "gdtb_vector_add": (C.c_int, [_P, _P, _P, C.c_int64]),
    "gdtb_vector_upload": (C.c_int, [_P, _P, _P, C.c_int64]),
    "gdtb_vector_download": (C.c_int, [_P, _P, _P, C.c_int64]),
    "gdtb_matop_local_rows": (C.c_int, [_P, _I64P, _I64P, _I64P]),
    "gdtb_matop_local_row_ranges": (C.c_int, [_P, C.c_int32, _I64P, _I64P, _I64P, _I64P, C.POINTER(C.c_int32)]),
    "gdtb_assemble_host": (C.c_int, [_P, _P, _DP, _DP]),
    "gdtb_fvop_create": (C.c_int, [_P, _P, C.POINTER(Flux), _PP]),
    "gdtb_fvop_destroy": (C.c_int, [_P]),
    "gdtb_fvop_apply": (C.c_int, [_P, _P, _P]),
    "gdtb_fvop_apply_host": (C.c_int, [_P, _DP, _DP]),
    "gdtb_fvop_euler": (C.c_int, [_P, _P, C.c_double, C.c_int64]),
    "gdtb_fvop_euler_host": (C.c_int, [_P, _DP, C.c_double, C.c_int64]),
    "gdtb_fvop_set_slab": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "gdtb_fvop_ghost_layer_size": (C.c_int64, [_P]),
    "gdtb_fvop_p2p_alloc": (C.c_int, [_P, _PP, _PP, _P]),
    "gdtb_fvop_p2p_connect": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.c_int64, C.c_int]),
    "gdtb_fvop_p2p_step": (C.c_int, [_P, C.c_int, C.c_double]),
    "gdtb_fvop_p2p_current": (C.c_int, [_P, _PP, _I64P]),
    "gdtb_fvop_p2p_check": (C.c_int, [_P]),
    "gdtb_fvop_step_async": (C.c_int, [_P, _P, _P, C.c_int, C.c_double, C.c_int64, C.c_int64]),
    "gdtb_fvop_append_boundary": (C.c_int, [_P, C.POINTER(FvBoundary)]),
    "gdtb_fv_estimate_dt": (C.c_int, [_P, _P, _DP, _DP]),
    "gdtb_fv_estimate_dt_host": (C.c_int, [_P, _DP, _DP, _DP]),
    "gdtb_rk_create": (C.c_int, [_P, C.c_int, C.c_int, _DP, _DP, _DP, C.c_double, C.c_double, _PP]),
    "gdtb_rk_destroy": (C.c_int, [_P]),
    "gdtb_rk_current_time": (C.c_double, [_P]),
    "gdtb_rk_step": (C.c_int, [_P, _P, C.c_double, C.c_double, _DP]),
    "gdtb_rk_step_host": (C.c_int, [_P, _DP, C.c_double, C.c_double, _DP]),
    "gdtb_rk_solve": (C.c_int, [_P, _P, C.c_double, C.c_double, _I64P, _DP]),
    "gdtb_rk_solve_host": (C.c_int, [_P, _DP, C.c_double, C.c_double, _I64P, _DP]),
    "gdtb_rk_p2p_handles": (C.c_int, [_P, _PP, _P]),
    "gdtb_rk_p2p_connect": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.c_int64, C.c_int]),
    "gdtb_rk_p2p_check": (C.c_int, [_P]),
    "gdtb_fv_interpolate": (C.c_int, [_P, _P, C.POINTER(Function), _P]),
    "gdtb_fv_interpolate_host": (C.c_int, [_P, _P, C.POINTER(Function), _DP]),
    # callers on either side of the hot path
    "gdtb_dirichlet_create": (C.c_int, [_P, _P, C.c_uint32, _PP]),
    "gdtb_dirichlet_destroy": (C.c_int, [_P]),
    "gdtb_dirichlet_size": (C.c_int64, [_P]),
    "gdtb_dirichlet_dofs_download": (C.c_int, [_P, _I64P]),
    "gdtb_dirichlet_device": (C.c_int, [_P, _PP, _PP]),
    "gdtb_dirichlet_apply": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "gdtb_dirichlet_apply_device": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int]),
    "gdtb_matop_apply": (C.c_int, [_P, _P, _P]),
    "gdtb_matop_apply_host": (C.c_int, [_P, _DP, _DP]),
    "gdtb_csr_apply_device": (C.c_int, [_P, _P, _P, _P, _P]),
    "gdtb_matop_apply_inverse": (C.c_int, [_P, _P, _P, C.POINTER(SolverOpts), C.POINTER(SolverInfo)]),
    "gdtb_matop_apply_inverse_host": (C.c_int, [_P, _DP, _DP, C.POINTER(SolverOpts), C.POINTER(SolverInfo)]),
    "gdtb_csr_apply_inverse_device": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(SolverOpts), C.POINTER(SolverInfo)]),
    "gdtb_matop_pattern_device": (C.c_int, [_P, _PP, _PP]),
    "gdtb_bilinear_form_quadrature_order": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(Form), _I32P]),
    "gdtb_bilinear_form_apply2": (C.c_int, [_P, _P, _P, C.POINTER(Function), C.POINTER(Form), _DP]),
    "gdtb_bilinear_form_apply2_host": (C.c_int, [_P, _P, _DP, C.POINTER(Function), C.POINTER(Form), _DP]),
    "gdtb_lagrange_interpolate": (C.c_int, [_P, _P, C.POINTER(Function), _P]),
    "gdtb_lagrange_interpolate_host": (C.c_int, [_P, _P, C.POINTER(Function), _DP]),
}

_lib = None


def lib():
    """The loaded libgdtb.so; raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        msg = lib().gdtb_last_error().decode("utf-8", "replace")
        raise _ERRORS.get(status, GdtError)(msg)


def dptr(array):
    """double* of a contiguous float64 numpy array"""
    return array.ctypes.data_as(_DP)
