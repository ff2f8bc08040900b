"""Shared helpers of the parity tests."""
import numpy as np


def rel_err(a, b):
    """norm-wise parity metric of SURVEY.md section 8(c): max|a-b| / max|b|"""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.max(np.abs(b))
    if scale == 0.0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / scale)


TOL = 1e-12  # north_star: matrix, RHS and FV update values within 1e-12 relative in FP64
