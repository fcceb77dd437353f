"""ctypes front end of oracle/softdtw_oracle.c (test infrastructure only; see that file's header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libsoftdtw_oracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = ctypes.CDLL(so)
        dp, i, d = ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_double
        _lib.sdtw_forward.argtypes = [dp, i, i, i, d, d, dp]
        _lib.sdtw_backward.argtypes = [dp, dp, i, i, i, d, d, dp]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def forward(D, gamma, bandwidth=0.0):
    """D [B,N,M] -> padded R [B,N+2,M+2] float64 (compute_softdtw, soft_dtw_cuda.py:184-206)."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    B, N, M = D.shape
    R = np.empty((B, N + 2, M + 2), dtype=np.float64)
    _load().sdtw_forward(_p(D), B, N, M, float(gamma), float(bandwidth), _p(R))
    return R


def backward(D, R, gamma, bandwidth=0.0):
    """-> E [B,N,M] float64 (compute_softdtw_backward, :209-239); R is copied, not modified."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    R = np.array(R, dtype=np.float64, copy=True, order="C")
    B, N, M = D.shape
    E = np.empty((B, N, M), dtype=np.float64)
    _load().sdtw_backward(_p(D), _p(R), B, N, M, float(gamma), float(bandwidth), _p(E))
    return E


def sqeuclidean(X, Y):
    """SoftDTW._euclidean_dist_func (:319-329): [B,N,d] x [B,M,d] -> [B,N,M]."""
    X, Y = np.asarray(X), np.asarray(Y)
    return ((X[:, :, None, :] - Y[:, None, :, :]) ** 2).sum(-1)
